/*
 * tvs_pool.h -- C ABI of the per-replicate multinomial bin statistics,
 *               part of "traversome_b200/libtvs.so".
 *
 * Replaces, for a whole batch of bootstrap / jackknife replicates at once, the per-replicate Python object graph of the
 * reference:
 *   Traversome.sample_sub_paths                    traversome/traversome.py:1636-1699  (regrouping of the drawn records by
 *                                                  read path, the (length, record id) order)
 *   Traversome.generate_multinomial_bin_stats      traversome/traversome.py:1305-1384
 *   Traversome._identify_bins                      traversome/traversome.py:1386-1462  (length ranges from the read paths'
 *                                                  shortest / longest possible alignment)
 *   _generate_align_len_id_lookup_table / _get_id_range_in_increasing_lengths   :1797-1842 (id range of a length range)
 *   get_averaged_possible_alignment_start_points   traversome/traversome.py:1492-1521  (num_possible_X = exact integer sum
 *   get_num_of_possible_alignment_start_points     traversome/traversome.py:1523-1579   of start points / number of records)
 *   the pruning rule of PathMultinomialModel.get_like_formula, traversome/ModelGenerator.py:138-151 (bins with X < 1)
 * and returns what the solver consumes per replicate ("lane"): the weight column w[r] = sum of num_matched over the
 * un-pruned bins of read path r, N = sum w, C = sum num_matched * log(num_possible_X), and the replicate's observed read
 * paths.  The random DRAWS are not made here: the caller replays the reference's `random.choices` / `random.sample`
 * (traversome.py:1652-1655, module-level RNG seeded by --rs) and passes the drawn record indices, so a run with the same
 * seed sees the same replicates.  Integer outputs are a bit-exact contract (w, N, observed; num_possible_X is an exact
 * integer sum divided once in IEEE fp64); C is a sum of n*log(X) in a fixed order (agrees with the host to rounding).
 *
 * Conventions are those of tvs.h: 0 or a negative TVS_E* code, tvs_last_error() for the message, no CPU fallback.
 * The arrays of tvs_pool_create are HOST arrays (they are indexed and sorted on the host once); samples / masked_rows and
 * the outputs of tvs_pool_lanes may be host or device pointers.
 */
#ifndef TVS_POOL_H_
#define TVS_POOL_H_

#include <stdint.h>
#include "tvs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvs_record_pool* tvs_pool_handle;

/* The record pool of one run (Traversome._prepare_for_sampling, traversome.py:1622-1634) in array form.
 *   record_ids[n_records]  strictly increasing ids of the valid records (records_pool_sorted)
 *   row_of[n_records]      read path (row of all_sub_paths) of each record
 *   align_len[n_records]   align_len_at_path_map[record]
 *   per read path r < n_rows: multi[r] = more than one contig; inner_len[r], full_len[r] (SubPathInfo);
 *   left_room[r] / right_room[r] = length of the first / last contig minus the overlap with its neighbour (:1558-1573) */
int tvs_pool_create(tvs_pool_handle* out, int device, int64_t n_records, int64_t n_rows, const int64_t* record_ids,
                    const int32_t* row_of, const int64_t* align_len, const uint8_t* multi, const int64_t* inner_len,
                    const int64_t* full_len, const int64_t* left_room, const int64_t* right_room);

/* Lanes of B replicates.  samples[B * n_draw]: replicate b drew the records samples[b*n_draw .. +n_draw) (indices into the
 * pool, duplicates allowed, -1 = no draw: pads a shorter jackknife sample); masked_rows[n_rows] u8 or NULL
 * (read_paths_masked).  Outputs, lane-minor like every per-lane matrix of tvs.h: w_out[n_rows * B] int32,
 * observed_out[n_rows * B] u8 (the replicate's all_sub_paths: sampled, not masked, legal), C_out[B], N_out[B].
 * Any output may be NULL. */
int tvs_pool_lanes(tvs_pool_handle p, int64_t B, int64_t n_draw, const int32_t* samples, const uint8_t* masked_rows,
                   int32_t* w_out, double* C_out, int64_t* N_out, uint8_t* observed_out);

int tvs_pool_destroy(tvs_pool_handle p);

#ifdef __cplusplus
}
#endif
#endif /* TVS_POOL_H_ */
