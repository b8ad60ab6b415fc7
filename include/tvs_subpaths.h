/*
 * tvs_subpaths.h -- C ABI of the read-path x variant occurrence counting,
 *                   part of "traversome_b200/libtvs.so".
 *
 * Replaces the per-variant Python loop of the reference:
 *   VariantSubPathsGenerator.gen_subpaths      traversome/utils.py:416-501
 *   Traversome.gen_all_informative_sub_paths   traversome/traversome.py:1210-1263 (the counting part)
 * which, for every start position of a variant path, grows the longest window whose accumulated length
 * (contig length minus the overlap of the edge entered, summed over the appended contigs) stays <= max_alignment_len-2,
 * standardises each prefix of that window (Assembly.get_standardized_path / get_standardized_path_circ,
 * traversome/Assembly.py:1174-1202) and counts those that are observed read paths.  The result is the integer
 * matrix A[read path, variant] (`SubPathInfo.from_variants`) that tvs_create consumes: a bit-exact contract.
 *
 * The graph never crosses the boundary.  The caller encodes
 *   - every variant as a sequence of oriented-contig tokens (any injective int32 code >= 0 of (contig, strand), with
 *     the strand of palindromic contigs forced to forward: Assembly.correct_path_with_palindromic_repeats
 *     Assembly.py:643-653) and, per position, the step = len(contig) - overlap(previous contig -> this contig)
 *     (for position 0 of a circular variant the previous contig is the last one; unused for linear variants);
 *   - every read path as the token sequences ("keys") a window must equal to be standardised onto it: the path and
 *     its reverse complement and, for windows of linear variants that are themselves circular paths, all rotations
 *     (see traversome_b200/subpaths.py, which builds them).  A key carries two rows because the two standardisers
 *     can send the same window to different read paths: key_row_circular[k] = the read path a window of a CIRCULAR
 *     variant equal to key k stands for, key_row_linear[k] = likewise for windows of LINEAR variants; -1 = none.
 *     Keys must be pairwise distinct sequences.
 *
 * Conventions are those of tvs.h: 0 or a negative TVS_E* code, tvs_last_error() for the message, no CPU fallback.
 * var_ptr and key_ptr are validated on the host and must be HOST arrays; the other inputs may be host or device.
 * Outputs of tvs_subpaths_fetch are written to HOST or DEVICE memory.
 */
#ifndef TVS_SUBPATHS_H_
#define TVS_SUBPATHS_H_

#include <stdint.h>
#include "tvs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvs_subpath_counts* tvs_subpaths_handle;

/* Counts, on `device`, the occurrences of every read path in every variant.
 *   var_ptr[n_variants+1], var_tok[var_ptr[n_variants]], var_step[same], var_circular[n_variants]
 *   n_rows = number of read paths; key_ptr[n_keys+1], key_tok[key_ptr[n_keys]],
 *   key_row_circular[n_keys] and key_row_linear[n_keys] in [-1, n_rows)
 *   max_internal_len = max_alignment_len - 2 (utils.py:452, :479)
 * On success *out owns the result; *nnz_out = number of non-zero (read path, variant) pairs,
 * *n_windows_out = number of (start, prefix) windows examined (the unit of the throughput figure),
 * *kernel_ms_out = device time of the counting kernels (CUDA events); the last three may be null. */
int tvs_subpaths_count(tvs_subpaths_handle* out, int device,
                       int64_t n_variants, const int64_t* var_ptr, const int32_t* var_tok, const int32_t* var_step,
                       const uint8_t* var_circular,
                       int64_t n_rows, int64_t n_keys, const int64_t* key_ptr, const int32_t* key_tok,
                       const int32_t* key_row_circular, const int32_t* key_row_linear,
                       int64_t max_internal_len,
                       int64_t* nnz_out, int64_t* n_windows_out, double* kernel_ms_out);

/* CSR by read path of the counts held by `h` (the layout tvs_create takes):
 *   row_ptr[n_rows+1] (int64), col[nnz] = variant (increasing inside a row), val[nnz] = occurrences,
 *   first[nnz] = rank of the first window of that variant standardised onto that read path, in the reference's
 *   enumeration order (start position ascending, then window length descending: utils.py:441-470); it orders the
 *   keys of the reference's per-variant dict and hence `all_sub_paths` (traversome.py:1246-1263).
 * Any output pointer may be null to skip it. */
int tvs_subpaths_fetch(tvs_subpaths_handle h, int64_t* row_ptr, int32_t* col, int32_t* val, int64_t* first);

int tvs_subpaths_destroy(tvs_subpaths_handle h);

#ifdef __cplusplus
}
#endif
#endif  /* TVS_SUBPATHS_H_ */
