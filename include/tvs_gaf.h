/*
 * tvs_gaf.h -- C ABI of the alignment-file parser, part of "traversome_b200/libtvs.so".
 *              HOST code: byte work on the text of the file, no device is touched and none is needed.
 *
 * Replaces the per-line Python object construction of the reference:
 *   GAFRecord.__init__                     traversome/GraphAlignRecords.py:33-69   (12 mandatory GAF columns, optional
 *                                          tags "xx:i:", "xx:f:", "xx:Z:", identity = id tag or num_match / align_len)
 *   gaf_str_to_path                        traversome/utils.py:1065-1082           (">53<45>33", stable-path coordinates
 *                                          after ':' dropped, a step without '>' / '<' is forward)
 *   _gaf_parse_worker                      traversome/GraphAlignRecords.py:151-171 (strip, split on tabs, keep records with
 *                                          identity >= min_record_identity, count every line)
 *   GraphAlignRecords.parse_alignment_file(_single|_mul)   :1024-1134 (the file loop; the reference forks a process pool
 *                                          and pickles the record objects back -- here the lines are cut into blocks and
 *                                          parsed by threads into column arrays)
 *   SPATSVRecord.__init__ / _tsv_parse_worker             :89-148, :178-196 (SPAligner TSV) -- tvs_gaf_parse_spa
 *
 * The result is a TABLE in array form: one row per kept record, integer columns, and three string pools (query names,
 * segment names of the path steps, CIGAR strings).  Integers are a bit-exact contract; identity is the IEEE double the
 * reference computes (strtod of the tag, or num_match / (double)align_len).
 *
 * A line the reference would raise on (fewer than 12 columns, a non-integer where an integer is required, a strand that
 * is neither '+' nor '-', a non-numeric id tag) makes the call fail with TVS_EINVAL and the 1-based line number in
 * tvs_last_error() -- nothing is skipped silently.
 *
 * Conventions are those of tvs.h: 0 or a negative TVS_E* code, tvs_last_error() for the message.
 */
#ifndef TVS_GAF_H_
#define TVS_GAF_H_

#include <stdint.h>
#include "tvs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvs_gaf_table* tvs_gaf_handle;

/* Parse GAF text (`text[0..n_bytes)`, lines separated by '\n', a final line without '\n' counts).  Records with
 * identity < min_identity are dropped (GraphAlignRecords.py:169).  n_threads <= 0: one per core (at most 64).
 * keep_cigar != 0 keeps the "cg:Z:" strings (the reference's parse_cigar switch). */
int tvs_gaf_parse(const char* text, int64_t n_bytes, double min_identity, int n_threads, int keep_cigar, tvs_gaf_handle* out);

/* The same for SPAligner TSV (GraphAlignRecords.py:89-148): lines whose second column holds a ',' (several
 * non-overlapping sub-paths) are counted and skipped (:190-191).  Columns not in the format are returned as -1
 * (p_len, num_match, mapq) or NaN (identity); path_align_lengths of the records are the `extra` values of
 * tvs_gaf_steps. */
int tvs_gaf_parse_spa(const char* text, int64_t n_bytes, int n_threads, tvs_gaf_handle* out);

/* n_lines = lines read (n_file_records of the reference), n_records = records kept, n_steps = path steps of all kept
 * records, and the sizes in bytes of the three string pools. */
int tvs_gaf_dims(tvs_gaf_handle t, int64_t* n_lines, int64_t* n_records, int64_t* n_steps, int64_t* name_bytes,
                 int64_t* seg_bytes, int64_t* cigar_bytes);

/* Per-record columns, each [n_records] (ptr arrays [n_records + 1]); any pointer may be NULL.
 *   line_no    0-based line of the record in the file (its raw id before the identity filter)
 *   q_strand   1 = '+', 0 = '-'
 *   step_ptr   steps of record i = [step_ptr[i], step_ptr[i+1])
 *   name_ptr   query name of record i = names[name_ptr[i] .. name_ptr[i+1])
 *   cigar_ptr  CIGAR of record i (empty when absent or not kept)
 *   tag_off    byte offset in `text` of the first optional tag of the record's line (-1: none), tag_end: end of the line
 *              (stripped) -- the caller re-reads rarely used tags from the text it still holds */
int tvs_gaf_columns(tvs_gaf_handle t, int64_t* line_no, int64_t* query_len, int64_t* q_start, int64_t* q_end,
                    int8_t* q_strand, int64_t* p_len, int64_t* p_start, int64_t* p_end, int64_t* num_match,
                    int64_t* align_len, int64_t* mapq, double* identity, int64_t* step_ptr, int64_t* name_ptr,
                    int64_t* cigar_ptr, int64_t* tag_off, int64_t* tag_end);

/* String pools and per-step columns: names[name_bytes], step_forward[n_steps] (1 = '>' or no prefix, 0 = '<'),
 * seg_ptr[n_steps + 1] into segs[seg_bytes], extra[n_steps] (SPA: aligned length on the step; GAF: 0),
 * cigars[cigar_bytes].  Any pointer may be NULL. */
int tvs_gaf_steps(tvs_gaf_handle t, char* names, int8_t* step_forward, int64_t* seg_ptr, char* segs, int64_t* extra,
                  char* cigars);

/* Distinct segment names of the path steps, numbered in order of first appearance in the file (a graph has far fewer
 * segments than the file has steps): counts first, then step_segment[n_steps] (the id of every step), dseg_ptr[n_distinct + 1]
 * into dsegs[distinct_bytes].  tvs_gaf_segment_ids before tvs_gaf_segments is TVS_ESTATE. */
int tvs_gaf_segments(tvs_gaf_handle t, int64_t* n_distinct, int64_t* distinct_bytes);
int tvs_gaf_segment_ids(tvs_gaf_handle t, int64_t* step_segment, int64_t* dseg_ptr, char* dsegs);

int tvs_gaf_destroy(tvs_gaf_handle t);

#ifdef __cplusplus
}
#endif
#endif  /* TVS_GAF_H_ */
