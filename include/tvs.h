/*
 * tvs.h -- C ABI of the B200-native maximum-likelihood variant-frequency fitter
 *          ("traversome_b200/libtvs.so").
 *
 * The reference (JianjunJin/Traversome v0.1.7.1) is pure Python and has no FFI for this
 * path; the drop-in boundary is its Python class `ModelFitMaxLike`
 * (traversome/ModelFitMaxLike.py:61-580).  The entry points below are what that class's
 * arithmetic is replaced by; `traversome_b200/ModelFitMaxLike.py` is the only caller and
 * INTEGRATION.md shows the ctypes stub a maintainer would add to the reference.
 *
 * Math (traversome/ModelGenerator.py:129-176), for one batch "lane" b (= one bootstrap
 * replicate and/or one candidate-variant subset):
 *
 *     LL_b(theta) = C_b + sum_r w[r,b] * log( (sum_v A[r,v]*theta[v,b]) / (sum_v L[v]*theta[v,b]) )
 *
 *   A  R x V  integer read-path x variant occurrence counts (`from_variants`), CSR, shared by all lanes
 *   L  V      integer variant sizes (`variant_sizes`)
 *   w  R x B  integer observed records per read path and lane (sum of `num_matched` over the
 *             lane's bins of that read path)
 *   C  B      sum over the lane's bins of num_matched*log(num_possible_X)   (constant in theta)
 *   mask V x B  1 = variant v is in lane b's candidate subset (`within_variant_ids`)
 *
 * Conventions
 *   - every function returns 0 on success or a negative TVS_E* code; nothing throws across the
 *     boundary; `tvs_last_error()` returns a thread-local message for the last failure.
 *   - array arguments may be HOST or DEVICE pointers unless stated otherwise (copies use
 *     cudaMemcpyDefault); all per-lane matrices are "lane-minor": element (i, b) at [i*B + b].
 *   - a handle owns one CUDA stream on one device; calls on one handle are serialised by the
 *     caller; different handles are independent (re-entrant per handle).
 *   - there is NO CPU fallback: without a usable CUDA device tvs_create fails with TVS_ECUDA.
 *   - output arrays of tvs_fit / tvs_loglik may be DEVICE pointers too: a multi-GPU caller gathers the per-lane results
 *     with NCCL straight from them (traversome_b200/lanes.py: LaneSharder).
 *   - candidate sets too wide for the shared-memory tiles of the staged kernel (V > ~330) and sparse (< 22 % of A
 *     non-zero) run the blocked kernel of csrc/tvs_blk.cuh: register accumulators, operand tiles fetched by bulk
 *     asynchronous copies (TMA); TVS_BLK=0 in the environment forbids it (falls back to the round-1 L2-gather kernel).
 *   - shapes with more than ~400 variants of which at least 22 % of A is non-zero keep a dense fp64 copy of A
 *     (R*V*8 bytes) and run the likelihood pass as two cuBLAS DGEMMs (the library links libcublas); TVS_DENSE=0
 *     in the environment forbids that, TVS_DENSE=1 forces it.
 */
#ifndef TVS_H_
#define TVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvs_problem* tvs_handle;

enum {
    TVS_OK = 0,
    TVS_EINVAL = -1,   /* bad argument (null pointer, negative size, unsorted CSR, ...)          */
    TVS_ECUDA = -2,    /* CUDA runtime error or no device                                         */
    TVS_ENOMEM = -3,   /* host or device allocation failed                                        */
    TVS_ELIMIT = -4,   /* V > 65535, a count A[r,v] > 65535 or a chunk too large for this build   */
    TVS_ESTATE = -5    /* call order violated (e.g. tvs_fit before tvs_set_lanes)                 */
};

/* per-lane status written by tvs_fit */
enum {
    TVS_CONVERGED = 0,   /* KKT residuals below tol                                               */
    TVS_MAXITER = 1,     /* max_iter reached; best point so far returned                          */
    TVS_STALLED = 2,     /* line search could not improve at fp64 resolution; residuals returned  */
    TVS_INFEASIBLE = 3   /* some read path with w>0 is covered by no variant of the lane's subset
                            (the reference's objective is -inf: ModelFitMaxLike.py:560-580 guards
                            it with cover_all_observed_subpaths)                                  */
};

typedef struct tvs_fit_options {
    double tol;          /* KKT tolerance: max_v phi_v*|g_v-1| and max_{v: phi_v~0} (g_v-1); <=0 -> 1e-9    */
    int32_t max_iter;    /* maximum accepted L-BFGS iterations per lane; <=0 -> 1000                        */
    int32_t history;     /* L-BFGS memory (4, 8 or 16); <=0 -> 16                                                */
    int32_t check_every; /* host polls the device-side "active lanes" counter every k passes; <=0 -> 4      */
    int32_t reserved;
} tvs_fit_options;

typedef struct tvs_fit_stats {
    int64_t passes;          /* fused likelihood/gradient passes launched by the last tvs_fit             */
    int64_t kernel_launches; /* all kernels launched by the last tvs_fit                                  */
    double  pass_ms;         /* CUDA-event time spent in the pass kernel (sum over passes)                */
    double  total_ms;        /* CUDA-event time of the whole tvs_fit on the handle's stream               */
    int64_t lane_passes;     /* sum over passes of lanes processed (for roofline arithmetic)              */
    double  update_ms;       /* CUDA-event time spent in the L-BFGS update kernel (when timing is requested) */
    int64_t compactions;     /* lane compactions performed                                                */
} tvs_fit_stats;

/* library / device ------------------------------------------------------------------------- */
int         tvs_version(void);
const char* tvs_last_error(void);
int         tvs_device_count(int* n_out);

/* problem = the shared CSR (replaces the symengine expression build of
 * ModelFitMaxLike.get_neg_likelihood_of_var_freq, ModelFitMaxLike.py:515-558, and
 * PathMultinomialModel.get_like_formula's bins x variants loop, ModelGenerator.py:153-169).
 * row_ptr[R+1], col[nnz], val[nnz] int32 with 0 <= col < V sorted within a row, val >= 1; L[V] int64 > 0. */
int tvs_create(tvs_handle* out, int device, int64_t R, int64_t V, int64_t nnz,
               const int32_t* row_ptr, const int32_t* col, const int32_t* val, const int64_t* L);
int tvs_destroy(tvs_handle h);

/* lanes (replaces the per-replicate / per-subset rebuild of the objective: traversome.py:515-545,
 * ModelFitMaxLike.py:349-370).  w[R*B] int32 >= 0; mask[V*B] uint8 or NULL (= all variants);
 * C[B] or NULL (= 0).  Computes N_b = sum_r w[r,b] on the device. */
int tvs_set_lanes(tvs_handle h, int64_t B, const int32_t* w, const uint8_t* mask, const double* C);

/* same with 16-bit weights (counts <= 65535, e.g. bootstrap replicates): halves the host->device traffic of a batch.
 * The weights STAY 16-bit on the device (the likelihood pass reads them in that form; whenever every count of a batch
 * fits, the other tvs_set_lanes* / tvs_resample entry points make the same 16-bit copy on the device).  This call does
 * not synchronise: the copy is queued on the handle's stream and tvs_fit / tvs_loglik follow on it, so `w` (when it is
 * pinned host memory) must stay unchanged until the next call on the handle has returned. */
int tvs_set_lanes_u16(tvs_handle h, int64_t B, const uint16_t* w, const uint8_t* mask, const double* C);

/* same, for batches in which many lanes share one weight column (all candidate subsets of one
 * replicate in a backward-selection step, ModelFitMaxLike.py:206-262): w_cols[R*n_cols] int32
 * (element (r, c) at [r*n_cols + c]) holds the DISTINCT weight columns and lane b uses column
 * col_of_lane[b]; the R x B lane-minor matrix is expanded on the device, so only R*n_cols integers
 * cross PCIe. */
int tvs_set_lanes_indexed(tvs_handle h, int64_t B, int64_t n_cols, const int32_t* w_cols, const int32_t* col_of_lane,
                          const uint8_t* mask, const double* C);

/* Device-resident weight columns.  The distinct weight vectors of a bootstrap / model selection (one per replicate:
 * traversome.py:515-545 rebuilds them, and the objective, for every fit) are registered ONCE in a column store on the
 * device; every later batch names its lanes' columns by slot, so no weights cross PCIe per batch.
 *   tvs_columns_reserve     capacity of the store (columns); growing keeps the registered columns
 *   tvs_columns_upload[_u16] w_cols[R*n] (element (r, c) at [r*n + c], host or device) -> slots first .. first+n-1
 *   tvs_columns_resample    slots first .. first+n-1 <- bootstrap replicates drawn on the device as in tvs_resample; the
 *                           Philox stream of slot first+j is keyed by first_id + j (a column depends on its id only)
 *   tvs_set_lanes_columns   lane b uses slot col_of_lane[b]; mask / C as in tvs_set_lanes */
int tvs_columns_reserve(tvs_handle h, int64_t n_cols);
int tvs_columns_upload(tvs_handle h, int64_t first, int64_t n, const int32_t* w_cols);
int tvs_columns_upload_u16(tvs_handle h, int64_t first, int64_t n, const uint16_t* w_cols);
int tvs_columns_resample(tvs_handle h, int64_t first, int64_t n, const int32_t* n_obs, uint64_t seed, uint64_t first_id);
int tvs_set_lanes_columns(tvs_handle h, int64_t B, const int32_t* col_of_lane, const uint8_t* mask, const double* C);

/* device-side bootstrap: lane b's w[.,b] ~ Multinomial(N, n_obs/N), N = sum n_obs, drawn record by
 * record with Philox4x32-10(key=(seed, b), counter=draw index) -- the batched equivalent of
 * `random.choices(records_pool_sorted, k=N)` (traversome.py:1652) for synthetic workloads.
 * Lane 0 is the unresampled data when keep_lane0 != 0.  mask/C as in tvs_set_lanes. */
int tvs_resample(tvs_handle h, int64_t B, const int32_t* n_obs, uint64_t seed, int keep_lane0,
                 const uint8_t* mask, const double* C);

/* copies the current lane weights (R*B int32, lane-minor) to `w_out` (host or device). */
int tvs_get_weights(tvs_handle h, int32_t* w_out);

/* LL_b(theta[.,b]) for every lane (replaces the lambdified objective, ModelFitMaxLike.py:534-545).
 * theta need not be normalised; entries outside the lane's mask are ignored.  -inf where the
 * reference's objective is -inf. */
int tvs_loglik(tvs_handle h, const double* theta, double* ll_out);

/* maximum-likelihood fit of every lane (replaces minimize_neg_likelihood, ModelFitMaxLike.py:19-58).
 * theta0[V*B] or NULL (= uniform over the lane's mask).  A lane without any record (all weights zero) has a constant
 * likelihood: it is reported TVS_CONVERGED at its starting point with ll = C.  Outputs (any may be NULL):
 * theta_out[V*B] (sums to 1 over the mask, 0 outside), ll_out[B] (includes C), kkt_out[2*B]
 * (complementarity, dual residual), iters_out[B], status_out[B].
 * Any B is served by the 2- and 4-lane kernels: a lane count that is not a multiple of 4 is re-laid once, inside the call,
 * at the next multiple (1..3 inert padding lanes; one extra copy of the lane weights at that width stays allocated on
 * the handle).  TVS_PAD=0 in the environment keeps the caller's width (1-lane kernels for odd B). */
int tvs_fit(tvs_handle h, const double* theta0, const tvs_fit_options* opt,
            double* theta_out, double* ll_out, double* kkt_out, int32_t* iters_out, int32_t* status_out);

int tvs_last_fit_stats(tvs_handle h, tvs_fit_stats* out);

/* sizes of the handle, for callers that only hold the handle */
int tvs_dims(tvs_handle h, int64_t* R, int64_t* V, int64_t* nnz, int64_t* B);

/* fp64 FMA peak micro-benchmark (MEASURED_PEAKS.json has no fp64 entry; SURVEY.md 8d asks for one).
 * Returns TFLOP/s of dependent-chain-free DFMA on `device`. */
int tvs_measure_fp64_peak(int device, double* tflops_out);
/* shared-memory read peak: GB/s of conflict-free 16-byte loads (LDS.128) summed over all SMs -- the roof that bounds the
 * sparse products of the likelihood pass (one 8-byte shared-memory operand per DFMA). */
int tvs_measure_smem_peak(int device, double* gbs_out);

/* self-test of the device reciprocal / logarithm used inside the likelihood pass (they replace the
 * IEEE division and libdevice log there): log_out[i] = log(p[i]), rcp_out[i] = 1/p[i], host arrays. */
int tvs_selftest_math(int device, int64_t n, const double* p, double* log_out, double* rcp_out);
/* the table-driven logarithm of the staged likelihood pass (uses the handle's table): log_out[i] = log(p[i]). */
int tvs_selftest_table_log(tvs_handle h, int64_t n, const double* p, double* log_out);

#ifdef __cplusplus
}
#endif
#endif /* TVS_H_ */
