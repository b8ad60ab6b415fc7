// Internal declarations shared by the translation units of libtvs.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/tvs.h"

namespace tvs {

constexpr int kWarps = 8;            // warps per CTA of the pass kernel
constexpr int kThreads = kWarps * 32;
constexpr int kMaxHistory = 16;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#define TVS_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return tvs::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// Read paths are processed in "chunks" of consecutive rows.  Inside a chunk the non-zeros are
// stored twice: row-major (forward product p = A x) and column-major (transposed product
// t = A^T (w/p)), so that both products read their dense operand conflict-free from shared memory.
struct Csr {
    int64_t R = 0, V = 0, nnz = 0;
    int n_chunks = 0;
    int rows_per_chunk = 0;          // upper bound
    // device arrays
    int32_t*  row_ptr = nullptr;     // [R+1]
    uint32_t* ent = nullptr;         // [nnz]  col | val<<16, row-major
    int32_t*  chunk_row0 = nullptr;  // [n_chunks+1]
    int32_t*  ccol_ptr = nullptr;    // [n_chunks+1] -> index into ccol_id / ccol_start
    int32_t*  ccol_id = nullptr;     // [n_ccol] column id of each non-empty (chunk, column)
    int32_t*  ccol_start = nullptr;  // [n_ccol+1] -> index into cent
    uint32_t* cent = nullptr;        // [nnz]  local_row | val<<16, column-major within chunk
    uint32_t* ent3 = nullptr;        // [nnz]  col | (top 16 bits of the fp64 count) << 16: pass3_kernel's form (no conversion per entry)
    uint32_t* cent3 = nullptr;       // [nnz]  the same for the column-major entries
    bool      ent3_ok = false;       // every count is exact in 16 bits of an fp64 (1..32, even numbers to 64, ...)
    int       max_chunk_nnz = 0;     // largest number of entries in one chunk
    void*     desc = nullptr;        // [n_chunks] ChunkDesc (tvs_kernels.cuh)
    void*     logtab = nullptr;      // [128] double2 {c_i, -log c_i} (table_log_pos)
    // Newton phase (V <= 128): products A_ru A_rv (u >= v) grouped by H entry inside blocks of h_HB rows (tvs_newton.cuh)
    int64_t*  slab_ptr = nullptr;    // [h_nblk * (h_NG + 1)] step offsets
    uint16_t* h_row = nullptr;       // [h_steps * 32] block-local row (h_HB = padding)
    float*    h_val = nullptr;       // [h_steps * 32] product (exact: < 2^24 checked)
    uint16_t* h_slot = nullptr;      // [h_nblk * h_NG * 32] packed H index u(u+1)/2+v, 0xFFFF = none
    int       h_HB = 0, h_nblk = 0, h_NG = 0;
    int32_t*  wide_ptr = nullptr;    // [w_nblk * (NP + 1)] entry-major lists inside blocks of w_HB rows (hess_wide_kernel)
    void*     wide_rec = nullptr;    // uint2 {row * 128, fp32 bits of the product}
    int       w_HB = 0, w_nblk = 0, w_cap = 0;
    int64_t   h_steps = 0;
    double*   dense = nullptr;       // [R*V] row-major fp64 copy of A for the DGEMM pass (dense shapes only, see tvs_create)
    // blocked copies of A (major = read path) and A^T (major = variant) for blk_kernel (wide candidate sets, tvs_blk.cuh)
    struct Blk { uint32_t* seg_off = nullptr; void* blob = nullptr; int n_chunks = 0, n_panels = 0, max_seg_bytes = 0, dealt = 0; };
    Blk blk_f, blk_b;
    bool has_blk = false;
    // Newton phase (V <= 128): the Hessian as ONE more blocked product -- G[(u,v), r] = A_ru A_rv (u >= v, packed index
    // u(u+1)/2+v) times the dense w/p^2 [R x lanes] -- run by blk_kernel's backward mode
    Blk blk_h;
    bool has_blk_h = false;
    double*   invL = nullptr;        // [V] 1/L_v
    double*   Ld = nullptr;          // [V] L_v as double
    int64_t   n_ccol = 0;
};

struct Lanes {
    int64_t B = 0;
    int32_t* w = nullptr;      // [R*B]
    uint8_t* mask = nullptr;   // [V*B]
    double*  C = nullptr;      // [B]
    double*  N = nullptr;      // [B]
    int64_t  cap_w = 0, cap_mask = 0, cap_b = 0;
    uint16_t* w16 = nullptr;   // [R*B] the same weights as 16-bit counts (valid when has16)
    int64_t  cap_w16 = 0;
    bool     has32 = false;    // w holds the current lanes
    bool     has16 = false;    // w16 holds the current lanes (every count <= 65535); the pass kernels prefer it
    int32_t* flag = nullptr;   // device word: overflow flag of the narrowing kernel
    // device-resident column store (tvs_columns_*): [R x col_cap] int32, element (r, c) at [r*col_cap + c]
    int32_t* cols = nullptr; int64_t col_cap = 0;
    uint16_t* cols16 = nullptr; bool cols16_ok = true;   // 16-bit mirror of the store (valid while no registered count exceeded 65,535)
    int32_t* slot = nullptr; int64_t cap_slot = 0;   // staging of the lanes' column slots
    unsigned long long* acc = nullptr;   // integer lane sums (finish_lanes)
    int64_t  cap_acc = 0;
};

// One set of per-lane arrays at some width.  tvs_fit keeps two of them and "compacts" the lanes
// that are still iterating from one into the other whenever at least 30 % of the lanes have finished,
// so that the pass kernel never spends time on converged lanes.
struct LaneSet {
    int64_t width = 0;
    // inputs (set 0 views the arrays of Lanes; after the first compaction they are owned copies)
    const int32_t* w = nullptr; const uint16_t* w16 = nullptr;   // weights: the 16-bit form when the lanes carry it, else int32
    const uint8_t* mask = nullptr; const double* N = nullptr; const double* C = nullptr;
    int32_t* orig = nullptr;      // [cap] original lane id of each column
    // solver state, all lane-minor [rows * width]
    double *z = nullptr, *zt = nullptr, *g = nullptr, *d = nullptr, *x = nullptr;          // V rows
    double *S = nullptr, *Y = nullptr;                                                     // m*V rows
    double *rho = nullptr;                                                                 // m rows
    double *gamma = nullptr, *F = nullptr, *dg = nullptr, *alpha = nullptr, *llsum = nullptr, *sumphi = nullptr;
    double *kkt = nullptr;                                                                 // 2 rows
    double *gram = nullptr;                                                                // (2m+1)(2m+2)/2 rows (vector-free L-BFGS)
    int32_t *ls = nullptr, *iters = nullptr, *head = nullptr, *cnt = nullptr, *done = nullptr, *status = nullptr;
};

struct FitState {               // all device, sized for (V, B, history)
    int64_t V = 0, B = 0, R = 0; int m = 0;
    LaneSet set[2];
    int32_t* wbuf[2] = {nullptr, nullptr}; uint16_t* wbuf16[2] = {nullptr, nullptr};   // compaction buffers of the weights (allocated on first use)
    uint8_t* mbuf[2] = {nullptr, nullptr};
    double* nbuf[2] = {nullptr, nullptr}; double* cbuf[2] = {nullptr, nullptr};
    int64_t half = 0;           // capacity (lanes) of the compaction buffers above
    // inputs of a batch whose lane count is not a multiple of 4, re-laid at the padded width (tvs_fit, set 0)
    void* pad_w = nullptr; size_t pad_w_cap = 0; uint8_t* pad_mask = nullptr; size_t pad_mask_cap = 0;
    double *pad_N = nullptr, *pad_C = nullptr; size_t pad_b_cap = 0;
    double *gt = nullptr, *gh = nullptr;            // scratch, V rows at full width
    double *tpart = nullptr, *llpart = nullptr; size_t tpart_cap = 0, llpart_cap = 0;
    double *t_red = nullptr, *ll_red = nullptr;     // partials summed over the row splits
    int32_t *idx = nullptr;                          // compaction gather list
    // Newton phase scratch
    double *gha = nullptr, *lam = nullptr, *Hpart = nullptr, *cr = nullptr; float* crf = nullptr; int32_t* need = nullptr; size_t Hpart_cap = 0, cr_cap = 0, crf_cap = 0;
    int32_t *nactive = nullptr;   // device counters (ring)
    int32_t *h_nactive = nullptr; // pinned host mirror
    // output staging at the ORIGINAL width
    double *theta_io = nullptr, *ll_out = nullptr, *kkt_out = nullptr; int32_t *iters_out = nullptr, *status_out = nullptr;
    std::vector<void*> owned;   // every cudaMalloc'ed block above
};

}  // namespace tvs

struct tvs_problem {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr;
    tvs::Csr csr;
    tvs::Lanes lanes;
    tvs::FitState fs;
    tvs_fit_stats stats{};
    bool lanes_set = false;
    bool attrs_set = false;     // cudaFuncSetAttribute of the update kernels done
    // dense (DGEMM) pass: cuBLAS handle bound to `stream`, p / w/p matrix [R x width], per-split log-likelihood partials
    void* cublas = nullptr;
    double* dense_p = nullptr; size_t dense_p_cap = 0;
    double* dense_ll = nullptr; size_t dense_ll_cap = 0;
    // blk pass: w/p of the forward product [R x width], read back by the backward product
    double* blk_rr = nullptr; size_t blk_rr_cap = 0;
    double* ll_scratch = nullptr; size_t ll_scratch_cap = 0;     // tvs_loglik: theta, x, denominators, partials, result
};
