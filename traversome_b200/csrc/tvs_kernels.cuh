// CUDA kernels of libtvs.so (sm_100a).  Included once by tvs_api.cu.
//
// Layout rule for every per-lane matrix: lane-minor, element (i, b) at [i*B + b]; a warp's 32
// threads are 32 consecutive lanes, so every global access below is a contiguous 128/256-byte
// segment and every shared-memory access is a conflict-free 64-bit column.
#pragma once
#include "tvs_internal.h"
#include <math.h>

namespace tvs {

// ------------------------------------------------------------------------------------------------
// Fused likelihood pass:  p = A x ;  rr = w / p ;  ll += w log p ;  t = A^T rr      (per lane)
//
// replaces: PathMultinomialModel.get_like_formula (ModelGenerator.py:153-174) evaluated by the
// lambdified objective (ModelFitMaxLike.py:534-545), once per SLSQP function evaluation.
//
// grid = (lane tiles, row splits); CTA (tile, split) walks chunks split, split+S, ...
//   phase 1 (row-major entries):   warp <- rows of the chunk, thread <- LPT lanes of the tile
//   phase 2 (column-major entries): warp <- non-empty columns of the chunk, same lanes
// XS: x tile and t accumulators live in shared memory (V*LT*16 bytes); otherwise x is read from
// global/L2 and t is accumulated in the CTA's private slice of the partial buffer.
// ------------------------------------------------------------------------------------------------
struct LogCoef { double c[11]; };                  // see tvs_pass2.cuh
struct __align__(16) ChunkDesc { int row0, nrows, e0, nent, c0, ncol, ce0, ncent; };   // one per chunk of rows

struct PassArgs {
    const int32_t* row_ptr; const uint32_t* ent;
    const int32_t* chunk_row0; const int32_t* ccol_ptr; const int32_t* ccol_id; const int32_t* ccol_start;
    const uint32_t* cent;
    const int32_t* w; const double* x; const int32_t* done;
    const uint16_t* w16;  // the same weights as 16-bit counts (preferred by pass3_kernel / blk_kernel when not null: half the traffic)
    double* tpart; double* llpart;
    int64_t B; int V; int n_chunks; int rows_per_chunk;
    int stage_cap;       // capacity (entries) of the shared-memory staging buffers of pass3_kernel
    const ChunkDesc* desc;
    LogCoef logc;
    double* cr;          // optional [R*B]: w / p^2 per (row, lane) for the Newton phase's Hessian (pass3_kernel only)
    float* crf;          // the same in fp32 (wide lane sets: hess_wide_kernel)
    const double2* logtab;   // [128] {c_i, -log c_i} for table_log_pos, or null (atanh series)
    const uint32_t* ent3; const uint32_t* cent3;   // entries with the count as the top 16 bits of its fp64 form (pass3_kernel)
};

template <int LPT, bool XS, bool BACKWARD>
__global__ void __launch_bounds__(kThreads) pass_kernel(const PassArgs a) {
    constexpr int LT = 32 * LPT;
    extern __shared__ double smem[];
    const int t = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int tile = blockIdx.x, split = blockIdx.y, S = gridDim.y;
    const int64_t B = a.B;
    const int V = a.V;
    const int64_t lane0 = (int64_t)tile * LT;

    // lanes of this thread
    int64_t lane[LPT]; bool lv[LPT];
#pragma unroll
    for (int j = 0; j < LPT; ++j) { lane[j] = lane0 + j * 32 + t; lv[j] = lane[j] < B; }

    // skip tiles whose lanes have all finished (uniform over the CTA)
    if (a.done != nullptr) {
        int alive = 0;
#pragma unroll
        for (int j = 0; j < LPT; ++j) alive |= (lv[j] && a.done[lane[j]] == 0);
        if (!__syncthreads_or(alive)) return;
    }

    double* x_s = smem;                                   // [V*LT]        (XS)
    double* t_s = x_s + (XS ? (size_t)V * LT : 0);        // [V*LT]        (XS && BACKWARD)
    double* rr_s = t_s + ((XS && BACKWARD) ? (size_t)V * LT : 0);  // [rows_per_chunk*LT]

    if (XS) {
        for (int i = threadIdx.x; i < V * LT; i += kThreads) {
            const int v = i / LT, l = i - v * LT;
            const int64_t b = lane0 + l;
            x_s[i] = (b < B) ? a.x[(int64_t)v * B + b] : 0.0;
            if (BACKWARD) t_s[i] = 0.0;
        }
        __syncthreads();
    }

    double ll[LPT];
#pragma unroll
    for (int j = 0; j < LPT; ++j) ll[j] = 0.0;

    for (int c = split; c < a.n_chunks; c += S) {
        const int row0 = a.chunk_row0[c], row1 = a.chunk_row0[c + 1];
        // ---------------- phase 1: p = A x over the rows of the chunk
        for (int r = row0 + wi; r < row1; r += kWarps) {
            const int k0 = __ldg(a.row_ptr + r), k1 = __ldg(a.row_ptr + r + 1);
            double p[LPT];
#pragma unroll
            for (int j = 0; j < LPT; ++j) p[j] = 0.0;
            for (int k = k0; k < k1; ++k) {
                const uint32_t e = __ldg(a.ent + k);
                const int col = e & 0xffffu;
                const double f = (double)(e >> 16);
#pragma unroll
                for (int j = 0; j < LPT; ++j) {
                    double xv;
                    if (XS) xv = x_s[col * LT + j * 32 + t];
                    else xv = lv[j] ? __ldg(a.x + (int64_t)col * B + lane[j]) : 0.0;
                    p[j] = fma(f, xv, p[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < LPT; ++j) {
                const int wv = !lv[j] ? 0 : (a.w16 != nullptr ? (int)__ldg(a.w16 + (int64_t)r * B + lane[j]) : __ldg(a.w + (int64_t)r * B + lane[j]));
                double rr = 0.0;
                if (wv > 0) {
                    const double wd = (double)wv;
                    rr = wd / p[j];
                    ll[j] = fma(wd, log(p[j]), ll[j]);
                }
                if (BACKWARD) rr_s[(r - row0) * LT + j * 32 + t] = rr;
            }
        }
        if (BACKWARD) {
            __syncthreads();
            // ------------ phase 2: t += A^T rr over the non-empty columns of the chunk
            const int c0 = a.ccol_ptr[c], c1 = a.ccol_ptr[c + 1];
            for (int ci = c0 + wi; ci < c1; ci += kWarps) {
                const int v = __ldg(a.ccol_id + ci);
                const int k0 = __ldg(a.ccol_start + ci), k1 = __ldg(a.ccol_start + ci + 1);
                double acc[LPT];
#pragma unroll
                for (int j = 0; j < LPT; ++j) acc[j] = 0.0;
                for (int k = k0; k < k1; ++k) {
                    const uint32_t e = __ldg(a.cent + k);
                    const int rl = e & 0xffffu;
                    const double f = (double)(e >> 16);
#pragma unroll
                    for (int j = 0; j < LPT; ++j) acc[j] = fma(f, rr_s[rl * LT + j * 32 + t], acc[j]);
                }
#pragma unroll
                for (int j = 0; j < LPT; ++j) {
                    if (XS) t_s[v * LT + j * 32 + t] += acc[j];
                    else if (lv[j]) a.tpart[((int64_t)split * V + v) * B + lane[j]] += acc[j];
                }
            }
            __syncthreads();
        }
    }

    // ---------------- epilogue: per-split partials, summed later in fixed split order (deterministic)
    if (XS && BACKWARD) {
        for (int i = threadIdx.x; i < V * LT; i += kThreads) {
            const int v = i / LT, l = i - v * LT;
            const int64_t b = lane0 + l;
            if (b < B) a.tpart[((int64_t)split * V + v) * B + b] = t_s[i];
        }
    }
    __syncthreads();
    double* red = rr_s;  // reuse: [kWarps*LT]; rows_per_chunk >= kWarps is enforced on the host
#pragma unroll
    for (int j = 0; j < LPT; ++j) red[wi * LT + j * 32 + t] = ll[j];
    __syncthreads();
    if (wi == 0) {
#pragma unroll
        for (int j = 0; j < LPT; ++j) {
            double s = 0.0;
            for (int q = 0; q < kWarps; ++q) s += red[q * LT + j * 32 + t];
            if (lv[j]) a.llpart[(int64_t)split * B + lane[j]] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Sum of the per-split partials in FIXED split order (bitwise reproducible), one thread per (row, lane):
// t[v,b] = sum_q tpart[q,v,b] (rows 0..V-1) and ll[b] = sum_q llpart[q,b] (row V).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ tpart, const double* __restrict__ llpart,
                                                              int S_t, int S_ll, int V, int64_t B, double* __restrict__ t_red,
                                                              double* __restrict__ ll_red) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)(V + 1) * B) return;
    const bool is_ll = e >= (int64_t)V * B;
    const int S = is_ll ? S_ll : S_t;
    const double* src = is_ll ? llpart + (e - (int64_t)V * B) : tpart + e;
    const int64_t stride = is_ll ? B : (int64_t)V * B;
    double s0 = 0.0;
    int q = 0;
    for (; q + 8 <= S; q += 8) {             // 8 independent loads in flight, summed in order
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = src[(int64_t)(q + i) * stride];
#pragma unroll
        for (int i = 0; i < 8; ++i) s0 += v[i];
    }
    for (; q < S; ++q) s0 += src[(int64_t)q * stride];
    if (is_ll) ll_red[e - (int64_t)V * B] = s0; else t_red[e] = s0;
}

// ------------------------------------------------------------------------------------------------
// small per-lane kernels
// ------------------------------------------------------------------------------------------------
// N_b = sum_r w[r,b]: grid.x = lane tiles of 32, grid.y = row slices; exact integer partial sums, one atomicAdd
// (64-bit integer: order-independent) per (slice, lane).  lane_sum_finish_kernel converts to double.
template <typename T>
__global__ void __launch_bounds__(256) lane_sum_kernel(const T* __restrict__ w, int64_t R, int64_t B, unsigned long long* __restrict__ acc) {
    __shared__ long long red[8][32];
    const int t = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * 32 + t;
    const int64_t rows = (R + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * rows, r1 = min(R, r0 + rows);
    long long s = 0;
    if (b < B)
        for (int64_t r = r0 + wi; r < r1; r += 8) s += (long long)w[r * B + b];
    red[wi][t] = s;
    __syncthreads();
    if (wi == 0 && b < B) {
        long long tot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) tot += red[q][t];
        atomicAdd(acc + b, (unsigned long long)tot);
    }
}
__global__ void lane_sum_finish_kernel(const unsigned long long* __restrict__ acc, int64_t B, double* __restrict__ N) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) N[b] = (double)(long long)acc[b];
}

// z0 from theta0 (or uniform over the mask): phi_v ~ theta_v * L_v, z = sqrt(phi)
__global__ void init_z_kernel(const double* __restrict__ theta0, const uint8_t* __restrict__ mask,
                              const double* __restrict__ Ld, const double* __restrict__ invL, int V, int64_t B, int64_t B0,
                              double* __restrict__ zt, double* __restrict__ x) {
    // B = width of the lane set (stride of mask / zt / x), B0 <= B = lanes of the caller's theta0 (its stride); lanes
    // b >= B0 are the padding of a set whose width was rounded up
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (b >= B0) theta0 = nullptr;
    double s = 0.0;
    for (int v = 0; v < V; ++v) {
        const bool m = mask ? mask[(int64_t)v * B + b] != 0 : true;
        double ph = 0.0;
        if (m) ph = theta0 ? fmax(theta0[(int64_t)v * B0 + b], 0.0) * Ld[v] : 1.0;
        s += ph;
    }
    for (int v = 0; v < V; ++v) {
        const bool m = mask ? mask[(int64_t)v * B + b] != 0 : true;
        double ph = 0.0;
        if (m) {
            ph = theta0 ? fmax(theta0[(int64_t)v * B0 + b], 0.0) * Ld[v] : 1.0;
            ph = (s > 0.0) ? ph / s : 0.0;
            ph = fmax(ph, 1e-12);      // an exact zero inside the subset is a stationary point of the z-parametrisation
        }
        const double zz = sqrt(ph);
        zt[(int64_t)v * B + b] = zz;
        x[(int64_t)v * B + b] = zz * zz * invL[v];
    }
}

// x for tvs_loglik: x_v = mask ? theta_v : 0 ; denom_b = sum_v x_v L_v
__global__ void prep_theta_kernel(const double* __restrict__ theta, const uint8_t* __restrict__ mask,
                                  const double* __restrict__ Ld, int V, int64_t B, double* __restrict__ x,
                                  double* __restrict__ denom) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s = 0.0;
    for (int v = 0; v < V; ++v) {
        const bool m = mask ? mask[(int64_t)v * B + b] != 0 : true;
        const double th = m ? theta[(int64_t)v * B + b] : 0.0;
        x[(int64_t)v * B + b] = th;
        s = fma(th, Ld[v], s);
    }
    denom[b] = s;
}

__global__ void finish_loglik_kernel(const double* __restrict__ llpart, int S, int64_t B, const double* __restrict__ N,
                                     const double* __restrict__ C, const double* __restrict__ denom, double* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s = 0.0;
    for (int q = 0; q < S; ++q) s += llpart[(int64_t)q * B + b];
    out[b] = s - N[b] * log(denom[b]) + (C ? C[b] : 0.0);
}

// ------------------------------------------------------------------------------------------------
// Batched L-BFGS step in z = sqrt(phi).
//
// minimise  F(z) = -(1/N) sum_r w_r log( sum_v A_rv z_v^2 / L_v ) + sum_v z_v^2      (no constraints)
// whose minimisers are exactly the KKT points of the reference's constrained problem
// (max LL s.t. theta >= 0, sum theta = 1: ModelFitMaxLike.py:27-40) with phi_v = z_v^2 = theta_v L_v / sum(theta L).
// replaces: scipy SLSQP multistart (ModelFitMaxLike.py:19-58).
// ------------------------------------------------------------------------------------------------
struct UpdArgs {
    const double* tpart; const double* llpart; int S; int S_ll;   // split counts of the t / ll partials
    const double* invL; const double* N; const uint8_t* mask;
    double *z, *zt, *g, *gt, *gh, *d, *x, *Sh, *Yh, *rho, *gamma;
    double *F, *dg, *alpha, *llsum, *kkt, *sumphi;
    int32_t *ls, *iters, *status, *head, *cnt, *done, *nactive;
    int V; int64_t B; int m; int max_iter; double tol, tol_dual;
    int32_t* nreset;     // counter zeroed by this launch for a LATER launch of the same poll group (or null)
};

// ------------------------------------------------------------------------------------------------
// Update kernel, "vector-free" form (Chen et al. 2014): the two-loop recursion runs on the
// (2M+1)x(2M+1) Gram matrix of the basis {s_k, y_k, g} held per lane (packed symmetric, lane-minor),
// so that one update needs only two sweeps over V (dot products of the new vectors with the basis,
// then d = sum_k delta_k b_k) instead of the 2M dependent sweeps of the textbook two-loop recursion.
// All dot products of the NEW vectors are computed directly (not derived algebraically) to keep the
// Gram entries accurate to fp64 rounding.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tri_idx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// sum_q src[q*stride], q = 0..S-1, added in order.  Loads are issued 8 at a time with the out-of-range ones
// predicated off (adding +0.0 leaves the sum unchanged): no serial tail loop.
__device__ __forceinline__ double sum_splits(const double* __restrict__ src, int S, int64_t stride) {
    double s0 = 0.0;
    for (int q = 0; q < S; q += 8) {
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (q + i < S) ? src[(int64_t)(q + i) * stride] : 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s0 += v[i];
    }
    return s0;
}

constexpr int kU2Lanes = 8;                                // lanes per CTA (4: 41.0 us per full-width cfg2 update against 38.9)
constexpr int kU2Slots = 32;
constexpr int kU2Threads = kU2Lanes * kU2Slots;
constexpr int kU2Warps = kU2Threads / 32;

template <int M>
struct Vl2Layout {
    static constexpr int NB = 2 * M + 1;
    static constexpr int NG = NB * (NB + 1) / 2;
    static constexpr int NACC = 6 * M + 9;
    // Gs[NG][8] dl[NB][8] acc[NACC+4][8] part[W][NACC+4][8]
    static constexpr size_t doubles = (size_t)(NG + NB + (NACC + 4) + (size_t)kU2Warps * (NACC + 4)) * kU2Lanes + 8 * kU2Lanes;
};

// sums (or maxima for q >= first_max) of vals[0..n) over the 64 slots of each lane -> acc[q][l]
template <int N>
__device__ __forceinline__ void cta_reduce(double* vals, int first_max, double* part, double (*acc)[kU2Lanes], int w, int t, int l) {
#pragma unroll
    for (int q = 0; q < N; ++q) {
        double v = vals[q];
#pragma unroll
        for (int off = kU2Lanes; off < 32; off <<= 1) {          // over the slots of this warp (thread = slot * lanes + lane)
            const double o = __shfl_xor_sync(0xffffffffu, v, off);
            v = (q >= first_max) ? fmax(v, o) : v + o;
        }
        if (t < kU2Lanes) part[((size_t)w * N + q) * kU2Lanes + l] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * kU2Lanes; i += kU2Threads) {
        const int q = i / kU2Lanes, ll = i - q * kU2Lanes;
        double r = part[(size_t)q * kU2Lanes + ll];
        for (int ww = 1; ww < kU2Warps; ++ww) {
            const double o = part[((size_t)ww * N + q) * kU2Lanes + ll];
            r = (q >= first_max) ? fmax(r, o) : r + o;
        }
        acc[q][ll] = r;
    }
    __syncthreads();
}

template <int M>
__global__ void __launch_bounds__(kU2Threads) lbfgs_update_vl2_kernel(const UpdArgs a, double* __restrict__ Gm) {
    constexpr int NB = Vl2Layout<M>::NB, NG = Vl2Layout<M>::NG, NACC = Vl2Layout<M>::NACC;
    constexpr int GI = 2 * M;
    constexpr int NR = NACC + 2;                                  // + comp, dual (max-reduced)
    extern __shared__ double sm2[];
    double (*Gs)[kU2Lanes] = reinterpret_cast<double (*)[kU2Lanes]>(sm2);                                   // [NG]
    double (*dl)[kU2Lanes] = reinterpret_cast<double (*)[kU2Lanes]>(sm2 + (size_t)NG * kU2Lanes);           // [NB]
    double (*acc)[kU2Lanes] = reinterpret_cast<double (*)[kU2Lanes]>(sm2 + (size_t)(NG + NB) * kU2Lanes);   // [NACC+4]
    double* part = sm2 + (size_t)(NG + NB + NACC + 4) * kU2Lanes;                                           // [W][NACC+4][8]
    double (*misc)[kU2Lanes] = reinterpret_cast<double (*)[kU2Lanes]>(part + (size_t)kU2Warps * (NACC + 4) * kU2Lanes);   // [4]

    const int tid = threadIdx.x, t = tid & 31, w = tid >> 5;
    const int l = tid & (kU2Lanes - 1), vs = tid / kU2Lanes;
    const int64_t B = a.B;
    const int64_t b = (int64_t)blockIdx.x * kU2Lanes + l;
    const bool valid = b < B;
    const int64_t bb = valid ? b : B - 1;
    const int V = a.V;
    const bool active = valid && a.done[bb] == 0;
    if (a.nreset != nullptr && blockIdx.x == 0 && tid == 0) *a.nreset = 0;
    if (!__syncthreads_or(active)) return;

    // the Gram matrix is first needed in phase D: fetch it asynchronously so that the loads of phases A and B are not
    // queued behind it (a plain load + shared store stalls the warp until the data arrives)
    for (int q = vs; q < NG; q += kU2Slots) {
        const unsigned dsts = (unsigned)__cvta_generic_to_shared(&Gs[q][l]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dsts), "l"(Gm + (int64_t)q * B + bb));
    }
    asm volatile("cp.async.commit_group;");
    if (active) {
        // the likelihood pass has just streamed the weight matrix through L2, so the solver state is in DRAM: start
        // fetching everything the later sweeps of this thread will read (profiles/r1_final_summary.md: 12 warps per
        // issue stalled on long_scoreboard, 21 MB of DRAM reads per launch)
        for (int v = vs; v < V; v += kU2Slots) {
            const int64_t i = (int64_t)v * B + bb;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.z + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.g + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.d + i));
#pragma unroll
            for (int k = 0; k < M; ++k) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Sh + ((int64_t)k * V + v) * B + bb));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Yh + ((int64_t)k * V + v) * B + bb));
            }
        }
    }

    const double N = a.N[bb];
    // ---- A: sum the per-split partials of the pass evaluated at zt in FIXED split order (bitwise reproducible)
    const double llsum = sum_splits(a.llpart + bb, a.S_ll, B);
    double r2[2] = {0.0, 0.0};                                  // sum phi, d.gt
    for (int v = vs; v < V; v += kU2Slots) {
        const int64_t i = (int64_t)v * B + bb;
        const double tv = sum_splits(a.tpart + i, a.S, (int64_t)V * B);
        const double gh = tv * a.invL[v] / N;
        const double ztv = a.zt[i];
        const bool mk = a.mask ? a.mask[i] != 0 : true;
        const double gtv = mk ? 2.0 * ztv * (1.0 - gh) : 0.0;
        if (active) { a.gt[i] = gtv; a.gh[i] = gh; }
        r2[0] = fma(ztv, ztv, r2[0]);
        r2[1] = fma(a.d[i], gtv, r2[1]);
    }
    cta_reduce<2>(r2, 2, part, acc, w, t, l);
    const double sumphi = acc[0][l], dgt = acc[1][l];
    double Ft = -llsum / N + sumphi;
    const bool fin = isfinite(llsum) && isfinite(sumphi) && sumphi > 0.0;
    if (!fin) Ft = INFINITY;

    const int it = a.iters[bb];
    const bool first = it < 0;
    const double F = a.F[bb], dg = a.dg[bb], alpha = a.alpha[bb];
    bool accept;
    if (first) accept = fin;
    else accept = fin && (Ft <= F + 1e-4 * alpha * dg ||
                          (fabs(Ft - F) <= 1e-14 * fabs(F) && fabs(dgt) <= 0.9 * fabs(dg)));
    accept = accept && active;
    bool finished = false;
    int status = TVS_CONVERGED;
    if (active && first && !fin) { finished = true; status = TVS_INFEASIBLE; }

    // ---- B: one sweep: dot products of the new vectors (s, y, g_new) with the whole basis + KKT residuals
    __syncthreads();                                            // acc is rewritten below
    if (__syncthreads_or(accept)) {
        double ac[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) ac[q] = 0.0;
        ac[NACC + 1] = -INFINITY;
        if (accept) {
            for (int v = vs; v < V; v += kU2Slots) {
                const int64_t i = (int64_t)v * B + bb;
                const double ztv = a.zt[i], gn = a.gt[i], go = a.g[i];
                const double sv = ztv - a.z[i], yv = gn - go;
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const double Sk = a.Sh[((int64_t)k * V + v) * B + bb], Yk = a.Yh[((int64_t)k * V + v) * B + bb];
                    ac[k] = fma(sv, Sk, ac[k]);                 ac[M + k] = fma(sv, Yk, ac[M + k]);
                    ac[2 * M + k] = fma(yv, Sk, ac[2 * M + k]); ac[3 * M + k] = fma(yv, Yk, ac[3 * M + k]);
                    ac[4 * M + k] = fma(gn, Sk, ac[4 * M + k]); ac[5 * M + k] = fma(gn, Yk, ac[5 * M + k]);
                }
                ac[6 * M + 0] = fma(sv, sv, ac[6 * M + 0]); ac[6 * M + 1] = fma(sv, yv, ac[6 * M + 1]);
                ac[6 * M + 2] = fma(yv, yv, ac[6 * M + 2]); ac[6 * M + 3] = fma(gn, gn, ac[6 * M + 3]);
                ac[6 * M + 4] = fma(sv, gn, ac[6 * M + 4]); ac[6 * M + 5] = fma(yv, gn, ac[6 * M + 5]);
                ac[6 * M + 6] = fma(sv, go, ac[6 * M + 6]); ac[6 * M + 7] = fma(yv, go, ac[6 * M + 7]);
                ac[6 * M + 8] = fma(gn, go, ac[6 * M + 8]);
                const bool mk = a.mask ? a.mask[i] != 0 : true;
                if (mk) {
                    const double ghn = a.gh[i] * sumphi;
                    ac[NACC] = fmax(ac[NACC], ztv * ztv / sumphi * fabs(ghn - 1.0));
                    ac[NACC + 1] = fmax(ac[NACC + 1], ghn - 1.0);
                }
            }
        }
        cta_reduce<NR>(ac, NACC, part, acc, w, t, l);
    }
    const double ss = acc[6 * M + 0][l], sy = acc[6 * M + 1][l], yy = acc[6 * M + 2][l], gg = acc[6 * M + 3][l];
    const double comp = acc[NACC][l], dual = acc[NACC + 1][l];

    int head = a.head[bb], cnt = a.cnt[bb];
    const bool store = accept && !first && sy > 1e-10 * sqrt(ss * yy) && sy > 0.0;
    // ---- C: move accepted lanes, write the new curvature pair into ring slot `head`
    if (accept) {
        for (int v = vs; v < V; v += kU2Slots) {
            const int64_t i = (int64_t)v * B + bb;
            const double ztv = a.zt[i], gtv = a.gt[i];
            if (store) {
                a.Sh[((int64_t)head * V + v) * B + bb] = ztv - a.z[i];
                a.Yh[((int64_t)head * V + v) * B + bb] = gtv - a.g[i];
            }
            a.z[i] = ztv; a.g[i] = gtv;
        }
    }
    int iters = it;
    if (accept) {
        iters = it + 1;
        if (comp <= a.tol && dual <= a.tol_dual) { finished = true; status = TVS_CONVERGED; }
        else if (iters >= a.max_iter) { finished = true; status = TVS_MAXITER; }
    }
    const bool newdir = accept && !finished;
    double anew = 1.0;

    // ---- D: Gram update and two-loop recursion on the coefficients.  The recursion is a sequence of 2*cnt rank-one
    //      updates of the NB-vector u; thread (lane l, slot vs) owns u[vs] (and u[vs + 32] when NB > 32), the pivot
    //      element of each step travels through shared memory (double-buffered, one barrier per step).  One thread
    //      per lane did all of it before: ~10k dependent instructions at M = 16, a third of the kernel's time.
    //      Every element sees the same operations in the same order, so the results are bitwise unchanged.
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                             // Gram matrix landed
    {
        double (*bc)[kU2Lanes] = misc + 2;                       // [2] pivot broadcast, [1] scalar broadcast
        double* al = part;                                       // [M][8] scratch (the partials are consumed)
        const bool own0 = vs < NB, own1 = vs + 32 < NB;
        if (accept && vs < M) {
            const int k = vs;
            Gs[tri_idx(GI, k)][l] = acc[4 * M + k][l]; Gs[tri_idx(GI, M + k)][l] = acc[5 * M + k][l];
            if (store) {
                const int h = head;
                Gs[tri_idx(h, k)][l] = acc[k][l];             Gs[tri_idx(h, M + k)][l] = acc[M + k][l];
                Gs[tri_idx(M + h, k)][l] = acc[2 * M + k][l]; Gs[tri_idx(M + h, M + k)][l] = acc[3 * M + k][l];
            }
        }
        __syncthreads();
        if (accept && vs == 0) {                                 // entries the loop above wrote provisionally
            Gs[tri_idx(GI, GI)][l] = gg;
            if (store) {
                const int h = head;
                Gs[tri_idx(h, h)][l] = ss; Gs[tri_idx(h, M + h)][l] = sy; Gs[tri_idx(M + h, M + h)][l] = yy;
                Gs[tri_idx(GI, h)][l] = acc[6 * M + 4][l]; Gs[tri_idx(GI, M + h)][l] = acc[6 * M + 5][l];
            }
        }
        if (store) { head = (head + 1) % M; cnt = min(cnt + 1, M); }
        __syncthreads();
        double dgn = -gg;
        double u0 = 0.0, u1 = 0.0;
        if (newdir) {
            if (own0) { dl[vs][l] = 0.0; u0 = -Gs[tri_idx(vs, GI)][l]; }
            if (own1) { dl[vs + 32][l] = 0.0; u1 = -Gs[tri_idx(vs + 32, GI)][l]; }
        }
        __syncthreads();
        if (newdir && vs == 0) dl[GI][l] = -1.0;
        for (int j = 0; j < M; ++j) {                             // newest -> oldest
            const bool on = newdir && j < cnt;
            const int sl = ((head - 1 - j) % M + M) % M;          // sl < M <= 16: owned in slot 0 of thread vs == sl
            if (on && vs == sl) bc[j & 1][l] = u0;
            __syncthreads();
            if (on) {
                const double aj = bc[j & 1][l] / Gs[tri_idx(sl, M + sl)][l];
                if (vs == 0) { al[j * kU2Lanes + l] = aj; dl[M + sl][l] -= aj; }
                if (own0) u0 = fma(-aj, Gs[tri_idx(vs, M + sl)][l], u0);
                if (own1) u1 = fma(-aj, Gs[tri_idx(vs + 32, M + sl)][l], u1);
            }
        }
        __syncthreads();
        if (newdir && cnt > 0) {
            const int sl = ((head - 1) % M + M) % M;
            const double gamma = Gs[tri_idx(sl, M + sl)][l] / Gs[tri_idx(M + sl, M + sl)][l];
            if (own0) { dl[vs][l] *= gamma; u0 *= gamma; }
            if (own1) { dl[vs + 32][l] *= gamma; u1 *= gamma; }
        }
        __syncthreads();
        for (int jj = 0; jj < M; ++jj) {                          // oldest -> newest
            const int j = M - 1 - jj;
            const bool on = newdir && j < cnt;
            const int sl = ((head - 1 - j) % M + M) % M;          // M + sl < 2M <= 32: slot 0 of thread vs == M + sl
            if (on && vs == M + sl) bc[jj & 1][l] = u0;
            __syncthreads();
            if (on) {
                const double cf = al[j * kU2Lanes + l] - bc[jj & 1][l] / Gs[tri_idx(sl, M + sl)][l];
                if (vs == 0) dl[sl][l] += cf;
                if (own0) u0 = fma(cf, Gs[tri_idx(vs, sl)][l], u0);
                if (own1) u1 = fma(cf, Gs[tri_idx(vs + 32, sl)][l], u1);
            }
        }
        if (newdir && ((GI < 32 && vs == GI) || (GI >= 32 && vs == GI - 32))) misc[4][l] = (GI < 32) ? u0 : u1;
        __syncthreads();
        if (newdir) {
            dgn = misc[4][l];
            if (!(dgn < 0.0) || cnt == 0) {                       // not a descent direction / no history
                if (own0) dl[vs][l] = (vs == GI) ? -1.0 : 0.0;
                if (own1) dl[vs + 32][l] = (vs + 32 == GI) ? -1.0 : 0.0;
                if (!(dgn < 0.0)) cnt = 0;
                dgn = -gg;
                anew = fmin(1.0, 1.0 / sqrt(fmax(gg, 1e-300)));
            }
            if (gg == 0.0) { finished = true; status = TVS_CONVERGED; }
        }
        if (vs == 0 && accept) {
            a.F[bb] = Ft; a.llsum[bb] = llsum; a.sumphi[bb] = sumphi;
            a.kkt[bb] = comp; a.kkt[B + bb] = dual; a.iters[bb] = iters; a.ls[bb] = 0;
            a.dg[bb] = dgn; a.alpha[bb] = anew;
            a.head[bb] = head; a.cnt[bb] = cnt;
        }
    }
    __syncthreads();

    // ---- E: d = sum_k delta_k b_k ; next trial point
    if (newdir && !finished) {
        // Saddle escape (see solve_kernel in tvs_newton.cuh): z_v = 0 with g_v > 1 has zero gradient and negative
        // curvature, so the quasi-Newton direction never moves it.  Once the complementarity residual has converged and
        // only the multiplier test fails, such components are placed at a small positive share; F decreases to second
        // order and the line search judges the trial point as usual.
        const bool kick = comp <= a.tol && dual > a.tol_dual;
        for (int v = vs; v < V; v += kU2Slots) {
            const int64_t i = (int64_t)v * B + bb;
            double dv = dl[GI][l] * a.g[i];
#pragma unroll
            for (int k = 0; k < M; ++k) {
                dv = fma(dl[k][l], a.Sh[((int64_t)k * V + v) * B + bb], dv);
                dv = fma(dl[M + k][l], a.Yh[((int64_t)k * V + v) * B + bb], dv);
            }
            const double zo = a.z[i];
            double zn = fma(anew, dv, zo);
            if (kick) {
                const bool mk = a.mask ? a.mask[i] != 0 : true;
                if (mk && a.gh[i] * sumphi - 1.0 > a.tol_dual && zo * zo < 1e-11 * sumphi) {
                    zn = sqrt(1e-8 * sumphi);
                    dv = (zn - zo) / anew;
                }
            }
            a.d[i] = dv;
            a.zt[i] = zn; a.x[i] = zn * zn * a.invL[v];
        }
    }
    if (accept)
        for (int q = vs; q < NG; q += kU2Slots) Gm[(int64_t)q * B + bb] = Gs[q][l];

    // ---- F: rejected lanes backtrack along the stored direction
    const bool reject = active && !accept && !finished;
    if (reject) {
        const int ls = a.ls[bb] + 1;
        double ar = alpha;
        bool stall = false;
        if (ls > 40) stall = true;
        else {
            double aq = 0.5 * alpha;
            if (isfinite(Ft)) {
                const double den = 2.0 * (Ft - F - dg * alpha);
                if (den > 0.0) aq = -dg * alpha * alpha / den;
                aq = fmin(fmax(aq, 0.1 * alpha), 0.5 * alpha);
            } else aq = 0.1 * alpha;
            ar = aq;
        }
        if (!stall) {
            for (int v = vs; v < V; v += kU2Slots) {
                const int64_t i = (int64_t)v * B + bb;
                const double zn = fma(ar, a.d[i], a.z[i]);
                a.zt[i] = zn; a.x[i] = zn * zn * a.invL[v];
            }
        }
        if (vs == 0) { a.alpha[bb] = ar; a.ls[bb] = ls; }
        if (stall) {
            const double c0 = a.kkt[bb], d0 = a.kkt[B + bb];
            finished = true;
            status = (c0 <= 1e2 * a.tol && d0 <= 1e2 * a.tol_dual) ? TVS_CONVERGED : TVS_STALLED;
        }
    }

    if (vs == 0 && active && finished) { a.status[bb] = status; a.done[bb] = 1; }
    if (w == 0) {
        const unsigned still = __ballot_sync(0xffffffffu, t < kU2Lanes && active && !finished);
        if (t == 0 && still) atomicAdd(a.nactive, __popc(still));
    }
}

// theta_out, ll_out, ... from the last accepted point of the lanes of one LaneSet, written at the lanes'
// ORIGINAL ids (orig[]) into arrays of the original width B0.  only_done: skip lanes still iterating.
struct FinArgs {
    const double *z, *invL, *llsum, *sumphi, *N, *C, *kkt; const int32_t *iters, *status, *done, *orig;
    double *theta_out, *ll_out, *kkt_out; int32_t *iters_out, *status_out; int V; int64_t B, B0; int only_done;
};
__global__ void finalize_fit_kernel(const FinArgs a) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const bool dn = a.done[b] != 0;
    if (a.only_done && !dn) return;
    const int st = dn ? a.status[b] : TVS_MAXITER;   // still running when the pass budget ran out
    const int64_t o = a.orig[b];
    if (o < 0) return;                                // padding lane of a compacted set
    if (a.theta_out) {
        double s = 0.0;
        for (int v = 0; v < a.V; ++v) { const double zz = a.z[(int64_t)v * a.B + b]; s += zz * zz * a.invL[v]; }
        for (int v = 0; v < a.V; ++v) {
            const double zz = a.z[(int64_t)v * a.B + b];
            a.theta_out[(int64_t)v * a.B0 + o] = (st == TVS_INFEASIBLE) ? nan("") : zz * zz * a.invL[v] / s;
        }
    }
    if (a.ll_out)
        a.ll_out[o] = (st == TVS_INFEASIBLE) ? -INFINITY
                                             : a.llsum[b] - a.N[b] * log(a.sumphi[b]) + (a.C ? a.C[b] : 0.0);
    if (a.kkt_out) { a.kkt_out[2 * o] = a.kkt[b]; a.kkt_out[2 * o + 1] = a.kkt[a.B + b]; }
    if (a.iters_out) a.iters_out[o] = a.iters[b] < 0 ? 0 : a.iters[b];
    if (a.status_out) a.status_out[o] = st;
}

// ------------------------------------------------------------------------------------------------
// Lane compaction: idx = columns with done == 0 in increasing order; gather_rows copies those columns
// of a [rows x Bs] lane-minor matrix into a [rows x Bd] one.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) build_index_kernel(const int32_t* __restrict__ done, int64_t B,
                                                           int32_t* __restrict__ idx, int32_t* __restrict__ count) {
    __shared__ int32_t part[1024];
    const int t = threadIdx.x;
    const int64_t seg = (B + 1023) / 1024;
    const int64_t b0 = (int64_t)t * seg, b1 = min(B, b0 + seg);
    int32_t c = 0;
    for (int64_t b = b0; b < b1; ++b) c += (done[b] == 0);
    part[t] = c;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {         // inclusive scan
        const int32_t v = (t >= off) ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int32_t pos = part[t] - c;
    for (int64_t b = b0; b < b1; ++b)
        if (done[b] == 0) idx[pos++] = (int32_t)b;
    if (t == 1023) {
        *count = part[1023];
        for (int q = 0; q < 3; ++q) idx[part[1023] + q] = -1;     // padding entries (see gather_rows_kernel)
    }
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows, int64_t Bs, int64_t Bd,
                                   const int32_t* __restrict__ idx) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Bd) return;
    const int64_t sj = idx[j];                              // -1: padding lane -> zeros
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) dst[r * Bd + j] = (sj >= 0) ? src[r * Bs + sj] : T(0);
}

// All per-lane matrices of a lane set in ONE launch: task k copies columns idx[j] of src (rows x Bs, 4- or 8-byte
// elements) into dst (rows x Bd); idx[j] < 0 -> zeros.  grid.y walks (task, row) pairs.
struct GatherTask { const void* src; void* dst; int64_t rows; int32_t elem; int32_t row0; };   // row0 = first global row index of the task
constexpr int kMaxGatherTasks = 32;
struct GatherArgs { GatherTask t[kMaxGatherTasks]; int n; int64_t total_rows; int64_t Bs, Bd; const int32_t* idx; };

__global__ void __launch_bounds__(128) gather_all_kernel(const GatherArgs a) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.Bd) return;
    const int64_t sj = a.idx[j];
    for (int64_t gr = blockIdx.y; gr < a.total_rows; gr += gridDim.y) {
        int k = 0;
        while (k + 1 < a.n && gr >= a.t[k + 1].row0) ++k;         // tasks are few: linear search, uniform across the CTA
        const int64_t r = gr - a.t[k].row0;
        if (a.t[k].elem == 8) {
            const double v = (sj >= 0) ? static_cast<const double*>(a.t[k].src)[r * a.Bs + sj] : 0.0;
            static_cast<double*>(a.t[k].dst)[r * a.Bd + j] = v;
        } else if (a.t[k].elem == 4) {
            const int32_t v = (sj >= 0) ? static_cast<const int32_t*>(a.t[k].src)[r * a.Bs + sj] : 0;
            static_cast<int32_t*>(a.t[k].dst)[r * a.Bd + j] = v;
        } else if (a.t[k].elem == 2) {
            const uint16_t v = (sj >= 0) ? static_cast<const uint16_t*>(a.t[k].src)[r * a.Bs + sj] : (uint16_t)0;
            static_cast<uint16_t*>(a.t[k].dst)[r * a.Bd + j] = v;
        } else {
            const uint8_t v = (sj >= 0) ? static_cast<const uint8_t*>(a.t[k].src)[r * a.Bs + sj] : (uint8_t)0;
            static_cast<uint8_t*>(a.t[k].dst)[r * a.Bd + j] = v;
        }
    }
}

__global__ void widen_u16_kernel(const uint16_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 4 <= n) {
        const ushort4 v = *reinterpret_cast<const ushort4*>(src + i);
        *reinterpret_cast<int4*>(dst + i) = make_int4(v.x, v.y, v.z, v.w);
    } else {
        for (int64_t k = i; k < n; ++k) dst[k] = src[k];
    }
}

// int32 -> uint16 copy of the weights; *overflow is set when a count does not fit (the 16-bit copy is then not used)
__global__ void narrow_u16_kernel(const int32_t* __restrict__ src, uint16_t* __restrict__ dst, int64_t n, int32_t* __restrict__ overflow) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    bool bad = false, neg = false;                        // flag bit 0: a count above 65535, bit 1: a NEGATIVE weight
    if (i + 4 <= n) {
        const int4 v = *reinterpret_cast<const int4*>(src + i);
        bad = ((unsigned)v.x | (unsigned)v.y | (unsigned)v.z | (unsigned)v.w) > 65535u;
        neg = (v.x | v.y | v.z | v.w) < 0;
        *reinterpret_cast<ushort4*>(dst + i) = make_ushort4((unsigned short)v.x, (unsigned short)v.y, (unsigned short)v.z, (unsigned short)v.w);
    } else {
        for (int64_t k = i; k < n; ++k) { bad = bad || (unsigned)src[k] > 65535u; neg = neg || src[k] < 0; dst[k] = (uint16_t)src[k]; }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(overflow, 1);
    if (__any_sync(0xffffffffu, neg) && (threadIdx.x & 31) == 0) atomicOr(overflow, 2);
}

// lane columns from the device-resident column store: dst[r, b] = store[r, slot[b]]  (replaces the per-batch rebuild of the
// replicate weights, traversome.py:515-545)
template <typename T>
__global__ void gather_columns_kernel(const T* __restrict__ store, int64_t cap, const int32_t* __restrict__ slot, int64_t R, int64_t B,
                                      T* __restrict__ dst) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t sj = slot[b];
    for (int64_t r = blockIdx.y; r < R; r += gridDim.y) dst[r * B + b] = store[r * cap + sj];
}
// 16-bit mirror of columns first .. first+n-1 of the store; *flag bit 0 = a count above 65535, bit 1 = a negative weight
__global__ void narrow_columns_kernel(const int32_t* __restrict__ store, uint16_t* __restrict__ store16, int64_t cap, int64_t first, int64_t n,
                                      int64_t R, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * n) return;
    const int64_t r = i / n, c = i - r * n;
    const int32_t v = store[r * cap + first + c];
    store16[r * cap + first + c] = (uint16_t)v;
    if ((unsigned)v > 65535u) atomicOr(flag, v < 0 ? 3 : 1);
}
__global__ void widen_columns_kernel(const uint16_t* __restrict__ src, int64_t n, int64_t R, int32_t* __restrict__ store, int64_t cap, int64_t first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // src [R x n] lane-minor
    if (i >= R * n) return;
    const int64_t r = i / n, c = i - r * n;
    store[r * cap + first + c] = src[i];
}

// All the solver-state arrays of a fit reset in ONE launch (tvs_fit issued ~25 memsets before): task k fills `bytes`
// (a multiple of 4) at p with the 32-bit pattern `fill`.  grid.y = task.
struct FillTask { void* p; size_t bytes; uint32_t fill; };
constexpr int kMaxFillTasks = 28;
struct FillArgs { FillTask t[kMaxFillTasks]; int n; };
__global__ void __launch_bounds__(256) fill_many_kernel(const FillArgs a) {
    const FillTask k = a.t[blockIdx.y];
    const size_t n16 = k.bytes / 16;
    const uint4 v = make_uint4(k.fill, k.fill, k.fill, k.fill);
    uint4* p16 = static_cast<uint4*>(k.p);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p16[i] = v;
    uint32_t* p4 = static_cast<uint32_t*>(k.p) + n16 * 4;
    const size_t rest = (k.bytes - n16 * 16) / 4;
    if (blockIdx.x == 0 && threadIdx.x < rest) p4[threadIdx.x] = k.fill;
}

__global__ void iota_kernel(int32_t* p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

// ------------------------------------------------------------------------------------------------
// Device bootstrap: Philox4x32-10, key = (seed_lo, seed_hi), counter = (draw_lo, draw_hi, lane, 0x54565321)
// record index = floor(u64 * Ntot / 2^64); row = first r with cum[r] > index.
// replaces: random.choices(records_pool_sorted, k=N) + regrouping by read path (traversome.py:1652-1660)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                               uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Lane b (grid.y) is written with stride `ld` (element (r, b) at w[r*ld + b]); its Philox stream is keyed by id0 + b, so a
// column of the column store (tvs_columns_resample) depends on its id only.  keep_lane0: stream 0 is the observed data.
__global__ void resample_kernel(const int64_t* __restrict__ cum, int64_t R, int64_t Ntot, int64_t ld, uint64_t seed,
                                int keep_lane0, const int32_t* __restrict__ n_obs, int32_t* __restrict__ w, uint32_t id0) {
    const int64_t b = blockIdx.y;
    if (keep_lane0 && b == 0 && id0 == 0) {
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x)
            w[r * ld] = n_obs[r];
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Ntot; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t o[4];
        philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)b + id0, 0x54565321u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
        const uint64_t u = ((uint64_t)o[1] << 32) | o[0];
        const uint64_t idx = __umul64hi(u, (uint64_t)Ntot);
        int64_t lo = 0, hi = R - 1;                 // first r with cum[r] > idx
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(cum + mid) > (int64_t)idx) hi = mid; else lo = mid + 1;
        }
        atomicAdd(w + lo * ld + b, 1);
    }
}

// fp64 FMA peak: 8 independent chains per thread
__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// shared-memory read peak: every thread streams 16-byte words of a 32 KB tile, conflict-free (a warp reads 512 contiguous
// bytes per instruction), 8 independent loads in flight
__global__ void __launch_bounds__(512) smem_peak_kernel(double* out, int iters) {
    __shared__ double2 tile[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = make_double2(1.0 + i, 2.0);
    __syncthreads();
    double2 acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = make_double2(0.0, 0.0);
    unsigned idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double2 v = tile[(idx + 256u * j) & 2047u];     // 8 distinct rows of 32 x 16 bytes
            acc[j].x += v.x; acc[j].y += v.y;
        }
        idx = (idx + 32u) & 2047u;
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j].x + acc[j].y;
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace tvs
