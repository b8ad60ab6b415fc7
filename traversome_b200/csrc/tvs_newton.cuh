// Damped Newton phase of tvs_fit for small lane sets (sm_100a).  Included by tvs_api.cu.
//
// replaces: the last ~2/3 of the iterations of scipy's SLSQP multistart (ModelFitMaxLike.py:19-58) on the
// few lanes that batched L-BFGS leaves behind (profiles/r1_v3_summary.md: ~130 iterations at <= 48 lanes,
// ~55 us each whatever the width), and the whole fit of small batches (model-selection chunks, cfg1).
//
// Same objective as the L-BFGS phase, F(z) = -(1/N) sum_r w_r log (M z^2)_r + sum_v z_v^2, whose Hessian is
//     Hz = 2 diag(1 - g) + 4 (z z^T) o H ,   H_uv = (1/N) sum_r w_r M_ru M_rv / p_r^2 ,  M_rv = A_rv / L_v .
// One Newton iteration of a lane set =
//   hess_kernel     H of every lane that moved, as per-split partial sums (one warp = one (lane, row range);
//                   the warp owns a private packed-triangular H in shared memory and walks its rows in order,
//                   so the summation order is fixed: bitwise reproducible, no atomics);
//   solve_kernel    one CTA per lane: sums the partials, forms Hz + lambda I, Cholesky (lambda grows until
//                   the factorisation succeeds: Levenberg-Marquardt damping, the z-parametrisation is not
//                   convex), solves for the step, writes the trial point;
//   pass kernel     the ordinary fused likelihood pass at the trial point;
//   newton_update_kernel   Armijo accept / reject (reject -> lambda x 10 and re-solve with the same H),
//                   KKT residuals, convergence flags.
// Available when V <= 128 (H is 2,080 doubles per lane at V = 64, 8,256 at 128) and the row-product lists fit
// (tvs_create builds them).
#pragma once
#include "tvs_kernels.cuh"
#include "tvs_pass2.cuh"

namespace tvs {

constexpr int kNewtonMaxV = 128;   // solve_kernel keeps V(V+1)/2 / 256 matrix entries per thread in registers: instantiated for 64 and 128

// Hessian accumulation.  tvs_create cuts the rows into blocks of HB rows and, per block, stores the products
// A_ru A_rv (u >= v) of every row GROUPED BY H ENTRY: entries sorted by their number of contributing rows, 32
// entries to a slab, slab e = len_e steps of 32 (local row, product) pairs (thread t of the warp owns entry
// t of the slab; shorter lists padded with (row HB, 0), cs[HB] = 0).  A warp = one (lane, range of blocks):
// it loads c_r = w_r / p_r^2 of the block's rows (written by the likelihood pass at the accepted point) into
// shared memory, then streams the block's steps -- coalesced loads, private accumulators, no atomics, a fixed
// summation order -- and adds each slab's sums into its packed-triangular H in shared memory.
struct HessArgs {
    const int64_t* slab_ptr;      // [nblk * (NG + 1)] step offsets
    const uint16_t* h_row;        // [steps * 32]
    const float* h_val;           // [steps * 32]
    const uint16_t* h_slot;       // [nblk * NG * 32] packed H index of (block, slab, thread); 0xFFFF = none
    const double* cr;             // [R * W] w / p^2 at the accepted point (lane-minor)
    const int32_t* done; const int32_t* need;
    double* Hpart;                // [S][W][NP]
    int64_t W, R; int V; int HB; int nblk; int NG; int blocks_per_split;
};

constexpr int kHessBatch = 16;    // steps whose loads are issued together
constexpr int kHessWarps = 4;     // warps of one CTA = one (lane, range of blocks); they split each block's slabs

struct HessBuf { int rw[kHessBatch]; float vl[kHessBatch]; };

__device__ __forceinline__ void hess_load(HessBuf& b, const HessArgs& a, int64_t base, int64_t send, int t) {
#pragma unroll
    for (int i = 0; i < kHessBatch; ++i) {
        const int64_t q = (base + i < send) ? base + i : send - 1;      // clamp: a valid address, weight 0 below
        b.rw[i] = __ldg(a.h_row + q * 32 + t);
        b.vl[i] = (base + i < send) ? __ldg(a.h_val + q * 32 + t) : 0.f;
    }
}

__global__ void __launch_bounds__(32 * kHessWarps) hess_kernel(const HessArgs a) {
    extern __shared__ double hs[];                 // Hs[NP] | cs[HB + 1] | sp[NG + 1] (int64) | slot[NG * 32] (uint16)
    const int V = a.V, NP = V * (V + 1) / 2, HB = a.HB, NG = a.NG;
    double* Hs = hs;
    double* cs = hs + NP;
    int64_t* sp = reinterpret_cast<int64_t*>(cs + HB + 1);
    uint16_t* slots = reinterpret_cast<uint16_t*>(sp + NG + 1);
    const int tid = threadIdx.x, t = tid & 31, wi = tid >> 5;
    const int64_t lane = blockIdx.x;
    const int split = blockIdx.y;
    if (a.done[lane] != 0 || a.need[lane] == 0) return;
    for (int i = tid; i < NP; i += 32 * kHessWarps) Hs[i] = 0.0;
    const int b0 = split * a.blocks_per_split, b1 = min(a.nblk, b0 + a.blocks_per_split);
    for (int blk = b0; blk < b1; ++blk) {
        __syncthreads();
        const int64_t r0 = (int64_t)blk * HB;
        const int nrows = (int)min((int64_t)HB, a.R - r0);
        for (int i = tid; i <= HB; i += 32 * kHessWarps) cs[i] = (i < nrows) ? __ldg(a.cr + (r0 + i) * a.W + lane) : 0.0;
        for (int e = tid; e <= NG; e += 32 * kHessWarps) sp[e] = __ldg(a.slab_ptr + (int64_t)blk * (NG + 1) + e);
        for (int i = tid; i < NG * 32; i += 32 * kHessWarps) slots[i] = __ldg(a.h_slot + (int64_t)blk * NG * 32 + i);
        __syncthreads();
        // this warp's slabs: a contiguous range holding ~1/kHessWarps of the block's steps (cut at slab boundaries)
        const int64_t tot0 = sp[0], tot = sp[NG] - tot0;
        int e_lo = 0, e_hi = NG;
        {
            const int64_t lo_target = tot0 + tot * wi / kHessWarps, hi_target = tot0 + tot * (wi + 1) / kHessWarps;
            while (e_lo < NG && sp[e_lo] < lo_target) ++e_lo;
            e_hi = e_lo;
            while (e_hi < NG && sp[e_hi] < hi_target) ++e_hi;
            if (wi == kHessWarps - 1) e_hi = NG;
        }
        if (e_lo < e_hi) {
            int64_t s = sp[e_lo];
            const int64_t send = sp[e_hi];
            int e = e_lo;
            int64_t nb = sp[e + 1];
            double acc = 0.0;
            auto flush = [&]() {
                const int slot = slots[e * 32 + t];
                if (slot != 0xFFFF) Hs[slot] += acc;          // slabs are owned by exactly one warp: no race
                acc = 0.0;
            };
            auto compute = [&](const HessBuf& b) {
#pragma unroll
                for (int i = 0; i < kHessBatch; ++i) {
                    if (s < send) {
                        while (s == nb) { flush(); ++e; nb = sp[e + 1]; }
                        acc = fma(cs[b.rw[i]], (double)b.vl[i], acc);
                        ++s;
                    }
                }
            };
            HessBuf A, Bf;
            if (s < send) hess_load(A, a, s, send, t);
            while (s < send) {
                if (s + kHessBatch < send) hess_load(Bf, a, s + kHessBatch, send, t);
                compute(A);
                if (!(s < send)) break;
                if (s + kHessBatch < send) hess_load(A, a, s + kHessBatch, send, t);
                compute(Bf);
            }
            flush();
        }
    }
    __syncthreads();
    double* out = a.Hpart + ((int64_t)split * a.W + lane) * NP;
    for (int i = tid; i < NP; i += 32 * kHessWarps) out[i] = Hs[i];
}


// Hessian accumulation for WIDE lane sets: the same sum as hess_kernel organised like the transposed product of the
// likelihood pass.  A warp = 32 lanes (lane-minor, coalesced) x 32 H entries; a CTA = 8 warps = 256 entries of one
// lane tile.  Rows are walked in blocks of HB: c_r = w_r / p_r^2 of the block (fp32) and the block's records
// {row, entry, A_ru A_rv} of the CTA's entries -- stored ROW-MAJOR per (block, warp subset), so the inner loop is one
// flat loop with no per-entry control flow -- are staged by cp.async (double-buffered).  Each record is one
// read-modify-write of the warp's private fp32 accumulators in shared memory (conflict-free: lane-minor); every
// `kWideFlush` blocks the fp32 sums are widened with integer ops (F2F.F64.F32 runs on the XU pipe at ~4 lanes/clk/SM:
// profiles/r1_v4_summary.md) and added to fp64 registers.  The Hessian of a damped Newton method tolerates the 1e-7
// relative error of fp32 partial sums: it only steers the step, the KKT test certifies the answer.
struct HessWideArgs {
    const int32_t* wide_ptr;      // [nblk * (NS + 1)] record offsets per (block, warp subset); NS = ceil(NP / 32)
    const uint2* wide_rec;        // x = row | entry_in_subset << 16 ; y = fp32 bits of A_ru A_rv
    const float* crf;             // [R * W] w / p^2 (lane-minor)
    double* H;                    // [NP][W]
    int64_t W, R; int V; int HB; int nblk; int cap;    // cap = most records of one (block, 8 consecutive subsets)
};
constexpr int kWideWarps = 8;
constexpr int kWideEnt = 32;      // H entries per warp
constexpr int kWideSet = kWideWarps * kWideEnt;
constexpr int kWideFlush = 8;     // blocks between two widenings of the fp32 accumulators

__device__ __forceinline__ double widen_pos(float x) {          // exact float -> double for x = 0 or normal positive
    const unsigned b = __float_as_uint(x);
    const double w = __hiloint2double((int)(((b >> 3) & 0x0fffffffu) + 0x38000000u), (int)(b << 29));
    return (b >= 0x00800000u && b < 0x7f800000u) ? w : (double)x;
}

// shared: Cs[2][HB*32] float | rec[2][cap] uint2 | ptr[2][kWideWarps + 1] int | acc[kWideWarps][32][32] float
__global__ void __launch_bounds__(32 * kWideWarps) hess_wide_kernel(const HessWideArgs a) {
    extern __shared__ __align__(16) unsigned char smw[];
    const int V = a.V, NP = V * (V + 1) / 2, HB = a.HB, cap = a.cap;
    const int NS = (NP + kWideEnt - 1) / kWideEnt;
    float* Cs = reinterpret_cast<float*>(smw);
    uint2* recs = reinterpret_cast<uint2*>(smw + (size_t)2 * HB * 32 * sizeof(float));
    int* ptrs = reinterpret_cast<int*>(recs + (size_t)2 * cap);
    float* accs = reinterpret_cast<float*>(ptrs + 2 * (kWideWarps + 1) + 2);
    const int tid = threadIdx.x, t = tid & 31, wi = tid >> 5;
    const int64_t tile0 = (int64_t)blockIdx.x * 32;
    const int64_t lane = tile0 + t;
    const int sub0 = blockIdx.y * kWideWarps;                   // first warp subset of this CTA
    const int nsub = min(kWideWarps, NS - sub0);
    const bool full_tile = tile0 + 32 <= a.W;
    float* my = accs + (size_t)wi * kWideEnt * 32 + t;          // my[e * 32] = accumulator of entry e, this lane

    auto stage = [&](int blk, int buf) {
        const int64_t r0 = (int64_t)blk * HB;
        const int nrows = (int)min((int64_t)HB, a.R - r0);
        float* dstC = Cs + (size_t)buf * HB * 32;
        if (full_tile) {                                        // 16-byte copies: W % 4 == 0 is required by the caller
            for (int i = tid; i < nrows * 8; i += 32 * kWideWarps) {
                const int row = i >> 3, c4 = (i & 7) * 4;
                const unsigned d = (unsigned)__cvta_generic_to_shared(dstC + row * 32 + c4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(a.crf + (r0 + row) * a.W + tile0 + c4));
            }
        } else {
            for (int i = tid; i < nrows * 32; i += 32 * kWideWarps) {
                const int row = i >> 5, l = i & 31;
                dstC[i] = (tile0 + l < a.W) ? __ldg(a.crf + (r0 + row) * a.W + tile0 + l) : 0.f;
            }
        }
        const int32_t* pp = a.wide_ptr + (int64_t)blk * (NS + 1) + sub0;
        const int qb = __ldg(pp), qe = __ldg(pp + nsub);
        for (int i = tid; i < qe - qb; i += 32 * kWideWarps) cp_async8(recs + (size_t)buf * cap + i, a.wide_rec + qb + i);
        for (int i = tid; i <= nsub; i += 32 * kWideWarps) cp_async4(ptrs + buf * (kWideWarps + 1) + i, pp + i);
    };

    double acc[kWideEnt];
#pragma unroll
    for (int j = 0; j < kWideEnt; ++j) { acc[j] = 0.0; my[j * 32] = 0.f; }
    stage(0, 0);
    cp_async_commit();
    for (int blk = 0; blk < a.nblk; ++blk) {
        const int buf = blk & 1;
        cp_async_wait_all();
        __syncthreads();                                        // block `blk` landed; buffer buf^1 is free again
        if (blk + 1 < a.nblk) stage(blk + 1, buf ^ 1);
        cp_async_commit();
        if (wi < nsub) {
            const int* pt = ptrs + buf * (kWideWarps + 1);
            const int base = pt[0];
            const uint2* rc = recs + (size_t)buf * cap;
            const float* cb = Cs + (size_t)buf * HB * 32 + t;
            const int q1 = pt[wi + 1] - base;
#pragma unroll 4
            for (int q = pt[wi] - base; q < q1; ++q) {
                const uint2 r = rc[q];
                float* slot = my + (r.x >> 16) * 32;
                *slot = fmaf(__uint_as_float(r.y), cb[(r.x & 0xffffu) * 32], *slot);
            }
        }
        if ((blk + 1) % kWideFlush == 0 || blk + 1 == a.nblk) {
#pragma unroll
            for (int j = 0; j < kWideEnt; ++j) { acc[j] += widen_pos(my[j * 32]); my[j * 32] = 0.f; }
        }
    }
    if (lane < a.W && wi < nsub) {
#pragma unroll
        for (int j = 0; j < kWideEnt; ++j) {
            const int e = (sub0 + wi) * kWideEnt + j;
            if (e < NP) a.H[(int64_t)e * a.W + lane] = acc[j];
        }
    }
}

struct SolveArgs {
    const double* Hpart; int S;
    int64_t h_entry, h_lane, h_split;     // element (split, lane, entry) of Hpart at split*h_split + lane*h_lane + entry*h_entry
    const double* z; const double* g; const double* gha; const double* invL; const double* N; const uint8_t* mask;
    const int32_t* done;
    double* lam; double* d; double* zt; double* x; double* dg; double* alpha;
    int64_t W; int V;
    const double* sumphi; double tol_dual;     // saddle escape (see solve_kernel)
    const int32_t* ls;                         // consecutive rejected trial points of the lane (shrinks the escape step)
};

constexpr int kSolveThreads = 256;

// one CTA per lane.  shared: H0[V][V+1] (Hz, kept for retries) | 2 column buffers | rhs[V] | zv[V]
template <int VMAX>
__global__ void __launch_bounds__(kSolveThreads) solve_kernel(const SolveArgs a) {
    extern __shared__ double ss[];
    const int V = a.V, LD = V + 1, NP = V * (V + 1) / 2;
    double* H0 = ss;                                           // Hz; overwritten by the unscaled columns of L at the end
    double* colbuf = ss + (size_t)V * LD;                      // two column buffers of V + 2
    double* rhs = colbuf + 2 * (V + 2);
    double* zv = rhs + V;
    double* kickv = zv + V;                                    // target z of a component that leaves a saddle, else 0
    __shared__ double red[kSolveThreads / 32];
    const int tid = threadIdx.x;
    const int64_t lane = blockIdx.x;
    if (a.done[lane] != 0) return;
    const double invN = 1.0 / a.N[lane];
    for (int v = tid; v < V; v += kSolveThreads) zv[v] = a.z[(int64_t)v * a.W + lane];
    __syncthreads();
    // ---- Hz from the per-split partial sums, added in fixed split order
    for (int i = tid; i < NP; i += kSolveThreads) {
        int u = (int)((sqrt(8.0 * i + 1.0) - 1.0) * 0.5);
        while ((u + 1) * (u + 2) / 2 <= i) ++u;
        while (u * (u + 1) / 2 > i) --u;
        const int v = i - u * (u + 1) / 2;                     // u >= v
        const double s = sum_splits(a.Hpart + lane * a.h_lane + (int64_t)i * a.h_entry, a.S, a.h_split);
        double h = 4.0 * zv[u] * zv[v] * a.invL[u] * a.invL[v] * s * invN;
        if (u == v) {
            const bool mk = a.mask ? a.mask[(int64_t)u * a.W + lane] != 0 : true;
            const double gh_u = a.gha[(int64_t)u * a.W + lane];
            // Saddle escape: z_u = 0 is a stationary point of F for every u; where the multiplier test fails there
            // (g_u > 1: the variant wants a positive share) the gradient is ~0 and the curvature 2(1 - g_u) negative, so
            // neither Newton nor L-BFGS ever leaves.  Such a component is taken out of the linear system and moved to the
            // one-dimensional Newton point of phi_u: phi_u = (g_u - 1) / H_uu  (seen once in 20,000 cfg2 replicates).
            // The same holds NEAR the saddle: with phi_u two orders of magnitude below that point the z-space step only
            // grows z_u by ~15 % per iteration (26 iterations on one lane of the 8-GPU bench workload), so such a
            // component is moved the same way.  A rejected trial point shrinks the move (x 1/4 per rejection).
            const double sp = a.sumphi[lane];
            const double phi_u = zv[u] * zv[u];
            kickv[u] = 0.0;
            if (mk && gh_u * sp - 1.0 > a.tol_dual) {
                const double hphi = s * a.invL[u] * a.invL[u] * invN;                 // d2F / dphi_u^2
                const double full = fmin(fmax(gh_u - 1.0, 0.0) / fmax(hphi, 1e-300), 1e-3 * sp);
                if (phi_u < 1e-11 * sp || phi_u < 1e-2 * full) {
                    const double target = phi_u + full * exp2(-2.0 * (double)min(a.ls[lane], 20));
                    kickv[u] = sqrt(fmax(target, 1e-16 * sp));
                }
            }
            const bool fixed = !mk || kickv[u] > 0.0;
            h = fixed ? 1.0 : h + 2.0 * (1.0 - gh_u);
            rhs[u] = fixed ? 0.0 : -a.g[(int64_t)u * a.W + lane];
        }
        H0[u * LD + v] = h;
        H0[v * LD + u] = h;
    }
    __syncthreads();
    // rows / columns of fixed components (masked or kicked) carry no coupling
    for (int i = tid; i < V * V; i += kSolveThreads) {
        const int u = i / V, v = i - u * V;
        if (u != v) {
            const bool mu = a.mask ? a.mask[(int64_t)u * a.W + lane] != 0 : true;
            const bool mv = a.mask ? a.mask[(int64_t)v * a.W + lane] != 0 : true;
            if (!mu || !mv || kickv[u] > 0.0 || kickv[v] > 0.0) H0[u * LD + v] = 0.0;
        }
    }
    __syncthreads();
    double lam = a.lam[lane];
    double* invD = zv;                                         // zv is no longer needed: reuse as 1 / D_k
    // ---- LDL^T of Hz + lam I with the matrix in REGISTERS: thread tid owns the packed lower-triangle entries
    //      tid, tid + 256, ... (<= 9 for V <= 64, <= 33 for V <= 128) and, for tid < V, rhs[tid].  Step k: the owners of column k publish
    //      it (and rhs_k) to one of two shared buffers, one barrier, everyone updates its own entries.  The forward
    //      substitution rides along.  lam is raised until every pivot is positive (Levenberg-Marquardt damping).
    constexpr int EPT = (VMAX * (VMAX + 1) / 2 + kSolveThreads - 1) / kSolveThreads;
    int ei[EPT], ej[EPT];
    double av[EPT];
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
        const int e = tid + kSolveThreads * m;
        if (e < NP) {
            int u = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while ((u + 1) * (u + 2) / 2 <= e) ++u;
            while (u * (u + 1) / 2 > e) --u;
            ei[m] = u; ej[m] = e - u * (u + 1) / 2;
        } else { ei[m] = -1; ej[m] = -1; }
    }
    double rv = 0.0;
    for (int attempt = 0; attempt < 60; ++attempt) {
#pragma unroll
        for (int m = 0; m < EPT; ++m) av[m] = (ei[m] >= 0) ? H0[ei[m] * LD + ej[m]] + (ei[m] == ej[m] ? lam : 0.0) : 0.0;
        rv = (tid < V) ? rhs[tid] : 0.0;
        bool bad = false;
        for (int k = 0; k < V; ++k) {
            double* col = colbuf + (k & 1) * (V + 2);
#pragma unroll
            for (int m = 0; m < EPT; ++m)
                if (ej[m] == k) col[ei[m]] = av[m];
            if (tid == k) col[V] = rv;
            __syncthreads();
            const double piv = col[k];
            if (!(piv > 1e-12 * fabs(H0[k * LD + k] + lam) && piv > 0.0 && piv < INFINITY)) { bad = true; break; }
            const double rp = fast_rcp(piv);
            if (tid == 0) invD[k] = rp;
            const double bk = col[V];
#pragma unroll
            for (int m = 0; m < EPT; ++m)
                if (ej[m] > k) av[m] = fma(-col[ei[m]] * rp, col[ej[m]], av[m]);
            if (tid > k && tid < V) rv = fma(-col[tid] * rp, bk, rv);
        }
        __syncthreads();
        if (!bad) break;
        lam = fmax(10.0 * lam, 1e-6);
    }
    // unscaled columns of L and y back to shared memory for the backward substitution
    double* A = H0;
#pragma unroll
    for (int m = 0; m < EPT; ++m)
        if (ei[m] >= 0) A[ei[m] * LD + ej[m]] = av[m];
    if (tid < V) rhs[tid] = rv;
    __syncthreads();
    // ---- D^-1, then L^T d = y with L_ki = A_ki / D_i (one warp)
    if (tid < 32) {
        for (int k = tid; k < V; k += 32) rhs[k] *= invD[k];
        __syncwarp();
        for (int k = V - 1; k >= 0; --k) {
            const double xk = rhs[k];
            __syncwarp();
            for (int i = tid; i < k; i += 32) rhs[i] = fma(-A[k * LD + i] * invD[i], xk, rhs[i]);
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- step: d = rhs ; trial point ; directional derivative
    double part = 0.0;
    for (int v = tid; v < V; v += kSolveThreads) {
        const int64_t i = (int64_t)v * a.W + lane;
        const double zo = a.z[i];
        const double dv = (kickv[v] > 0.0) ? kickv[v] - zo : rhs[v];
        a.d[i] = dv;
        const double zn = zo + dv;
        a.zt[i] = zn;
        a.x[i] = zn * zn * a.invL[v];
        part = fma(dv, a.g[i], part);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int q = 0; q < kSolveThreads / 32; ++q) s += red[q];
        a.dg[lane] = s;
        a.alpha[lane] = 1.0;
        a.lam[lane] = lam;
    }
}

// Entry into the Newton phase: g-hat at the ACCEPTED point recovered from the stored z-space gradient
// g_z = 2 z (1 - g-hat); damping and flags reset.
__global__ void newton_init_kernel(const double* __restrict__ z, const double* __restrict__ g, const uint8_t* __restrict__ mask,
                                   const double* __restrict__ invL, int V, int64_t B, double* __restrict__ gha, double* __restrict__ x,
                                   double* __restrict__ lam, int32_t* __restrict__ need, int32_t* __restrict__ ls) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int v = 0; v < V; ++v) {
        const int64_t i = (int64_t)v * B + b;
        const bool mk = mask ? mask[i] != 0 : true;
        const double zz = z[i];
        gha[i] = (mk && zz != 0.0) ? 1.0 - g[i] / (2.0 * zz) : 1.0;
        x[i] = zz * zz * invL[v];                   // the refresh pass evaluates c_r = w_r / p_r^2 at the ACCEPTED point
    }
    lam[b] = 0.0; need[b] = 1; ls[b] = 0;
}

// Accept / reject of the Newton trial point: phases A, B (residuals only), C of lbfgs_update_vl2_kernel.
struct NewtonUpdArgs {
    const double* tpart; const double* llpart; int S;
    const double* invL; const double* N; const uint8_t* mask;
    double *z, *zt, *g, *gt, *gh, *gha, *d;
    double *F, *dg, *alpha, *llsum, *kkt, *sumphi, *lam;
    int32_t *ls, *iters, *status, *done, *need, *nactive;
    int V; int64_t B; int max_iter; double tol, tol_dual;
};

__global__ void __launch_bounds__(kU2Threads) newton_update_kernel(const NewtonUpdArgs a) {
    __shared__ double part[kU2Warps * 4 * kU2Lanes];
    __shared__ double accs[4][kU2Lanes];
    double (*acc)[kU2Lanes] = accs;
    const int tid = threadIdx.x, t = tid & 31, w = tid >> 5;
    const int l = tid & (kU2Lanes - 1), vs = tid / kU2Lanes;
    const int64_t B = a.B;
    const int64_t b = (int64_t)blockIdx.x * kU2Lanes + l;
    const bool valid = b < B;
    const int64_t bb = valid ? b : B - 1;
    const int V = a.V;
    const bool active = valid && a.done[bb] == 0;
    if (!__syncthreads_or(active)) return;

    const double N = a.N[bb];
    const double llsum = sum_splits(a.llpart + bb, a.S, B);
    double r2[2] = {0.0, 0.0};
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
    const double F = a.F[bb], dg = a.dg[bb], alpha = a.alpha[bb];
    bool accept = fin && (Ft <= F + 1e-4 * alpha * dg ||
                          (fabs(Ft - F) <= 1e-14 * fabs(F) && fabs(dgt) <= 0.9 * fabs(dg)));
    accept = accept && active;
    __syncthreads();
    double r3[2] = {0.0, -INFINITY};                            // comp, dual
    if (accept) {
        for (int v = vs; v < V; v += kU2Slots) {
            const int64_t i = (int64_t)v * B + bb;
            const bool mk = a.mask ? a.mask[i] != 0 : true;
            if (mk) {
                const double ztv = a.zt[i];
                const double ghn = a.gh[i] * sumphi;
                r3[0] = fmax(r3[0], ztv * ztv / sumphi * fabs(ghn - 1.0));
                r3[1] = fmax(r3[1], ghn - 1.0);
            }
            a.z[i] = a.zt[i]; a.g[i] = a.gt[i]; a.gha[i] = a.gh[i];
        }
    }
    cta_reduce<2>(r3, 0, part, acc, w, t, l);
    const double comp = acc[0][l], dual = acc[1][l];
    bool finished = false;
    int status = TVS_CONVERGED;
    if (vs == 0 && active) {
        if (accept) {
            const int iters = a.iters[bb] + 1;
            if (comp <= a.tol && dual <= a.tol_dual) finished = true;
            else if (iters >= a.max_iter) { finished = true; status = TVS_MAXITER; }
            const double lam = a.lam[bb];
            a.lam[bb] = (lam > 1e-12) ? 0.1 * lam : 0.0;
            a.F[bb] = Ft; a.llsum[bb] = llsum; a.sumphi[bb] = sumphi;
            a.kkt[bb] = comp; a.kkt[B + bb] = dual; a.iters[bb] = iters; a.ls[bb] = 0;
            a.need[bb] = 1;
        } else {
            const int ls = a.ls[bb] + 1;
            a.ls[bb] = ls;
            a.lam[bb] = fmax(10.0 * a.lam[bb], 1e-6);
            a.need[bb] = 0;
            if (ls > 40) {
                const double c0 = a.kkt[bb], d0 = a.kkt[B + bb];
                finished = true;
                status = (c0 <= 1e2 * a.tol && d0 <= 1e2 * a.tol_dual) ? TVS_CONVERGED : TVS_STALLED;
            }
        }
        if (finished) { a.status[bb] = status; a.done[bb] = 1; }
    }
    if (w == 0) {
        const unsigned still = __ballot_sync(0xffffffffu, t < kU2Lanes && active && !finished);
        if (t == 0 && still) atomicAdd(a.nactive, __popc(still));
    }
}

}  // namespace tvs
