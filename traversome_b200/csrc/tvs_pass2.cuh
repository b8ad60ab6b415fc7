// pass2_kernel -- second-generation fused likelihood pass (sm_100a).  Included by tvs_api.cu.
//
//     p = A x ;  rr = w / p ;  ll += w log p ;  t = A^T rr            (per lane)
//
// replaces: PathMultinomialModel.get_like_formula (ModelGenerator.py:153-174) evaluated by the
// lambdified objective (ModelFitMaxLike.py:534-545), once per optimiser function evaluation.
//
// What changed against pass_kernel (profiles/r1_v1_summary.md: issue-bound, 310 warp instructions
// per (row, 32 lanes) of which 13 were useful DFMA):
//   * K consecutive lanes per thread (K = 1, 2, 4): one entry decode / address computation feeds K
//     FMAs, the x / rr operands come from shared memory as 128-bit loads (conflict-free layout),
//     the weights from global memory as one K*4-byte load per thread.
//   * 1/p from `rcp.approx.ftz.f64` + one third-order step (4 instructions instead of the ~40 of an
//     IEEE division) and log p from an inlined atanh series on the reduced mantissa (no table: the
//     binding resource of the sparse product is shared-memory operand bandwidth, so the logarithm
//     is kept on the fp64 pipe).  Both are accurate to ~1 ulp; tests/test_gpu_parity.py pins them.
//   * 16 warps per CTA.
// Lane l of a tile lives at slot(l) inside every shared-memory row so that the two 16-byte halves
// of a thread's K = 4 lanes are each contiguous across the warp.
#pragma once
#include "tvs_internal.h"
#include <math.h>

namespace tvs {

constexpr int kWarps2 = 16;
constexpr int kThreads2 = kWarps2 * 32;

// coefficients of atanh(s)/s - 1 = s^2/3 + s^4/5 + ... travel as KERNEL PARAMETERS (LogCoef inside the argument
// struct): bank-0 constants are legal DFMA operands, so the polynomial costs no loads or moves (a __constant__ array
// costs one LDCU per use: profiles/r1_v2_summary.md).
__host__ __device__ inline LogCoef make_log_coef() {
    LogCoef k;
    for (int i = 0; i < 9; ++i) k.c[i] = 1.0 / (double)(2 * i + 3);
    k.c[9] = 6.93147180369123816490e-01;      // ln2_hi, 32 significant bits
    k.c[10] = 1.90821492927058770002e-10;     // ln2_lo
    return k;
}

__device__ __forceinline__ bool is_pos_normal(double p) {       // finite, normal, > 0 (one integer compare)
    return (unsigned)(__double2hiint(p) - 0x00100000) < 0x7fe00000u;
}

// int32 -> double without the XU pipe (I2F.F64 and F2F.F64 run there at a few lanes/clk/SM and the pipe saturates:
// profiles/r1_v4_summary.md): place the integer in the mantissa of 2^52 and subtract.
__device__ __forceinline__ double u31_to_double(int v) {        // 0 <= v < 2^31
    return __hiloint2double(0x43300000, v) - 4503599627370496.0;
}
__device__ __forceinline__ double i32_to_double(int v) {        // any int32
    return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
}

__device__ __forceinline__ double fast_rcp(double p) {          // p finite, normal, > 0
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));        // MUFU.RCP64H, ~2^-20
    const double e = fma(-p, r, 1.0);                            // |e| <= ~2^-20 (the low 32 mantissa bits of p are ignored)
    return fma(r, fma(e, e, e), r);                              // r (1 + e + e^2) = (1 - e^3) / p: ~1 ulp in 3 FMAs (was 2 Newton steps, 4)
}

// natural logarithm for finite normal p > 0: p = 2^e m, m in [sqrt(1/2), sqrt(2)),
// log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716.
__device__ __forceinline__ double fast_log_pos(double p, const LogCoef& kc) {
    int hi = __double2hiint(p);
    const int lo = __double2loint(p);
    int e = (hi >> 20) - 1023;
    int mh = (hi & 0x000fffff) | 0x3ff00000;
    if (mh >= 0x3ff6a09f) { mh -= 0x00100000; e += 1; }
    const double m = __hiloint2double(mh, lo);
    const double u = fast_rcp(m + 1.0);
    const double s = (m - 1.0) * u;
    const double s2 = s * s;
    double q = kc.c[8];
    q = fma(q, s2, kc.c[7]);
    q = fma(q, s2, kc.c[6]);
    q = fma(q, s2, kc.c[5]);
    q = fma(q, s2, kc.c[4]);
    q = fma(q, s2, kc.c[3]);
    q = fma(q, s2, kc.c[2]);
    q = fma(q, s2, kc.c[1]);
    q = fma(q, s2, kc.c[0]);
    const double ed = i32_to_double(e);
    const double two_s = s + s;
    const double tail = fma(two_s * s2, q, ed * kc.c[10]);      // + e * ln2_lo
    return fma(ed, kc.c[9], two_s + tail);                      // e * ln2_hi is exact
}

// Table variant (pass3_kernel): p = 2^e m, m in [1,2); i = top 7 mantissa bits; tab[i] = {c_i ~ 1/m_i, -log c_i};
// r = m c_i - 1 (exact in one FMA, |r| <= 2^-8); log m = -log c_i + log1p(r), log1p by a degree-7 Taylor polynomial
// (truncation < 2^-67).  12 fp64 instructions instead of ~30; costs one 16-byte shared-memory read per call.
__device__ __forceinline__ double table_log_pos(double p, const double2* __restrict__ tab, const LogCoef& kc) {
    const int hi = __double2hiint(p);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(p));
    const double2 cl = tab[(hi >> 13) & 127];
    const double r = fma(m, cl.x, -1.0);
    double q = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    q = fma(q, r, 0.2);
    q = fma(q, r, -0.25);
    q = fma(q, r, kc.c[0]);                                      // 1/3
    q = fma(q, r, -0.5);
    const double ed = i32_to_double(e);
    const double tail = fma(r * r, q, fma(ed, kc.c[10], r));     // r + r^2 q + e ln2_lo
    return fma(ed, kc.c[9], cl.y) + tail;                        // e ln2_hi (exact) - log c_i
}

__device__ __forceinline__ double log_any(double p, const LogCoef& kc) {   // -inf at 0, nan below, slow path for odd inputs
    if (is_pos_normal(p)) return fast_log_pos(p, kc);
    return log(p);
}

__global__ void selftest_math_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ lg, double* __restrict__ rc,
                                     const LogCoef kc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lg[i] = log_any(p[i], kc);
    rc[i] = is_pos_normal(p[i]) ? fast_rcp(p[i]) : 1.0 / p[i];
}

__global__ void selftest_table_log_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ lg, const double2* __restrict__ tab,
                                          const LogCoef kc) {
    __shared__ double2 tab_s[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) tab_s[i] = tab[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lg[i] = is_pos_normal(p[i]) ? table_log_pos(p[i], tab_s, kc) : log(p[i]);
}

template <int K>
__device__ __forceinline__ int lane_slot(int l) {               // position of tile lane l inside a shared row
    if (K == 4) return ((l >> 1) & 1) * 64 + (l >> 2) * 2 + (l & 1);
    return l;
}

// entry decode of pass3_kernel: the count travels as the top 16 bits of its fp64 form (Csr::ent3 / cent3), so it is rebuilt
// with one LOP3 and no fp64 instruction; operand address = the thread's base + index * row bytes.
__device__ __forceinline__ double ent_count(unsigned e) { return __hiloint2double((int)(e & 0xffff0000u), 0); }
template <int K>
__device__ __forceinline__ void lds_lanes_at(const double* tbase, unsigned e, double* out) {   // tbase = tile + this thread's slot
    constexpr int LT = 32 * K;
    const char* row = reinterpret_cast<const char*>(tbase) + (e & 0xffffu) * (unsigned)(LT * 8);
    if (K == 1) out[0] = *reinterpret_cast<const double*>(row);
    else if (K == 2) { const double2 a = *reinterpret_cast<const double2*>(row); out[0] = a.x; out[1] = a.y; }
    else {
        const double2 a = *reinterpret_cast<const double2*>(row);
        const double2 b = *reinterpret_cast<const double2*>(row + 512);
        out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
    }
}

template <int K> struct WVec;
template <> struct WVec<1> { int v[1]; };
template <> struct __align__(8) WVec<2> { int v[2]; };
template <> struct __align__(16) WVec<4> { int v[4]; };

// K consecutive lanes' weights of one row: 16-bit counts when the lane set carries them, else int32
template <int K>
__device__ __forceinline__ WVec<K> load_w(const PassArgs& a, int64_t idx) {
    WVec<K> o;
    if (a.w16 != nullptr) {
        if (K == 1) o.v[0] = a.w16[idx];
        else if (K == 2) { const unsigned u = *reinterpret_cast<const unsigned*>(a.w16 + idx); o.v[0] = (int)(u & 0xffffu); o.v[K - 1] = (int)(u >> 16); }
        else {
            const uint2 u = *reinterpret_cast<const uint2*>(a.w16 + idx);
            o.v[0] = (int)(u.x & 0xffffu); o.v[1 % K] = (int)(u.x >> 16); o.v[2 % K] = (int)(u.y & 0xffffu); o.v[3 % K] = (int)(u.y >> 16);
        }
    } else o = *reinterpret_cast<const WVec<K>*>(a.w + idx);
    return o;
}

// The same in two steps, so that a load issued one row ahead is not waited for before the row it belongs to: the RAW
// words travel across the loop trip and are unpacked where the weights are used.  (With the unpacking inside the load the
// warp stalled on its own prefetch: 17 % of the stall samples of the pass, profiles/r2_pass3_full.csv.)
template <int K>
__device__ __forceinline__ WVec<K> load_w_raw(const PassArgs& a, int64_t idx) {
    WVec<K> o;
#pragma unroll
    for (int j = 0; j < K; ++j) o.v[j] = 0;
    if (a.w16 != nullptr) {
        if (K == 1) o.v[0] = a.w16[idx];
        else if (K == 2) o.v[0] = (int)*reinterpret_cast<const unsigned*>(a.w16 + idx);
        else { const uint2 u = *reinterpret_cast<const uint2*>(a.w16 + idx); o.v[0] = (int)u.x; o.v[1 % K] = (int)u.y; }
    } else o = *reinterpret_cast<const WVec<K>*>(a.w + idx);
    return o;
}
template <int K>
__device__ __forceinline__ WVec<K> unpack_w(const WVec<K>& raw, bool is16) {
    if (!is16 || K == 1) return raw;
    WVec<K> o;
    if (K == 2) { const unsigned u = (unsigned)raw.v[0]; o.v[0] = (int)(u & 0xffffu); o.v[K - 1] = (int)(u >> 16); }
    else {
        const unsigned ux = (unsigned)raw.v[0], uy = (unsigned)raw.v[1 % K];
        o.v[0] = (int)(ux & 0xffffu); o.v[1 % K] = (int)(ux >> 16); o.v[2 % K] = (int)(uy & 0xffffu); o.v[3 % K] = (int)(uy >> 16);
    }
    return o;
}

template <int K>
__device__ __forceinline__ void lds_lanes(const double* row, int t, double* out) {   // out[K] = lanes K*t .. K*t+K-1
    if (K == 1) out[0] = row[t];
    else if (K == 2) { const double2 a = reinterpret_cast<const double2*>(row)[t]; out[0] = a.x; out[1] = a.y; }
    else {
        const double2 a = reinterpret_cast<const double2*>(row)[t];
        const double2 b = reinterpret_cast<const double2*>(row)[32 + t];
        out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
    }
}
template <int K>
__device__ __forceinline__ void sts_lanes(double* row, int t, const double* in) {
    if (K == 1) row[t] = in[0];
    else if (K == 2) reinterpret_cast<double2*>(row)[t] = make_double2(in[0], in[1]);
    else {
        reinterpret_cast<double2*>(row)[t] = make_double2(in[0], in[1]);
        reinterpret_cast<double2*>(row)[32 + t] = make_double2(in[2], in[3]);
    }
}

// grid = (lane tiles, row splits); CTA (tile, split) walks chunks split, split+S, ...
// requires: B % K == 0 (vector loads of w), V*LT doubles of x (and t) fit in shared memory.
template <int K, bool BACKWARD>
__global__ void __launch_bounds__(kThreads2, 1) pass2_kernel(const PassArgs a) {
    constexpr int LT = 32 * K;
    extern __shared__ __align__(16) double smem2[];
    const int t = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int tile = blockIdx.x, split = blockIdx.y, S = gridDim.y;
    const int64_t B = a.B;
    const int V = a.V;
    const int64_t lane0 = (int64_t)tile * LT;
    const int64_t mylane = lane0 + K * t;                 // first of this thread's K lanes (B % K == 0)
    const bool lv = mylane < B;

    if (a.done != nullptr) {                              // skip tiles whose lanes have all finished
        int alive = 0;
        if (lv) {
#pragma unroll
            for (int j = 0; j < K; ++j) alive |= (a.done[mylane + j] == 0);
        }
        if (!__syncthreads_or(alive)) return;
    }

    double* x_s = smem2;                                            // [V][LT]
    double* t_s = x_s + (size_t)V * LT;                             // [V][LT]   (BACKWARD)
    double* rr_s = t_s + (BACKWARD ? (size_t)V * LT : 0);           // [rows_per_chunk][LT]

    for (int i = threadIdx.x; i < V * LT; i += kThreads2) {
        const int v = i / LT, l = i - v * LT;
        const int64_t b = lane0 + l;
        // 1.0 for padding and finished lanes: they stay on the fast rcp/log path, their results are never read
        const bool live = b < B && (a.done == nullptr || a.done[b] == 0);
        x_s[v * LT + lane_slot<K>(l)] = live ? a.x[(int64_t)v * B + b] : 1.0;
        if (BACKWARD) t_s[i] = 0.0;
    }
    __syncthreads();

    double ll[K];
#pragma unroll
    for (int j = 0; j < K; ++j) ll[j] = 0.0;

    for (int c = split; c < a.n_chunks; c += S) {
        const int row0 = a.chunk_row0[c], row1 = a.chunk_row0[c + 1];
        // ---------------- phase 1: p = A x over the rows of the chunk
        for (int r = row0 + wi; r < row1; r += kWarps2) {
            const int k0 = __ldg(a.row_ptr + r), k1 = __ldg(a.row_ptr + r + 1);
            WVec<K> wv;
            if (lv) wv = load_w<K>(a, (int64_t)r * B + mylane);
            else {
#pragma unroll
                for (int j = 0; j < K; ++j) wv.v[j] = 0;
            }
            double p[K];
#pragma unroll
            for (int j = 0; j < K; ++j) p[j] = 0.0;
#pragma unroll 2
            for (int k = k0; k < k1; ++k) {
                const uint32_t e = __ldg(a.ent + k);
                const double f = (double)(e >> 16);
                double xv[K];
                lds_lanes<K>(x_s + (e & 0xffffu) * LT, t, xv);
#pragma unroll
                for (int j = 0; j < K; ++j) p[j] = fma(f, xv[j], p[j]);
            }
            double rr[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double wd = (double)wv.v[j];
                const bool on = wv.v[j] > 0;
                double rcp, lg;
                if (is_pos_normal(p[j])) { rcp = fast_rcp(p[j]); lg = fast_log_pos(p[j], a.logc); }
                else { rcp = 1.0 / p[j]; lg = log(p[j]); }
                rr[j] = on ? wd * rcp : 0.0;
                ll[j] = on ? fma(wd, lg, ll[j]) : ll[j];
            }
            if (BACKWARD) sts_lanes<K>(rr_s + (r - row0) * LT, t, rr);
        }
        if (BACKWARD) {
            __syncthreads();
            // ------------ phase 2: t += A^T rr over the non-empty columns of the chunk
            const int c0 = a.ccol_ptr[c], c1 = a.ccol_ptr[c + 1];
            for (int ci = c0 + wi; ci < c1; ci += kWarps2) {
                const int v = __ldg(a.ccol_id + ci);
                const int k0 = __ldg(a.ccol_start + ci), k1 = __ldg(a.ccol_start + ci + 1);
                double acc[K];
                lds_lanes<K>(t_s + v * LT, t, acc);
#pragma unroll 2
                for (int k = k0; k < k1; ++k) {
                    const uint32_t e = __ldg(a.cent + k);
                    const double f = (double)(e >> 16);
                    double rv[K];
                    lds_lanes<K>(rr_s + (e & 0xffffu) * LT, t, rv);
#pragma unroll
                    for (int j = 0; j < K; ++j) acc[j] = fma(f, rv[j], acc[j]);
                }
                sts_lanes<K>(t_s + v * LT, t, acc);
            }
            __syncthreads();
        }
    }

    // ---------------- epilogue: per-split partials, summed later in fixed split order (deterministic)
    if (BACKWARD) {
        for (int i = threadIdx.x; i < V * LT; i += kThreads2) {
            const int v = i / LT, l = i - v * LT;
            const int64_t b = lane0 + l;
            if (b < B) a.tpart[((int64_t)split * V + v) * B + b] = t_s[v * LT + lane_slot<K>(l)];
        }
    }
    __syncthreads();
    double* red = rr_s;                                   // [kWarps2][LT]; rows_per_chunk >= kWarps2 enforced on the host
    sts_lanes<K>(red + wi * LT, t, ll);
    __syncthreads();
    if (wi == 0) {
        double s[K];
#pragma unroll
        for (int j = 0; j < K; ++j) s[j] = 0.0;
        for (int q = 0; q < kWarps2; ++q) {
            double v[K];
            lds_lanes<K>(red + q * LT, t, v);
#pragma unroll
            for (int j = 0; j < K; ++j) s[j] += v[j];
        }
        if (lv) {
#pragma unroll
            for (int j = 0; j < K; ++j) a.llpart[(int64_t)split * B + mylane + j] = s[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pass3_kernel: pass2 + the CSR entries of each chunk staged in shared memory by cp.async
// (profiles/r1_v2_summary.md: pass2 stalls on the dependent LDG(entry) -> LDS(x) -> DFMA chain).
//   * entries stay packed (column or chunk-local row | top 16 bits of the fp64 count << 16, one 32-bit word; Csr::ent3 /
//     cent3): the inner loops are LDS.32 (broadcast) + one LOP3 for the count (no fp64 conversion) + K/2 x LDS.128 + K x DFMA;
//   * the row-major entries of the NEXT chunk and the column-major entries of THIS chunk are in flight
//     while phase 1 runs (double buffer / single buffer);
//   * the weights of a warp's next row are loaded one row ahead (the only HBM stream of the pass) and travel across
//     the loop trip as RAW words (load_w_raw / unpack_w);
//   * V <= 4 * NW: the t accumulators of a warp's columns (wi, wi + NW, wi + 2 NW, wi + 3 NW) stay in registers for the
//     whole launch (pass3_treg): no t tile, so the chunks are longer (144 rows at V = 64, K = 4).
// shared: x_s[V][LT] (t_s[V][LT] unless pass3_treg) rr_s[RC][LT] | c1[2][cap] c2[cap] (packed entries) rp[2][RC+1]
//         cs[V+1] cid[V] (int32) | log table
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct Pass3Smem {            // byte offsets from the start of dynamic shared memory
    size_t x, t, rr, c1, c2, rp, cs, cid, tab, total;
};
// t accumulators in REGISTERS when every warp can own at most 4 columns for the whole launch (wi, wi + NW, wi + 2 NW,
// wi + 3 NW): no t tile in shared memory, so a chunk can hold 144 rows instead of 80 at V = 64, 4 lanes per thread
__host__ __device__ inline bool pass3_treg(int V, int nw) { return V <= 4 * nw; }
__host__ __device__ inline Pass3Smem pass3_layout(int V, int LT, int RC, int cap, bool backward, int nw) {
    Pass3Smem L;
    size_t o = 0;
    L.x = o; o += (size_t)V * LT * 8;
    L.t = o; o += (backward && !pass3_treg(V, nw)) ? (size_t)V * LT * 8 : 0;
    L.rr = o; o += (size_t)(RC > nw ? RC : nw) * LT * 8;
    L.c1 = o; o += (size_t)2 * cap * 4;
    L.c2 = o; o += backward ? (size_t)cap * 4 : 0;
    L.rp = o; o += (size_t)2 * (RC + 1) * 4;
    L.cs = o; o += backward ? (size_t)(V + 1) * 4 : 0;
    L.cid = o; o += backward ? (size_t)V * 4 : 0;
    o = (o + 15) & ~(size_t)15;
    L.tab = o; o += 128 * 16;                          // log table (table_log_pos)
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

template <int K, int NW, bool BACKWARD>
__global__ void __launch_bounds__(NW * 32, 1) pass3_kernel(const PassArgs a) {
    constexpr int LT = 32 * K;
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) unsigned char smem3[];
    const int t = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int tile = blockIdx.x, split = blockIdx.y, S = gridDim.y;
    const int64_t B = a.B;
    const int V = a.V, RC = a.rows_per_chunk, cap = a.stage_cap;
    const int64_t lane0 = (int64_t)tile * LT;
    const int64_t mylane = lane0 + K * t;
    const bool lv = mylane < B;

    if (a.done != nullptr) {
        int alive = 0;
        if (lv) {
#pragma unroll
            for (int j = 0; j < K; ++j) alive |= (a.done[mylane + j] == 0);
        }
        if (!__syncthreads_or(alive)) return;
    }

    const Pass3Smem L = pass3_layout(V, LT, RC, cap, BACKWARD, NW);
    const bool treg = BACKWARD && pass3_treg(V, NW);
    double* x_s = reinterpret_cast<double*>(smem3 + L.x);
    double* t_s = reinterpret_cast<double*>(smem3 + L.t);
    double* rr_s = reinterpret_cast<double*>(smem3 + L.rr);
    uint32_t* c1_s = reinterpret_cast<uint32_t*>(smem3 + L.c1);
    uint32_t* c2_s = reinterpret_cast<uint32_t*>(smem3 + L.c2);
    int* rp_s = reinterpret_cast<int*>(smem3 + L.rp);
    int* cs_s = reinterpret_cast<int*>(smem3 + L.cs);
    int* cid_s = reinterpret_cast<int*>(smem3 + L.cid);
    double2* tab_s = reinterpret_cast<double2*>(smem3 + L.tab);
    const double* x_t = x_s + (K == 1 ? t : 2 * t);      // this thread's slot inside an operand row (lds_lanes_at)
    const double* rr_t = rr_s + (K == 1 ? t : 2 * t);
    const bool is16 = a.w16 != nullptr;
    for (int i = threadIdx.x; i < 128; i += NT) tab_s[i] = a.logtab[i];        // the log table (tvs_create always uploads it)

    // row-major entries + row pointers of a chunk -> buffer `buf`
    auto stage_rows = [&](const ChunkDesc& d, int buf) {
        for (int i = threadIdx.x; i < d.nent; i += NT) {
            cp_async4(c1_s + (size_t)buf * cap + i, a.ent3 + d.e0 + i);
        }
        for (int i = threadIdx.x; i <= d.nrows; i += NT) cp_async4(rp_s + buf * (RC + 1) + i, a.row_ptr + d.row0 + i);
    };
    // column-major entries + column starts / ids of a chunk
    auto stage_cols = [&](const ChunkDesc& d) {
        for (int i = threadIdx.x; i < d.ncent; i += NT) {
            cp_async4(c2_s + i, a.cent3 + d.ce0 + i);
        }
        for (int i = threadIdx.x; i <= d.ncol; i += NT) cp_async4(cs_s + i, a.ccol_start + d.c0 + i);
        for (int i = threadIdx.x; i < d.ncol; i += NT) cp_async4(cid_s + i, a.ccol_id + d.c0 + i);
    };
    // chunk descriptors are fetched two chunks ahead: no dependent global loads inside the chunk loop
    auto load_desc = [&](int c) {
        ChunkDesc d;
        if (c < a.n_chunks) {
            const int4 lo = __ldg(reinterpret_cast<const int4*>(a.desc + c));
            const int4 hi = __ldg(reinterpret_cast<const int4*>(a.desc + c) + 1);
            d.row0 = lo.x; d.nrows = lo.y; d.e0 = lo.z; d.nent = lo.w; d.c0 = hi.x; d.ncol = hi.y; d.ce0 = hi.z; d.ncent = hi.w;
        } else d = ChunkDesc{0, 0, 0, 0, 0, 0, 0, 0};
        return d;
    };

    int c = split, buf = 0;
    ChunkDesc cur = load_desc(c), nxt = load_desc(c + S);
    stage_rows(cur, 0);
    cp_async_commit();

    {   // x tile: NT is a multiple of LT, so a thread always fills the same lane of the tile (one liveness test per thread)
        static_assert(NT % LT == 0, "threads per CTA must be a multiple of the lane tile");
        const int l = threadIdx.x % LT;
        const int64_t b = lane0 + l;
        const bool live = b < B && (a.done == nullptr || a.done[b] == 0);
        const int slot = lane_slot<K>(l);
        const double* src = a.x + (live ? b : 0);
#pragma unroll 4
        for (int v = threadIdx.x / LT; v < V; v += NT / LT) {
            x_s[v * LT + slot] = live ? src[(int64_t)v * B] : 1.0;
            if (BACKWARD && !treg) t_s[v * LT + l] = 0.0;
        }
    }

    double ll[K];
#pragma unroll
    for (int j = 0; j < K; ++j) ll[j] = 0.0;
    double tacc[4][K];                                     // treg: t of this warp's columns, in registers for the whole launch
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < K; ++j) tacc[q][j] = 0.0;

    for (; c < a.n_chunks; c += S, buf ^= 1) {
        const ChunkDesc nx2 = load_desc(c + 2 * S);        // consumed two trips from now
        const int row0 = cur.row0, row1 = cur.row0 + cur.nrows;
        cp_async_wait_all();
        __syncthreads();                                   // entries of chunk c (and x_s on the first trip) visible
        if (BACKWARD) stage_cols(cur);
        if (c + S < a.n_chunks) stage_rows(nxt, buf ^ 1);
        cp_async_commit();
        // ---------------- phase 1: p = A x over the rows of the chunk
        const uint32_t* c1 = c1_s + (size_t)buf * cap;
        const int* rp = rp_s + buf * (RC + 1);
        const int e0 = rp[0];
        WVec<K> wnext;                                     // RAW words of the next row's weights (load_w_raw)
#pragma unroll
        for (int j = 0; j < K; ++j) wnext.v[j] = 0;
        {
            const int r = row0 + wi;
            if (lv && r < row1) wnext = load_w_raw<K>(a, (int64_t)r * B + mylane);
        }
        for (int r = row0 + wi; r < row1; r += NW) {
            const WVec<K> wraw = wnext;
            if (lv && r + NW < row1) wnext = load_w_raw<K>(a, (int64_t)(r + NW) * B + mylane);
            const int k0 = rp[r - row0] - e0, k1 = rp[r - row0 + 1] - e0;
            double p[K];
#pragma unroll
            for (int j = 0; j < K; ++j) p[j] = 0.0;
#pragma unroll 4
            for (int k = k0; k < k1; ++k) {
                const unsigned e = c1[k];                            // column | top 16 bits of the fp64 count << 16
                const double f = ent_count(e);
                double xv[K];
                lds_lanes_at<K>(x_t, e, xv);
#pragma unroll
                for (int j = 0; j < K; ++j) p[j] = fma(f, xv[j], p[j]);
            }
            bool fast = true;
#pragma unroll
            for (int j = 0; j < K; ++j) fast = fast && is_pos_normal(p[j]);
            const WVec<K> wv = unpack_w<K>(wraw, is16);
            double rr[K];
            double c2[K];                                  // w / p^2 (Newton phase only)
            if (fast) {
                // p is finite, normal and positive here, so 1/p and log p are finite: a zero weight gives rr = +0 and leaves
                // ll unchanged without a select (weights are validated w >= 0 at upload; max() guards anyway)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const double wd = u31_to_double(max(wv.v[j], 0));
                    const double rcp = fast_rcp(p[j]);
                    rr[j] = wd * rcp;
                    c2[j] = rr[j] * rcp;
                    const double lg = table_log_pos(p[j], tab_s, a.logc);
                    ll[j] = fma(wd, lg, ll[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const double wd = (double)wv.v[j];
                    const bool on = wv.v[j] > 0;
                    rr[j] = on ? wd / p[j] : 0.0;
                    c2[j] = on ? rr[j] / p[j] : 0.0;
                    ll[j] = on ? fma(wd, log(p[j]), ll[j]) : ll[j];
                }
            }
            if (BACKWARD) sts_lanes<K>(rr_s + (r - row0) * LT, t, rr);
            if (a.cr != nullptr && lv) {
                double* dst = a.cr + (int64_t)r * B + mylane;
                if (K == 1) dst[0] = c2[0];
                else if (K == 2) *reinterpret_cast<double2*>(dst) = make_double2(c2[0], c2[1]);
                else {
                    reinterpret_cast<double2*>(dst)[0] = make_double2(c2[0], c2[1]);
                    reinterpret_cast<double2*>(dst)[1] = make_double2(c2[2], c2[3]);
                }
            }
            if (a.crf != nullptr && lv) {
                float* dst = a.crf + (int64_t)r * B + mylane;
                if (K == 1) dst[0] = (float)c2[0];
                else if (K == 2) *reinterpret_cast<float2*>(dst) = make_float2((float)c2[0], (float)c2[1]);
                else *reinterpret_cast<float4*>(dst) = make_float4((float)c2[0], (float)c2[1], (float)c2[2], (float)c2[3]);
            }
        }
        if (BACKWARD) {
            cp_async_wait_all();
            __syncthreads();                               // rr_s complete, column-major entries landed
            // ------------ phase 2: t += A^T rr over the non-empty columns of the chunk
            const int ncol = cur.ncol;
            const int ce0 = cs_s[0];
            if (treg) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int v = wi + NW * q;
                    if (v < V && ncol > 0) {
                        int ci = min(v, ncol - 1);                  // the chunk lists its non-empty columns in ascending order
                        while (ci > 0 && cid_s[ci] > v) --ci;
                        if (cid_s[ci] == v) {
                            const int k0 = cs_s[ci] - ce0, k1 = cs_s[ci + 1] - ce0;
#pragma unroll 4
                            for (int k = k0; k < k1; ++k) {
                                const unsigned e = c2_s[k];              // chunk-local row | top 16 bits of the fp64 count << 16
                                const double f = ent_count(e);
                                double rv[K];
                                lds_lanes_at<K>(rr_t, e, rv);
#pragma unroll
                                for (int j = 0; j < K; ++j) tacc[q][j] = fma(f, rv[j], tacc[q][j]);
                            }
                        }
                    }
                }
            } else {
                for (int ci = wi; ci < ncol; ci += NW) {
                    const int v = cid_s[ci];
                    const int k0 = cs_s[ci] - ce0, k1 = cs_s[ci + 1] - ce0;
                    double acc[K];
                    lds_lanes<K>(t_s + v * LT, t, acc);
#pragma unroll 4
                    for (int k = k0; k < k1; ++k) {
                        const unsigned e = c2_s[k];
                        const double f = ent_count(e);
                        double rv[K];
                        lds_lanes_at<K>(rr_t, e, rv);
#pragma unroll
                        for (int j = 0; j < K; ++j) acc[j] = fma(f, rv[j], acc[j]);
                    }
                    sts_lanes<K>(t_s + v * LT, t, acc);
                }
            }
        }
        cur = nxt; nxt = nx2;
    }
    cp_async_wait_all();
    __syncthreads();

    // ---------------- epilogue
    if (BACKWARD && treg) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int v = wi + NW * q;
            if (v < V && lv) {
                double* dst = a.tpart + ((int64_t)split * V + v) * B + mylane;
                if (K == 1) dst[0] = tacc[q][0];
                else if (K == 2) *reinterpret_cast<double2*>(dst) = make_double2(tacc[q][0], tacc[q][1]);
                else {
                    reinterpret_cast<double2*>(dst)[0] = make_double2(tacc[q][0], tacc[q][1]);
                    reinterpret_cast<double2*>(dst)[1] = make_double2(tacc[q][2], tacc[q][3]);
                }
            }
        }
    } else if (BACKWARD) {
        for (int i = threadIdx.x; i < V * LT; i += NT) {
            const int v = i / LT, l = i - v * LT;
            const int64_t b = lane0 + l;
            if (b < B) a.tpart[((int64_t)split * V + v) * B + b] = t_s[v * LT + lane_slot<K>(l)];
        }
    }
    double* red = rr_s;                                   // [NW][LT]
    sts_lanes<K>(red + wi * LT, t, ll);
    __syncthreads();
    if (wi == 0) {
        double s[K];
#pragma unroll
        for (int j = 0; j < K; ++j) s[j] = 0.0;
        for (int q = 0; q < NW; ++q) {
            double v[K];
            lds_lanes<K>(red + q * LT, t, v);
#pragma unroll
            for (int j = 0; j < K; ++j) s[j] += v[j];
        }
        if (lv) {
#pragma unroll
            for (int j = 0; j < K; ++j) a.llpart[(int64_t)split * B + mylane + j] = s[j];
        }
    }
}

}  // namespace tvs
