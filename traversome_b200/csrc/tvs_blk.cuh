// blk_kernel -- likelihood pass for WIDE candidate sets (V above what pass3_kernel can keep in shared memory:
// BASELINE configs[3], 2,048 variants x 200k read paths).  Included by tvs_api.cu.
//
//     forward :  p = A x ;  rr = w / p ;  ll += w log p        (rr written to global memory)
//     backward:  t = A^T rr                                    (per-split partials)
//
// replaces: PathMultinomialModel.get_like_formula (ModelGenerator.py:153-174) evaluated by the lambdified objective
// (ModelFitMaxLike.py:534-545), and its gradient (finite differences in the reference, ModelFitMaxLike.py:36-40).
//
// pass_kernel (round 1) gathered x from L2 for every non-zero and scattered t through L2: 197 GB of L2 traffic per
// 625-lane pass, 24.9 ms.  Here both products are the SAME kernel -- a sparse "major x minor" matrix times a dense
// [minor x lanes] operand with the accumulators of the major index in REGISTERS:
//   * the sparse matrix is stored in blocks ("segments") of 256 major x 128 minor indices (forward: major = read path,
//     minor = variant; backward: the transpose).  A segment is one contiguous blob: u16 entry pointers per major
//     index, then packed 32-bit entries (byte offset of the operand row inside the tile | top 16 bits of the fp64
//     count, so that unpacking needs no fp64 instruction);
//   * a CTA of 16 warps owns 64 lanes (2 per thread) and walks segments; per segment the 128-row operand tile (64 KB)
//     and the blob are fetched by bulk asynchronous copies (cp.async.bulk -> mbarrier complete_tx, TMA unit) into a
//     two-stage ring while the previous segment is being consumed;
//   * every warp owns 16 major indices of the chunk, i.e. 32 fp64 accumulators per thread that live in registers
//     across all segments of the chunk: no partial sums in shared or global memory, each operand element that enters
//     shared memory is used 256 x density (~13) times;
//   * forward epilogue per 256-row chunk: w/p, log p (same routines as pass3_kernel), rr -> global;  backward epilogue:
//     the chunk's 256 x 64 block of t -> per-split partial, summed in fixed order by the update kernel.
// Bound: one 8-byte shared-memory operand per DFMA (128 B/clk/SM), as in pass3_kernel; L2 -> shared fill traffic is
// nnz * lanes * 8 / 13 bytes per product.
#pragma once
#include "tvs_internal.h"
#include "tvs_pass2.cuh"
#include "tvs_blk_host.h"

namespace tvs {

constexpr int kBlkWarps = 16;
constexpr int kBlkThreads = kBlkWarps * 32;
constexpr int kBlkRPW = kBlkCH / kBlkWarps;         // major indices (accumulator rows) per warp

struct BlkMat {                                     // device view of one blocked matrix
    const uint32_t* seg_off;                        // [n_chunks * n_panels + 1] offsets into blob, 16-byte units
    const uint4* blob;
    int n_chunks, n_panels;
    int max_seg_bytes;
};

struct BlkArgs {
    BlkMat m;
    const double* src;                              // dense operand [n_minor x W], lane-minor (x or rr)
    int64_t n_minor, n_major, W;
    const int32_t* done;
    const int32_t* w32; const uint16_t* w16;        // forward: weights [R x W] (one of the two)
    double* rr;                                     // forward out: [R x W]
    double* llpart;                                 // forward out: [splits x W]
    double* tpart;                                  // backward out: [splits x V x W]
    int splits;                                     // backward: row-panel splits (grid.y = chunks * splits)
    LogCoef logc; const double2* logtab;
};

struct BlkSmem { size_t tile[2], blob[2], bar, tab, total; };
__host__ __device__ inline BlkSmem blk_layout(int LT, int max_seg_bytes) {
    BlkSmem L;
    size_t o = 0;
    const size_t tile = (size_t)kBlkPN * LT * 8;
    const size_t blob = ((size_t)max_seg_bytes + 127) & ~(size_t)127;
    L.tile[0] = o; o += tile; L.tile[1] = o; o += tile;
    L.blob[0] = o; o += blob; L.blob[1] = o; o += blob;
    L.bar = o; o += 64;
    L.tab = o; o += 128 * 16;
    L.total = o;
    return L;
}

// ---- mbarrier / bulk-copy primitives (PTX ISA: cp.async.bulk, mbarrier; SASS: UBLKCP, SYNCS)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// MODE 0: forward (major = read path, operand = x, epilogue w/p + log);  MODE 1: backward (major = variant, operand = rr).
// BULK: operand rows by cp.async.bulk (needs W even: 16-byte aligned rows); otherwise 8-byte cp.async by all threads.
// K lanes per thread (K = 2 needs W even).
template <int K, int MODE, bool BULK>
__global__ void __launch_bounds__(kBlkThreads, 1) blk_kernel(const BlkArgs a) {
    constexpr int LT = 32 * K;
    constexpr int ROWB = LT * 8;
    extern __shared__ __align__(128) unsigned char smb[];
    const int tid = threadIdx.x, t = tid & 31, wi = tid >> 5;
    const int64_t W = a.W;
    const int64_t lane0 = (int64_t)blockIdx.x * LT;
    const int64_t mylane = lane0 + K * t;
    const bool lv = mylane < W;                                    // W % K == 0: all K lanes valid or none
    const int nvalid = (int)((W - lane0 < LT) ? (W - lane0) : LT);
    const int NP = a.m.n_panels;

    if (a.done != nullptr) {                                       // skip tiles whose lanes have all finished
        int alive = 0;
        if (lv) {
#pragma unroll
            for (int j = 0; j < K; ++j) alive |= (a.done[mylane + j] == 0);
        }
        if (!__syncthreads_or(alive)) return;
    }

    const BlkSmem L = blk_layout(LT, a.m.max_seg_bytes);
    const size_t tile_bytes = L.tile[1] - L.tile[0], blob_bytes_cap = L.blob[1] - L.blob[0];
    auto tile_of = [&](int st) { return smb + L.tile[0] + (size_t)st * tile_bytes; };
    auto blob_of = [&](int st) { return smb + L.blob[0] + (size_t)st * blob_bytes_cap; };
    auto bar_of = [&](int st) { return smem_u32(smb + L.bar) + 8u * (unsigned)st; };
    double2* tab_s = reinterpret_cast<double2*>(smb + L.tab);

    // ---- iteration space
    int c, p, S;
    if (MODE == 0) { c = blockIdx.y; p = 0; S = gridDim.y; }
    else { c = blockIdx.y / a.splits; p = blockIdx.y - c * a.splits; S = a.splits; }
    const int split = (MODE == 0) ? (int)blockIdx.y : p;
    auto valid = [&](int cc, int pp) { return MODE == 0 ? cc < a.m.n_chunks : pp < NP; };
    auto advance = [&](int& cc, int& pp) {
        if (MODE == 0) { if (++pp == NP) { pp = 0; cc += S; } }
        else pp += S;
    };

    if (tid == 0) {
        mbar_init(bar_of(0), 1); mbar_init(bar_of(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // operand lanes beyond W are never copied: give them a defined value once (generic-proxy writes, then a proxy
    // fence so that the bulk copies below are ordered after them)
    for (int i = tid; i < 2 * kBlkPN * LT; i += kBlkThreads) reinterpret_cast<double*>(smb)[i] = 1.0;
    if (MODE == 0 && a.logtab != nullptr)
        for (int i = tid; i < 128; i += kBlkThreads) tab_s[i] = a.logtab[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // copies of one segment into stage `st`; returns nothing (uniform control flow)
    auto issue = [&](int st, int pp, unsigned o0, unsigned o1) {
        const unsigned blob_bytes = (o1 - o0) * 16u;
        if (blob_bytes == 0) return;
        const int64_t minor0 = (int64_t)pp * kBlkPN;
        const int nrows = (int)((a.n_minor - minor0 < kBlkPN) ? (a.n_minor - minor0) : kBlkPN);
        const double* src0 = a.src + minor0 * W + lane0;
        if (BULK) {
            if (wi == 0) {
                if (t == 0) {
                    mbar_expect_tx(bar_of(st), blob_bytes + (unsigned)nrows * (unsigned)nvalid * 8u);
                    bulk_g2s(smem_u32(blob_of(st)), a.m.blob + o0, blob_bytes, bar_of(st));
                }
                for (int r = t; r < nrows; r += 32)
                    bulk_g2s(smem_u32(tile_of(st) + (size_t)r * ROWB), src0 + (int64_t)r * W, (unsigned)nvalid * 8u, bar_of(st));
            }
        } else {
            if (tid == 0) {
                mbar_expect_tx(bar_of(st), blob_bytes);
                bulk_g2s(smem_u32(blob_of(st)), a.m.blob + o0, blob_bytes, bar_of(st));
            }
            for (int i = tid; i < nrows * LT; i += kBlkThreads) {
                const int r = i / LT, l = i - r * LT;
                if (l < nvalid) cp_async8(tile_of(st) + (size_t)r * ROWB + (size_t)l * 8, src0 + (int64_t)r * W + l);
            }
        }
    };

    double acc[kBlkRPW][K];
#pragma unroll
    for (int i = 0; i < kBlkRPW; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) acc[i][j] = 0.0;
    double ll[K];
#pragma unroll
    for (int j = 0; j < K; ++j) ll[j] = 0.0;

    const bool use_tab = a.logtab != nullptr;
    // forward epilogue of chunk cc: rr = w / p, ll += w log p
    auto epilogue_fwd = [&](int cc) {
        const int64_t r0 = (int64_t)cc * kBlkCH + wi * kBlkRPW;
        int wv[kBlkRPW][K];
#pragma unroll
        for (int i = 0; i < kBlkRPW; ++i) {
            const int64_t r = r0 + i;
#pragma unroll
            for (int j = 0; j < K; ++j) wv[i][j] = 0;
            if (lv && r < a.n_major) {
                if (a.w16 != nullptr) {
                    if (K == 2) { const unsigned u = *reinterpret_cast<const unsigned*>(a.w16 + r * W + mylane); wv[i][0] = (int)(u & 0xffffu); wv[i][K - 1] = (int)(u >> 16); }
                    else wv[i][0] = a.w16[r * W + mylane];
                } else {
                    if (K == 2) { const int2 u = *reinterpret_cast<const int2*>(a.w32 + r * W + mylane); wv[i][0] = u.x; wv[i][K - 1] = u.y; }
                    else wv[i][0] = a.w32[r * W + mylane];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kBlkRPW; ++i) {
            const int64_t r = r0 + i;
            double rr[K];
            bool fast = true;
#pragma unroll
            for (int j = 0; j < K; ++j) fast = fast && is_pos_normal(acc[i][j]);
            if (fast) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const bool on = wv[i][j] > 0;
                    const double wd = u31_to_double(on ? wv[i][j] : 0);
                    const double rcp = fast_rcp(acc[i][j]);
                    rr[j] = on ? wd * rcp : 0.0;
                    const double lg = use_tab ? table_log_pos(acc[i][j], tab_s, a.logc) : fast_log_pos(acc[i][j], a.logc);
                    ll[j] = on ? fma(wd, lg, ll[j]) : ll[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const double wd = (double)wv[i][j];
                    const bool on = wv[i][j] > 0;
                    rr[j] = on ? wd / acc[i][j] : 0.0;
                    ll[j] = on ? fma(wd, log(acc[i][j]), ll[j]) : ll[j];
                }
            }
            if (a.rr != nullptr && lv && r < a.n_major) {
                if (K == 2) *reinterpret_cast<double2*>(a.rr + r * W + mylane) = make_double2(rr[0], rr[K - 1]);
                else a.rr[r * W + mylane] = rr[0];
            }
#pragma unroll
            for (int j = 0; j < K; ++j) acc[i][j] = 0.0;
        }
    };

    if (valid(c, p)) {
        unsigned o0 = a.m.seg_off[(int64_t)c * NP + p], o1 = a.m.seg_off[(int64_t)c * NP + p + 1];
        issue(0, p, o0, o1);
        if (!BULK) cp_async_commit();
        int cn = c, pn = p;
        advance(cn, pn);
        unsigned n0 = 0, n1 = 0;
        if (valid(cn, pn)) { n0 = a.m.seg_off[(int64_t)cn * NP + pn]; n1 = a.m.seg_off[(int64_t)cn * NP + pn + 1]; }
        unsigned phase = 0u;                                    // bit st = parity of stage st
        for (int it = 0; valid(c, p); ++it) {
            const int st = it & 1;
            __syncthreads();                                       // iteration it-1 consumed: stage st^1 is free
            const bool more = valid(cn, pn);
            if (more) issue(st ^ 1, pn, n0, n1);
            if (!BULK) cp_async_commit();
            int c2 = cn, p2 = pn;
            unsigned m0 = 0, m1 = 0;
            if (more) {
                advance(c2, p2);
                if (valid(c2, p2)) { m0 = a.m.seg_off[(int64_t)c2 * NP + p2]; m1 = a.m.seg_off[(int64_t)c2 * NP + p2 + 1]; }
            }
            if (o1 != o0) {
                if (!BULK) { asm volatile("cp.async.wait_group 1;" ::: "memory"); __syncthreads(); }
                mbar_wait(bar_of(st), (phase >> st) & 1u);
                phase ^= 1u << st;
                // ---- the segment: warp wi owns major indices wi*RPW .. +RPW of the chunk
                const uint16_t* ptr = reinterpret_cast<const uint16_t*>(blob_of(st)) + wi * kBlkRPW;
                const unsigned* ent = reinterpret_cast<const unsigned*>(blob_of(st) + kBlkPtrBytes);
                const unsigned char* xb = tile_of(st) + t * (8 * K);
                int k = ptr[0];
#pragma unroll
                for (int i = 0; i < kBlkRPW; ++i) {
                    const int k1 = ptr[i + 1];
                    double a0 = acc[i][0], a1 = acc[i][K - 1];
#pragma unroll 4
                    for (; k < k1; ++k) {
                        const unsigned e = ent[k];
                        const double f = __hiloint2double((int)(e & 0xffff0000u), 0);
                        const unsigned off = (K == 2) ? (e & 0xffffu) : ((e & 0xffffu) >> 1);
                        if (K == 2) {
                            const double2 xv = *reinterpret_cast<const double2*>(xb + off);
                            a0 = fma(f, xv.x, a0); a1 = fma(f, xv.y, a1);
                        } else {
                            a0 = fma(f, *reinterpret_cast<const double*>(xb + off), a0);
                        }
                    }
                    acc[i][0] = a0;
                    if (K == 2) acc[i][K - 1] = a1;
                }
            }
            if (MODE == 0 && p == NP - 1) epilogue_fwd(c);
            c = cn; p = pn; o0 = n0; o1 = n1;
            cn = c2; pn = p2; n0 = m0; n1 = m1;
        }
    }
    if (!BULK) asm volatile("cp.async.wait_group 0;" ::: "memory");

    if (MODE == 1) {
        // ---- backward epilogue: the chunk's block of t for this split
        const int cc = blockIdx.y / a.splits;
        const int64_t v0 = (int64_t)cc * kBlkCH + wi * kBlkRPW;
        if (lv) {
#pragma unroll
            for (int i = 0; i < kBlkRPW; ++i) {
                const int64_t v = v0 + i;
                if (v < a.n_major) {
                    double* dst = a.tpart + ((int64_t)split * a.n_major + v) * W + mylane;
                    if (K == 2) *reinterpret_cast<double2*>(dst) = make_double2(acc[i][0], acc[i][K - 1]);
                    else dst[0] = acc[i][0];
                }
            }
        }
    } else {
        // ---- forward: log-likelihood partial of this split (sum over the warps in fixed order)
        __syncthreads();
        double* red = reinterpret_cast<double*>(smb);              // [warps][LT]
#pragma unroll
        for (int j = 0; j < K; ++j) red[wi * LT + K * t + j] = ll[j];
        __syncthreads();
        if (wi == 0 && lv) {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                double s = 0.0;
                for (int q = 0; q < kBlkWarps; ++q) s += red[q * LT + K * t + j];
                a.llpart[(int64_t)split * W + mylane + j] = s;
            }
        }
    }
}

}  // namespace tvs
