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
//   * the sparse matrix is stored in blocks ("segments") of 240 major x 128 minor indices (forward: major = read path,
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

constexpr int kBlkWarps = 15;                       // consumer warps; one more warp only issues the copies (16 warps: 128 registers each)
constexpr int kBlkThreads = (kBlkWarps + 1) * 32;
constexpr int kBlkRPW = kBlkCH / kBlkWarps;         // major indices (accumulator rows) per warp
static_assert(kBlkWarps == kBlkCW && kBlkRPW == kBlkRW, "host layout (tvs_blk_host.h) and kernel disagree");
__device__ __forceinline__ int64_t blk_major(int c, int wi, int i, int n_chunks, int dealt) {   // = blk_major_of (tvs_blk_host.h)
    return dealt ? ((int64_t)i * kBlkWarps + wi) * n_chunks + c : (int64_t)c * kBlkCH + wi * kBlkRPW + i;
}
constexpr int kBlkMaxStages = 6;

struct BlkMat {                                     // device view of one blocked matrix
    const uint32_t* seg_off;                        // [n_chunks * n_panels + 1] offsets into blob, 16-byte units
    const uint4* blob;
    int n_chunks, n_panels;
    int max_seg_bytes;
    int dealt;                                      // major indices dealt round robin over chunks and warps (blk_major)
};

struct BlkArgs {
    BlkMat m;
    const double* src;                              // dense operand [n_minor x W], lane-minor (x or rr)
    int64_t n_minor, n_major, W;
    const int32_t* done;
    const int32_t* need;                            // backward, optional: lanes with need == 0 keep their previous output (Newton phase:
                                                    // a lane whose trial point was rejected re-uses its Hessian)
    const int32_t* w32; const uint16_t* w16;        // forward: weights [R x W] (one of the two)
    double* rr;                                     // forward out: [R x W] (null: log-likelihood only)
    double* llpart;                                 // forward out: [splits x W]
    double* tpart;                                  // backward out: [splits x V x W]
    int stages;                                     // depth of the shared-memory ring (2..kBlkMaxStages)
    LogCoef logc; const double2* logtab;
};

// shared memory: [stages x operand tile][stages x blob][full barriers][empty barriers][log table]
struct BlkSmem { size_t tile_bytes, blob_bytes, tile0, blob0, bar, tab, total; };
__host__ __device__ inline BlkSmem blk_layout(int LT, int max_seg_bytes, int stages) {
    BlkSmem L;
    L.tile_bytes = (size_t)kBlkPN * LT * 8;
    L.blob_bytes = ((size_t)max_seg_bytes + 127) & ~(size_t)127;
    L.tile0 = 0;
    L.blob0 = (size_t)stages * L.tile_bytes;
    L.bar = L.blob0 + (size_t)stages * L.blob_bytes;
    L.tab = L.bar + 16 * kBlkMaxStages;
    L.total = L.tab + 128 * 16;
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
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// MODE 0: forward (major = read path, operand = x, epilogue w/p + log);  MODE 1: backward (major = variant, operand = rr).
// BULK: operand rows by cp.async.bulk (needs W even: 16-byte aligned rows); otherwise the producer warp copies them with
// 8-byte loads and stores.  K lanes per thread (K = 2 needs W even).
// grid: forward (lane tile, row-chunk split); backward (variant chunk, lane tile, row-panel split).
//
// Warp roles: warps 0..14 consume (each owns 16 major indices of the chunk and runs freely: a slow warp delays nobody
// until the ring is full), warp 15 produces (waits for a stage to be released by all 15 consumers, arms the stage's
// "full" barrier with the byte count and issues the copies).
template <int K, int MODE, bool BULK>
__global__ void __launch_bounds__(kBlkThreads, 1) blk_kernel(const BlkArgs a) {
    constexpr int LT = 32 * K;
    constexpr int ROWB = LT * 8;
    extern __shared__ __align__(128) unsigned char smb[];
    const int tid = threadIdx.x, t = tid & 31, wi = tid >> 5;
    const int64_t W = a.W;
    const int64_t lane0 = (int64_t)(MODE == 0 ? blockIdx.x : blockIdx.y) * LT;
    const int64_t mylane = lane0 + K * t;
    const bool lv = mylane < W;                                    // W % K == 0: all K lanes valid or none
    const int nvalid = (int)((W - lane0 < LT) ? (W - lane0) : LT);
    const int NP = a.m.n_panels;
    const int NST = a.stages;

    if (a.done != nullptr) {                                       // skip tiles whose lanes have all finished
        int alive = 0;
        if (lv && wi < kBlkWarps) {
#pragma unroll
            for (int j = 0; j < K; ++j) alive |= (a.done[mylane + j] == 0);
        }
        if (!__syncthreads_or(alive)) return;
    }

    const BlkSmem L = blk_layout(LT, a.m.max_seg_bytes, NST);
    const unsigned tile_s = smem_u32(smb + L.tile0), blob_s = smem_u32(smb + L.blob0);
    const unsigned full_s = smem_u32(smb + L.bar), empty_s = full_s + 8u * kBlkMaxStages;
    double2* tab_s = reinterpret_cast<double2*>(smb + L.tab);

    // ---- iteration space: (chunk c, panel p) pairs in the order both roles walk them
    int c, p, S;
    if (MODE == 0) { c = blockIdx.y; p = 0; S = gridDim.y; }
    else { c = blockIdx.x; p = blockIdx.z; S = gridDim.z; }
    const int split = (MODE == 0) ? (int)blockIdx.y : (int)blockIdx.z;
    auto valid = [&](int cc, int pp) { return MODE == 0 ? cc < a.m.n_chunks : pp < NP; };
    auto advance = [&](int& cc, int& pp) {
        if (MODE == 0) { if (++pp == NP) { pp = 0; cc += S; } }
        else pp += S;
    };

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full_s + 8u * s, BULK ? 1u : 2u);            // expect_tx arrive (+ the producer's arrive after its stores)
            mbar_init(empty_s + 8u * s, kBlkWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // operand lanes beyond W are never copied: give them a defined value once (generic-proxy writes, then a proxy
    // fence so that the bulk copies are ordered after them)
    for (int i = tid; i < NST * kBlkPN * LT; i += kBlkThreads) reinterpret_cast<double*>(smb + L.tile0)[i] = 1.0;
    if (MODE == 0)
        for (int i = tid; i < 128; i += kBlkThreads) tab_s[i] = a.logtab[i];        // the log table (tvs_create always uploads it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (wi == kBlkWarps) {
        // ================================================================ producer warp
        int st = 0;
        unsigned round = 0;                                        // how many times the ring has wrapped
        for (; valid(c, p); advance(c, p)) {
            const unsigned o0 = a.m.seg_off[(int64_t)c * NP + p], o1 = a.m.seg_off[(int64_t)c * NP + p + 1];
            const unsigned blob_bytes = (o1 - o0) * 16u;
            if (blob_bytes == 0) continue;                         // empty segment: neither role touches the ring
            if (round > 0) mbar_wait(empty_s + 8u * st, (round - 1) & 1u);
            const int64_t minor0 = (int64_t)p * kBlkPN;
            const int nrows = (int)((a.n_minor - minor0 < kBlkPN) ? (a.n_minor - minor0) : kBlkPN);
            const double* src0 = a.src + minor0 * W + lane0;
            const unsigned fb = full_s + 8u * st;
            const unsigned tdst = tile_s + (unsigned)st * (unsigned)L.tile_bytes;
            if (t == 0) {
                mbar_expect_tx(fb, blob_bytes + (BULK ? (unsigned)nrows * (unsigned)nvalid * 8u : 0u));
                bulk_g2s(blob_s + (unsigned)st * (unsigned)L.blob_bytes, a.m.blob + o0, blob_bytes, fb);
            }
            if (BULK) {
                for (int r = t; r < nrows; r += 32) bulk_g2s(tdst + (unsigned)r * ROWB, src0 + (int64_t)r * W, (unsigned)nvalid * 8u, fb);
            } else {
                double* trow = reinterpret_cast<double*>(smb + L.tile0 + (size_t)st * L.tile_bytes);
                if (t < nvalid) {
                    int r = 0;
                    for (; r + 8 <= nrows; r += 8) {
                        double v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = __ldg(src0 + (int64_t)(r + q) * W + t);
#pragma unroll
                        for (int q = 0; q < 8; ++q) trow[(r + q) * LT + t] = v[q];
                    }
                    for (; r < nrows; ++r) trow[r * LT + t] = __ldg(src0 + (int64_t)r * W + t);
                }
                __syncwarp();
                if (t == 0) mbar_arrive(fb);                       // release: the warp's stores above are visible to the waiters
            }
            if (++st == NST) { st = 0; ++round; }
        }
    } else {
        // ================================================================ consumer warps
        double acc[kBlkRPW][K];
#pragma unroll
        for (int i = 0; i < kBlkRPW; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) acc[i][j] = 0.0;
        double ll[K];
#pragma unroll
        for (int j = 0; j < K; ++j) ll[j] = 0.0;
        const unsigned tK = (unsigned)t * (8u * K);                // this thread's bytes inside a tile row (bits 3..8)

        int st = 0;
        unsigned round = 0;
        for (; valid(c, p); advance(c, p)) {
            const unsigned o0 = a.m.seg_off[(int64_t)c * NP + p], o1 = a.m.seg_off[(int64_t)c * NP + p + 1];
            // forward, 16-bit weights: the chunk's weight lines are requested into L2 when its first segment starts (one request
            // per row and warp, no registers), so the elementwise part at the end of the chunk does not wait for DRAM
            // (cfg3, 2,048 lanes: 2.26 -> 2.06 ms per pass; carrying the 16 packed words in registers instead measured the same
            // and spilled; cfg4, 16 panels per chunk: unchanged)
            if (MODE == 0 && p == 0 && a.w16 != nullptr && t == 0) {
#pragma unroll
                for (int i = 0; i < kBlkRPW; ++i) {
                    const int64_t r = blk_major(c, wi, i, a.m.n_chunks, a.m.dealt);
                    if (r < a.n_major) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.w16 + r * W + lane0));
                }
            }
            if (o1 != o0) {
                mbar_wait(full_s + 8u * st, round & 1u);
                // ---- the segment: warp wi owns major indices wi*RPW .. +RPW of the chunk; entry lists have even
                //      lengths, read as 8-byte pairs (one broadcast wavefront per two entries)
                const unsigned char* blob = smb + L.blob0 + (size_t)st * L.blob_bytes;
                const uint16_t* ptr = reinterpret_cast<const uint16_t*>(blob) + wi * kBlkRPW;
                const uint2* ent2 = reinterpret_cast<const uint2*>(blob + kBlkPtrBytes);
                const unsigned tile = tile_s + (unsigned)st * (unsigned)L.tile_bytes;     // shared-window address of the stage's tile
                int k = ptr[0] >> 1;
#pragma unroll
                for (int i = 0; i < kBlkRPW; ++i) {
                    const int k1 = ptr[i + 1] >> 1;
                    double a0 = acc[i][0], a1 = acc[i][K - 1];
#pragma unroll 2
                    for (; k < k1; ++k) {
                        const uint2 e = ent2[k];
                        // operand address: row offset (bits 9..15 of the entry; K = 1: halved) OR-ed with the thread's bytes
                        const unsigned ad0 = (K == 2) ? ((e.x & 0xffffu) | tK) : (((e.x & 0xffffu) >> 1) | tK);
                        const unsigned ad1 = (K == 2) ? ((e.y & 0xffffu) | tK) : (((e.y & 0xffffu) >> 1) | tK);
                        const double f0 = __hiloint2double((int)(e.x & 0xffff0000u), 0);
                        const double f1 = __hiloint2double((int)(e.y & 0xffff0000u), 0);
                        if (K == 2) {
                            const double2 x0 = lds_f64x2(tile + ad0);
                            const double2 x1 = lds_f64x2(tile + ad1);
                            a0 = fma(f0, x0.x, a0); a1 = fma(f0, x0.y, a1);
                            a0 = fma(f1, x1.x, a0); a1 = fma(f1, x1.y, a1);
                        } else {
                            a0 = fma(f0, lds_f64(tile + ad0), a0);
                            a0 = fma(f1, lds_f64(tile + ad1), a0);
                        }
                    }
                    acc[i][0] = a0;
                    if (K == 2) acc[i][K - 1] = a1;
                }
                __syncwarp();
                if (t == 0) mbar_arrive(empty_s + 8u * st);        // this warp is done with the stage
                if (++st == NST) { st = 0; ++round; }
            }
            if (MODE == 0 && p == NP - 1) {
                // ---- forward epilogue of chunk c: rr = w / p, ll += w log p; the accumulators start over
                int wv[kBlkRPW][K];
#pragma unroll
                for (int i = 0; i < kBlkRPW; ++i) {
                    const int64_t r = blk_major(c, wi, i, a.m.n_chunks, a.m.dealt);
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
                    const int64_t r = blk_major(c, wi, i, a.m.n_chunks, a.m.dealt);
                    double rr[K];
                    bool fast = true;
#pragma unroll
                    for (int j = 0; j < K; ++j) fast = fast && is_pos_normal(acc[i][j]);
                    if (fast) {
#pragma unroll
                        for (int j = 0; j < K; ++j) {        // p finite, normal, > 0: a zero weight needs no select (pass3_kernel)
                            const double wd = u31_to_double(max(wv[i][j], 0));
                            const double rcp = fast_rcp(acc[i][j]);
                            rr[j] = wd * rcp;
                            const double lg = table_log_pos(acc[i][j], tab_s, a.logc);
                            ll[j] = fma(wd, lg, ll[j]);
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
            }
        }

        if (MODE == 1) {
            // ---- backward epilogue: the chunk's block of t for this split
            if (lv) {
                bool on[K];
#pragma unroll
                for (int j = 0; j < K; ++j) on[j] = a.need == nullptr || (a.need[mylane + j] != 0 && (a.done == nullptr || a.done[mylane + j] == 0));
#pragma unroll
                for (int i = 0; i < kBlkRPW; ++i) {
                    const int64_t v = blk_major((int)blockIdx.x, wi, i, a.m.n_chunks, a.m.dealt);
                    if (v < a.n_major) {
                        double* dst = a.tpart + ((int64_t)split * a.n_major + v) * W + mylane;
                        if (K == 2 && on[0] && on[K - 1]) *reinterpret_cast<double2*>(dst) = make_double2(acc[i][0], acc[i][K - 1]);
                        else {
                            if (on[0]) dst[0] = acc[i][0];
                            if (K == 2 && on[K - 1]) dst[K - 1] = acc[i][K - 1];
                        }
                    }
                }
            }
        }
        if (MODE == 0) {
            // ---- forward: log-likelihood partial of this split, summed over the warps in fixed order.  Every consumer has
            // waited for all of its segments, so no copy is in flight and the first tile can be reused; named barrier over
            // the 15 consumer warps (the producer warp does not take part).
            asm volatile("bar.sync 1, %0;" ::"n"(kBlkWarps * 32) : "memory");
            double* red = reinterpret_cast<double*>(smb + L.tile0);  // [warps][LT]
#pragma unroll
            for (int j = 0; j < K; ++j) red[wi * LT + K * t + j] = ll[j];
            asm volatile("bar.sync 1, %0;" ::"n"(kBlkWarps * 32) : "memory");
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
}

}  // namespace tvs
