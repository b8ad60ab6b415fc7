// libtvs.so -- C ABI (include/tvs.h) over the sm_100a kernels in tvs_kernels.cuh.
// Host side: CSR validation and chunking, launch geometry, the batched L-BFGS driver loop.
#include "tvs_kernels.cuh"
#include "tvs_pass2.cuh"
#include "tvs_newton.cuh"
#include "tvs_blk.cuh"

#include <cublas_v2.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace tvs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? TVS_ENOMEM : TVS_ECUDA;
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

template <class T>
static int dev_alloc(T** p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    TVS_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    return TVS_OK;
}
template <class T>
static void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

static void free_csr(Csr& c) {
    dev_free(c.row_ptr); dev_free(c.ent); dev_free(c.chunk_row0); dev_free(c.ccol_ptr); dev_free(c.ccol_id);
    dev_free(c.ccol_start); dev_free(c.cent); dev_free(c.ent3); dev_free(c.cent3); dev_free(c.invL); dev_free(c.Ld);
    if (c.desc) { cudaFree(c.desc); c.desc = nullptr; }
    if (c.logtab) { cudaFree(c.logtab); c.logtab = nullptr; }
    dev_free(c.slab_ptr); dev_free(c.h_row); dev_free(c.h_val); dev_free(c.h_slot);
    dev_free(c.wide_ptr); if (c.wide_rec) { cudaFree(c.wide_rec); c.wide_rec = nullptr; }
    dev_free(c.dense);
    for (Csr::Blk* b : {&c.blk_f, &c.blk_b, &c.blk_h}) {
        dev_free(b->seg_off);
        if (b->blob) { cudaFree(b->blob); b->blob = nullptr; }
    }
    c.has_blk = false; c.has_blk_h = false;
}
static void free_lanes(Lanes& l) {
    dev_free(l.w); dev_free(l.mask); dev_free(l.C); dev_free(l.N); dev_free(l.w16); dev_free(l.flag); dev_free(l.cols); dev_free(l.cols16); dev_free(l.slot);
    if (l.acc) cudaFree(l.acc);
    l = Lanes();
}
static void free_fit(FitState& f) {
    for (void* p : f.owned) cudaFree(p);
    for (int i = 0; i < 2; ++i) { if (f.wbuf[i]) cudaFree(f.wbuf[i]); if (f.wbuf16[i]) cudaFree(f.wbuf16[i]); }
    if (f.h_nactive) cudaFreeHost(f.h_nactive);
    if (f.Hpart) cudaFree(f.Hpart);
    if (f.pad_w) cudaFree(f.pad_w);
    if (f.pad_mask) cudaFree(f.pad_mask);
    if (f.pad_N) cudaFree(f.pad_N);
    if (f.pad_C) cudaFree(f.pad_C);
    if (f.cr) cudaFree(f.cr);
    if (f.crf) cudaFree(f.crf);
    f = FitState();
}

// ---------------------------------------------------------------------------------- launch geometry
struct PassPlan {
    int lpt = 1; bool xs = false; int tiles = 0; int splits = 1; size_t smem = 0;
    bool dense = false;   // two cuBLAS DGEMMs on the dense copy of A + dense_mid_kernel (dense shapes, see launch_dense)
    bool v2 = false;      // pass2_kernel<lpt, BW> (lpt = lanes per thread)
    int v3_warps = 0;     // > 0: pass3_kernel<lpt, v3_warps, BW> (entries staged in shared memory)
    bool blk = false;     // blk_kernel<lpt, ., bulk>: forward (splits = row-chunk splits) + backward (bsplits row-panel splits)
    bool bulk = false;
    int bsplits = 1; size_t bsmem = 0; int stages = 2, bstages = 2;
};

static int splits_for(const tvs_problem* h, int slots, int tiles) {
    const Csr& c = h->csr;
    int s;
    if (tiles <= slots) {
        s = slots / std::max(tiles, 1);                        // one wave: every resident slot gets a CTA
    } else {
        // more lane tiles than resident slots: pick the split count whose CTA total fills whole waves best
        // (e.g. 79 tiles on 148 slots: 1 split = 53 % of one wave, 15 splits = 1185 CTAs = 8.0 waves)
        const int s_max = std::max(1, std::min(c.n_chunks / 4, 32));
        double best = -1.0;
        s = 1;
        for (int cand = 1; cand <= s_max; ++cand) {
            const int64_t total = (int64_t)tiles * cand;
            const int64_t waves = (total + slots - 1) / slots;
            const double eff = (double)total / (double)(waves * slots);
            if (eff > best + 0.02) { best = eff; s = cand; }    // prefer fewer splits unless clearly better
        }
    }
    s = std::max(1, std::min(s, c.n_chunks));
    s = std::min(s, env_int("TVS_MAX_SPLITS", 1024));
    const int forced = env_int("TVS_SPLITS", 0);
    if (forced > 0) s = std::min(forced, c.n_chunks);
    return s;
}

// pass2_kernel: K lanes per thread, x/t tiles in shared memory.  Fails with TVS_ELIMIT when the tiles do not fit.
template <int K, bool BW>
static int plan_v2(const tvs_problem* h, int64_t B, PassPlan* out) {
    const int LT = 32 * K;
    const Csr& c = h->csr;
    if (B % K != 0) return TVS_ELIMIT;
    const size_t dbl = (size_t)c.V * LT * (BW ? 2 : 1) + (size_t)std::max(c.rows_per_chunk, kWarps2) * LT;
    const size_t smem = dbl * sizeof(double);
    if (smem > 220 * 1024) return TVS_ELIMIT;
    auto kern = pass2_kernel<K, BW>;
    TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    TVS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads2, smem));
    if (occ < 1) return TVS_ELIMIT;
    out->v2 = true; out->lpt = K; out->xs = true; out->smem = smem;
    out->tiles = (int)((B + LT - 1) / LT);
    out->splits = splits_for(h, occ * h->sm_count, out->tiles);
    return TVS_OK;
}

// pass3_kernel: pass2 + cp.async staging of each chunk's entries (needs max_chunk_nnz * 12 bytes more).
template <int K, int NW, bool BW>
static int plan_v3(const tvs_problem* h, int64_t B, PassPlan* out) {
    const int LT = 32 * K;
    const Csr& c = h->csr;
    if (B % K != 0 || c.rows_per_chunk < NW) return TVS_ELIMIT;
    if (!c.ent3_ok) return TVS_ELIMIT;                 // a count that is not exact in 16 bits of an fp64: pass2_kernel
    const Pass3Smem L = pass3_layout((int)c.V, LT, c.rows_per_chunk, std::max(c.max_chunk_nnz, 1), BW, NW);
    if (L.total > 225 * 1024) return TVS_ELIMIT;
    auto kern = pass3_kernel<K, NW, BW>;
    TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    int occ = 0;
    TVS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, L.total));
    if (occ < 1) return TVS_ELIMIT;
    out->v2 = true; out->v3_warps = NW; out->lpt = K; out->xs = true; out->smem = L.total;
    out->tiles = (int)((B + LT - 1) / LT);
    out->splits = splits_for(h, occ * h->sm_count, out->tiles);
    return TVS_OK;
}

template <bool BW>
static int make_plan_v2(const tvs_problem* h, int64_t B, PassPlan* out) {
    // cheapest K by (padded lanes) x (relative cost per lane of the K-lane inner loops)
    const double rel[3] = {1.0, 0.70, 0.55};
    const int forced = env_int("TVS_K", 0);
    double best = 1e300;
    PassPlan bp;
    int rc_any = TVS_ELIMIT;
    for (int ki = 0; ki < 3; ++ki) {
        const int K = 1 << ki;
        if (forced && forced != K) continue;
        PassPlan p;
        int rc = TVS_ELIMIT;
        if (env_int("TVS_NO_STAGE", 0) == 0) {
            if (env_int("TVS_NW", 16) == 32)
                rc = (K == 1) ? plan_v3<1, 32, BW>(h, B, &p) : (K == 2) ? plan_v3<2, 32, BW>(h, B, &p) : plan_v3<4, 32, BW>(h, B, &p);
            else
                rc = (K == 1) ? plan_v3<1, 16, BW>(h, B, &p) : (K == 2) ? plan_v3<2, 16, BW>(h, B, &p) : plan_v3<4, 16, BW>(h, B, &p);
            if (rc == TVS_ECUDA) return rc;
        }
        if (rc != TVS_OK) rc = (K == 1) ? plan_v2<1, BW>(h, B, &p) : (K == 2) ? plan_v2<2, BW>(h, B, &p) : plan_v2<4, BW>(h, B, &p);
        if (rc == TVS_ECUDA) return rc;
        if (rc != TVS_OK) continue;
        rc_any = TVS_OK;
        const double cost = (double)p.tiles * 32 * K * rel[ki];
        if (cost < best) { best = cost; bp = p; }
    }
    if (rc_any == TVS_OK) *out = bp;
    return rc_any;
}

template <int LPT, bool XS, bool BW>
static int plan_one(const tvs_problem* h, int64_t B, PassPlan* out) {
    const int LT = 32 * LPT;
    const Csr& c = h->csr;
    size_t dbl = 0;
    if (XS) dbl += (size_t)c.V * LT * (BW ? 2 : 1);
    dbl += (size_t)std::max(c.rows_per_chunk, kWarps) * LT;
    const size_t smem = dbl * sizeof(double);
    if (smem > 200 * 1024) return TVS_ELIMIT;
    auto kern = pass_kernel<LPT, XS, BW>;
    TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    TVS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    if (occ < 1) return TVS_ELIMIT;
    const int slots = occ * h->sm_count;
    out->lpt = LPT; out->xs = XS; out->smem = smem;
    out->tiles = (int)((B + LT - 1) / LT);
    out->splits = splits_for(h, slots, out->tiles);
    return TVS_OK;
}


static int upload_blk(const BlkHost& hb, Csr::Blk* d) {
    int rc;
    if ((rc = dev_alloc(&d->seg_off, hb.seg_off.size())) != TVS_OK) return rc;
    uint32_t* blob = nullptr;
    if ((rc = dev_alloc(&blob, hb.blob.size())) != TVS_OK) return rc;
    d->blob = blob;
    TVS_CUDA(cudaMemcpy(d->seg_off, hb.seg_off.data(), hb.seg_off.size() * 4, cudaMemcpyHostToDevice));
    TVS_CUDA(cudaMemcpy(blob, hb.blob.data(), hb.blob.size() * 4, cudaMemcpyHostToDevice));
    d->n_chunks = hb.n_chunks; d->n_panels = hb.n_panels; d->max_seg_bytes = hb.max_seg_bytes; d->dealt = hb.dealt ? 1 : 0;
    return TVS_OK;
}

// blk_kernel geometry for a lane set of width W: 2 lanes per thread and bulk operand copies when W is even, else 1 lane
// per thread with 8-byte cp.async rows.  Fails with TVS_ELIMIT when a segment does not fit the shared-memory ring.
template <int K, int MODE, bool BULK>
static int blk_prepare(const tvs_problem* h, size_t smem, int* occ) {
    auto kern = blk_kernel<K, MODE, BULK>;
    TVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TVS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, kBlkThreads, smem));
    (void)h;
    return TVS_OK;
}

template <bool BW>
static int plan_blk(const tvs_problem* h, int64_t W, PassPlan* out) {
    const Csr& c = h->csr;
    if (!c.has_blk) return TVS_ELIMIT;
    const size_t limit = 227 * 1024;
    int K = (W % 2 == 0) ? 2 : 1;
    const bool bulk = (W % 2 == 0);
    const int forcedK = env_int("TVS_BLK_K", 0);
    if (forcedK == 1) K = 1;
    // deepest ring (2..kBlkMaxStages stages of operand tile + blob) that fits the shared memory of one CTA
    auto stages_for = [&](int k, int seg_bytes) {
        int n = std::min(kBlkMaxStages, std::max(2, env_int("TVS_BLK_STAGES", kBlkMaxStages)));
        while (n >= 2 && blk_layout(32 * k, seg_bytes, n).total > limit) --n;
        return n;
    };
    int nf = stages_for(K, c.blk_f.max_seg_bytes), nb = stages_for(K, c.blk_b.max_seg_bytes);
    if (K == 2 && (nf < 2 || (BW && nb < 2))) {
        K = 1;
        nf = stages_for(1, c.blk_f.max_seg_bytes); nb = stages_for(1, c.blk_b.max_seg_bytes);
    }
    if (nf < 2 || (BW && nb < 2)) return TVS_ELIMIT;
    const size_t sf = blk_layout(32 * K, c.blk_f.max_seg_bytes, nf).total, sb = blk_layout(32 * K, c.blk_b.max_seg_bytes, nb).total;
    int occ_f = 0, occ_b = 0, rc;
    if (K == 2) {
        if ((rc = blk_prepare<2, 0, true>(h, sf, &occ_f)) != TVS_OK) return rc;
        if (BW && (rc = blk_prepare<2, 1, true>(h, sb, &occ_b)) != TVS_OK) return rc;
    } else if (bulk) {
        if ((rc = blk_prepare<1, 0, true>(h, sf, &occ_f)) != TVS_OK) return rc;
        if (BW && (rc = blk_prepare<1, 1, true>(h, sb, &occ_b)) != TVS_OK) return rc;
    } else {
        if ((rc = blk_prepare<1, 0, false>(h, sf, &occ_f)) != TVS_OK) return rc;
        if (BW && (rc = blk_prepare<1, 1, false>(h, sb, &occ_b)) != TVS_OK) return rc;
    }
    if (occ_f < 1 || (BW && occ_b < 1)) return TVS_ELIMIT;
    *out = PassPlan();
    out->blk = true; out->bulk = bulk; out->lpt = K; out->xs = true; out->smem = sf; out->bsmem = sb;
    out->stages = nf; out->bstages = nb;
    const int LT = 32 * K;
    out->tiles = (int)((W + LT - 1) / LT);
    const int slots = h->sm_count;                                 // one CTA per SM
    // split count whose CTA total (units * s) fills whole waves of `slots` CTAs best; fewer splits unless clearly better
    auto fill_waves = [&](int units, int s_max) {
        int s = 1;
        double best = -1.0;
        for (int cand = 1; cand <= std::max(1, s_max); ++cand) {
            const int64_t total = (int64_t)units * cand;
            const int64_t waves = (total + slots - 1) / slots;
            const double eff = (double)total / (double)(waves * slots);
            if (eff > best + 0.02) { best = eff; s = cand; }
        }
        return s;
    };
    // forward: CTA (tile, split) walks the row chunks split, split + S, ...; every split writes one ll partial
    {
        int s = fill_waves(out->tiles, std::min(2 * slots, std::max(1, c.blk_f.n_chunks / 2)));
        const int forced = env_int("TVS_SPLITS", 0);
        if (forced > 0) s = forced;
        out->splits = std::max(1, std::min(s, c.blk_f.n_chunks));
    }
    // backward: CTA (tile, variant chunk, split) walks the row panels split, split + S, ...; tpart[split][V][W]
    if (BW) {
        int s = fill_waves(out->tiles * c.blk_b.n_chunks, std::min(std::max(24, slots / std::max(1, out->tiles * c.blk_b.n_chunks)), std::max(1, c.blk_b.n_panels / 4)));
        const int forced = env_int("TVS_BSPLITS", 0);
        if (forced > 0) s = forced;
        out->bsplits = std::max(1, std::min(s, c.blk_b.n_panels));
    }
    return TVS_OK;
}

template <bool BW>
static int make_plan(const tvs_problem* h, int64_t B, PassPlan* out) {
    if (h->csr.dense != nullptr && env_int("TVS_DENSE", 1) != 0) {
        *out = PassPlan();
        out->dense = true; out->tiles = 1; out->splits = 1;
        return TVS_OK;
    }
    const int want_blk = env_int("TVS_BLK", -1);                  // 1: prefer blk_kernel whenever it is built, 0: never
    if (want_blk == 1) {
        const int rcb = plan_blk<BW>(h, B, out);
        if (rcb == TVS_OK || rcb == TVS_ECUDA) return rcb;
    }
    if (env_int("TVS_PASS_V1", 0) == 0) {
        const int rc2 = make_plan_v2<BW>(h, B, out);
        if (rc2 == TVS_ECUDA) return rc2;
        if (rc2 == TVS_OK) {
            // staged kernel with one lane per thread only (its x / t tiles leave no room for more): on a wide lane set the
            // blocked kernel wins (it keeps no whole-V tile); narrow sets are latency-bound and stay on the staged kernel
            if (want_blk != 0 && out->lpt == 1 && h->csr.has_blk && B >= env_int("TVS_BLK_MIN_LANES", 192)) {
                PassPlan pb;
                if (plan_blk<BW>(h, B, &pb) == TVS_OK) *out = pb;
            }
            return TVS_OK;
        }                                                        // otherwise: tiles too large for shared memory
    }
    if (want_blk != 0) {
        const int rcb = plan_blk<BW>(h, B, out);
        if (rcb == TVS_OK || rcb == TVS_ECUDA) return rcb;       // otherwise: a segment too large for the ring
    }
    int lpt = (B > 32) ? 2 : 1;
    const int forced = env_int("TVS_LPT", 0);
    if (forced == 1 || forced == 2) lpt = forced;
    const bool no_xs = env_int("TVS_NO_XS", 0) != 0;
    int rc = TVS_ELIMIT;
    if (!no_xs) {
        if (lpt == 2) rc = plan_one<2, true, BW>(h, B, out);
        if (rc != TVS_OK) rc = plan_one<1, true, BW>(h, B, out);
    }
    if (rc != TVS_OK && lpt == 2) rc = plan_one<2, false, BW>(h, B, out);
    if (rc != TVS_OK) rc = plan_one<1, false, BW>(h, B, out);
    if (rc != TVS_OK && rc != TVS_ECUDA) set_error("no pass-kernel configuration fits (V=%lld)", (long long)h->csr.V);
    return rc;
}

// ---- dense pass: for shapes whose compatibility matrix is dense enough that a full fp64 contraction beats the sparse
// kernels (BASELINE configs[4]: 4,096 variants at 30 % density; the sparse pass there runs at ~12 % of the fp64 FMA
// peak because every non-zero gathers its operand from L2, a DGEMM runs near the fp64 tensor peak and wastes 70 %).
//   P^T [W x R] = X^T [W x V] * A^T [V x R]      (lane-minor arrays are column-major matrices with ld = W)
//   rr = w / p, ll += w log p                     (dense_mid_kernel, fixed summation order)
//   T^T [W x V] = RR^T [W x R] * A [R x V]
__global__ void dense_mid_kernel(double* __restrict__ P, const int32_t* __restrict__ w, const uint16_t* __restrict__ w16, const int32_t* __restrict__ done,
                                 int64_t W, int64_t R, int64_t rows_per_split, double* __restrict__ llpart, int backward) {
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= W) return;
    const bool live = done == nullptr || done[lane] == 0;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_split, r1 = min(R, r0 + rows_per_split);
    double ll = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t i = r * W + lane;
        const int32_t wv = w16 != nullptr ? (int32_t)w16[i] : w[i];
        double rr = 0.0;
        if (live && wv != 0) {
            const double p = P[i];
            const double wd = (double)wv;
            ll = fma(wd, log(p), ll);
            rr = wd / p;
        }
        if (backward) P[i] = rr;
    }
    llpart[(int64_t)blockIdx.y * W + lane] = ll;
}

__global__ void dense_ll_reduce_kernel(const double* __restrict__ part, int S, int64_t W, double* __restrict__ out) {
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= W) return;
    double s = 0.0;
    for (int q = 0; q < S; ++q) s += part[(int64_t)q * W + lane];
    out[lane] = s;
}

static int launch_dense(tvs_problem* h, const PassArgs& a, bool backward) {
    const Csr& c = h->csr;
    const int64_t W = a.B, R = c.R, V = c.V;
    if (h->cublas == nullptr) {
        cublasHandle_t hd = nullptr;
        if (cublasCreate(&hd) != CUBLAS_STATUS_SUCCESS) { set_error("cublasCreate failed"); return TVS_ECUDA; }
        cublasSetStream(hd, h->stream);
        cublasSetPointerMode(hd, CUBLAS_POINTER_MODE_HOST);
        h->cublas = hd;
    }
    cublasHandle_t hd = (cublasHandle_t)h->cublas;
    const int S = (int)std::max<int64_t>(1, std::min<int64_t>(R / 256 + 1, (int64_t)h->sm_count * 8 / ((W + 127) / 128)));
    const int64_t rows_per_split = (R + S - 1) / S;
    if (h->dense_p_cap < (size_t)R * (size_t)W) {
        if (h->dense_p) cudaFree(h->dense_p);
        h->dense_p = nullptr; h->dense_p_cap = 0;
        TVS_CUDA(cudaMalloc((void**)&h->dense_p, (size_t)R * (size_t)W * sizeof(double)));
        h->dense_p_cap = (size_t)R * (size_t)W;
    }
    if (h->dense_ll_cap < (size_t)S * (size_t)W) {
        if (h->dense_ll) cudaFree(h->dense_ll);
        h->dense_ll = nullptr; h->dense_ll_cap = 0;
        TVS_CUDA(cudaMalloc((void**)&h->dense_ll, (size_t)S * (size_t)W * sizeof(double)));
        h->dense_ll_cap = (size_t)S * (size_t)W;
    }
    const double one = 1.0, zero = 0.0;
    cublasStatus_t cs = cublasDgemm(hd, CUBLAS_OP_N, CUBLAS_OP_N, (int)W, (int)R, (int)V, &one, a.x, (int)W, c.dense, (int)V, &zero,
                                    h->dense_p, (int)W);
    if (cs != CUBLAS_STATUS_SUCCESS) { set_error("cublasDgemm (p = A x) failed: %d", (int)cs); return TVS_ECUDA; }
    dense_mid_kernel<<<dim3((unsigned)((W + 127) / 128), (unsigned)S), 128, 0, h->stream>>>(h->dense_p, a.w, a.w16, a.done, W, R, rows_per_split,
                                                                                          h->dense_ll, backward ? 1 : 0);
    dense_ll_reduce_kernel<<<(unsigned)((W + 127) / 128), 128, 0, h->stream>>>(h->dense_ll, S, W, a.llpart);
    TVS_CUDA(cudaGetLastError());
    if (backward) {
        cs = cublasDgemm(hd, CUBLAS_OP_N, CUBLAS_OP_T, (int)W, (int)V, (int)R, &one, h->dense_p, (int)W, c.dense, (int)V, &zero,
                         a.tpart, (int)W);
        if (cs != CUBLAS_STATUS_SUCCESS) { set_error("cublasDgemm (t = A^T rr) failed: %d", (int)cs); return TVS_ECUDA; }
    }
    return TVS_OK;
}

__global__ void densify_kernel(const int32_t* __restrict__ row_ptr, const uint32_t* __restrict__ ent, int64_t R, int64_t V,
                               double* __restrict__ dense) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    for (int32_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) dense[r * V + (ent[k] & 0xffffu)] = (double)(ent[k] >> 16);
}



// geometry of ONE blocked product in backward mode for a lane set of width W (the Hessian product of the Newton phase)
struct BlkOne { int K = 2; bool bulk = true; int stages = 2; size_t smem = 0; int tiles = 0; int splits = 1; };
static int plan_blk_backward(const tvs_problem* h, int64_t W, const Csr::Blk& m, BlkOne* out) {
    const size_t limit = 227 * 1024;
    int K = (W % 2 == 0) ? 2 : 1;
    const bool bulk = (W % 2 == 0);
    auto stages_for = [&](int k) {
        int n = std::min(kBlkMaxStages, std::max(2, env_int("TVS_BLK_STAGES", kBlkMaxStages)));
        while (n >= 2 && blk_layout(32 * k, m.max_seg_bytes, n).total > limit) --n;
        return n;
    };
    int ns = stages_for(K);
    if (K == 2 && ns < 2) { K = 1; ns = stages_for(1); }
    if (ns < 2) return TVS_ELIMIT;
    const size_t smem = blk_layout(32 * K, m.max_seg_bytes, ns).total;
    int occ = 0, rc;
    if (K == 2) rc = blk_prepare<2, 1, true>(h, smem, &occ);
    else if (bulk) rc = blk_prepare<1, 1, true>(h, smem, &occ);
    else rc = blk_prepare<1, 1, false>(h, smem, &occ);
    if (rc != TVS_OK) return rc;
    if (occ < 1) return TVS_ELIMIT;
    out->K = K; out->bulk = bulk; out->stages = ns; out->smem = smem;
    out->tiles = (int)((W + 32 * K - 1) / (32 * K));
    const int slots = h->sm_count, units = out->tiles * m.n_chunks;
    int best_s = 1;
    double best = -1.0;
    const int s_max = std::min(std::max(8, slots / std::max(1, units)), std::max(1, m.n_panels / 4));
    for (int cand = 1; cand <= std::max(1, s_max); ++cand) {
        const int64_t total = (int64_t)units * cand;
        const int64_t waves = (total + slots - 1) / slots;
        const double eff = (double)total / (double)(waves * slots);
        if (eff > best + 0.02) { best = eff; best_s = cand; }
    }
    out->splits = std::max(1, std::min(best_s, m.n_panels));
    return TVS_OK;
}

template <int K, bool BULK>
static void blk_launch_one(const BlkArgs& a, int mode, dim3 grid, size_t smem, cudaStream_t st) {
    if (mode == 0) blk_kernel<K, 0, BULK><<<grid, kBlkThreads, smem, st>>>(a);
    else blk_kernel<K, 1, BULK><<<grid, kBlkThreads, smem, st>>>(a);
}

template <bool BW>
static int launch_blk(tvs_problem* h, const PassPlan& pl, const PassArgs& a) {
    const Csr& c = h->csr;
    const int64_t W = a.B, R = c.R, V = c.V;
    if (BW && h->blk_rr_cap < (size_t)R * (size_t)W) {
        if (h->blk_rr) cudaFree(h->blk_rr);
        h->blk_rr = nullptr; h->blk_rr_cap = 0;
        TVS_CUDA(cudaMalloc((void**)&h->blk_rr, (size_t)R * (size_t)W * sizeof(double)));
        h->blk_rr_cap = (size_t)R * (size_t)W;
    }
    BlkArgs f{};
    f.m.seg_off = c.blk_f.seg_off; f.m.blob = (const uint4*)c.blk_f.blob; f.m.n_chunks = c.blk_f.n_chunks;
    f.m.n_panels = c.blk_f.n_panels; f.m.max_seg_bytes = c.blk_f.max_seg_bytes; f.m.dealt = c.blk_f.dealt;
    f.src = a.x; f.n_minor = V; f.n_major = R; f.W = W; f.done = a.done; f.w32 = a.w; f.w16 = a.w16;
    f.rr = BW ? h->blk_rr : nullptr; f.llpart = a.llpart; f.tpart = nullptr; f.stages = pl.stages;
    f.logc = a.logc; f.logtab = a.logtab;
    const dim3 gf((unsigned)pl.tiles, (unsigned)pl.splits);
    if (pl.lpt == 2) blk_launch_one<2, true>(f, 0, gf, pl.smem, h->stream);
    else if (pl.bulk) blk_launch_one<1, true>(f, 0, gf, pl.smem, h->stream);
    else blk_launch_one<1, false>(f, 0, gf, pl.smem, h->stream);
    TVS_CUDA(cudaGetLastError());
    if (BW) {
        BlkArgs b = f;
        b.m.seg_off = c.blk_b.seg_off; b.m.blob = (const uint4*)c.blk_b.blob; b.m.n_chunks = c.blk_b.n_chunks;
        b.m.n_panels = c.blk_b.n_panels; b.m.max_seg_bytes = c.blk_b.max_seg_bytes; b.m.dealt = c.blk_b.dealt;
        b.src = h->blk_rr; b.n_minor = R; b.n_major = V; b.rr = nullptr; b.llpart = nullptr; b.tpart = a.tpart; b.stages = pl.bstages;
        const dim3 gb((unsigned)c.blk_b.n_chunks, (unsigned)pl.tiles, (unsigned)pl.bsplits);   // the chunks of one (tile, split) are neighbours: they stream the same rr panels
        if (pl.lpt == 2) blk_launch_one<2, true>(b, 1, gb, pl.bsmem, h->stream);
        else if (pl.bulk) blk_launch_one<1, true>(b, 1, gb, pl.bsmem, h->stream);
        else blk_launch_one<1, false>(b, 1, gb, pl.bsmem, h->stream);
        TVS_CUDA(cudaGetLastError());
    }
    return TVS_OK;
}

template <bool BW>
static int launch_pass(tvs_problem* h, const PassPlan& pl, const PassArgs& a) {
    if (pl.dense) return launch_dense(h, a, BW);
    if (pl.blk) return launch_blk<BW>(h, pl, a);
    dim3 grid(pl.tiles, pl.splits), block(kThreads);
    if (pl.v3_warps == 16) {
        if (pl.lpt == 4) pass3_kernel<4, 16, BW><<<grid, 512, pl.smem, h->stream>>>(a);
        else if (pl.lpt == 2) pass3_kernel<2, 16, BW><<<grid, 512, pl.smem, h->stream>>>(a);
        else pass3_kernel<1, 16, BW><<<grid, 512, pl.smem, h->stream>>>(a);
        TVS_CUDA(cudaGetLastError());
        return TVS_OK;
    }
    if (pl.v3_warps == 32) {
        if (pl.lpt == 4) pass3_kernel<4, 32, BW><<<grid, 1024, pl.smem, h->stream>>>(a);
        else if (pl.lpt == 2) pass3_kernel<2, 32, BW><<<grid, 1024, pl.smem, h->stream>>>(a);
        else pass3_kernel<1, 32, BW><<<grid, 1024, pl.smem, h->stream>>>(a);
        TVS_CUDA(cudaGetLastError());
        return TVS_OK;
    }
    if (pl.v2) {
        dim3 block2(kThreads2);
        if (pl.lpt == 4) pass2_kernel<4, BW><<<grid, block2, pl.smem, h->stream>>>(a);
        else if (pl.lpt == 2) pass2_kernel<2, BW><<<grid, block2, pl.smem, h->stream>>>(a);
        else pass2_kernel<1, BW><<<grid, block2, pl.smem, h->stream>>>(a);
        TVS_CUDA(cudaGetLastError());
        return TVS_OK;
    }
    if (pl.lpt == 2 && pl.xs) pass_kernel<2, true, BW><<<grid, block, pl.smem, h->stream>>>(a);
    else if (pl.lpt == 1 && pl.xs) pass_kernel<1, true, BW><<<grid, block, pl.smem, h->stream>>>(a);
    else if (pl.lpt == 2) pass_kernel<2, false, BW><<<grid, block, pl.smem, h->stream>>>(a);
    else pass_kernel<1, false, BW><<<grid, block, pl.smem, h->stream>>>(a);
    TVS_CUDA(cudaGetLastError());
    return TVS_OK;
}

static PassArgs base_args(const tvs_problem* h) {
    PassArgs a{};
    const Csr& c = h->csr;
    a.row_ptr = c.row_ptr; a.ent = c.ent; a.chunk_row0 = c.chunk_row0; a.ccol_ptr = c.ccol_ptr; a.ccol_id = c.ccol_id;
    a.ccol_start = c.ccol_start; a.cent = c.cent; a.ent3 = c.ent3; a.cent3 = c.cent3; a.w = h->lanes.w; a.w16 = h->lanes.has16 ? h->lanes.w16 : nullptr; a.B = h->lanes.B; a.V = (int)c.V;
    a.n_chunks = c.n_chunks; a.rows_per_chunk = c.rows_per_chunk;
    a.stage_cap = std::max(c.max_chunk_nnz, 1);
    a.desc = (const ChunkDesc*)c.desc; a.logc = make_log_coef();
    a.logtab = (const double2*)c.logtab;
    return a;
}

// ------------------------------------------------------------------------------------- fit state
template <class T>
static int fit_alloc(FitState& f, T** p, size_t n) {
    int rc = dev_alloc(p, n);
    if (rc == TVS_OK) f.owned.push_back((void*)*p);
    return rc;
}

static int ensure_fit_state(tvs_problem* h, int m) {
    FitState& f = h->fs;
    const int64_t V = h->csr.V, R = h->csr.R;
    // every array below is [rows x lanes] with the CURRENT lane count as stride, so a state allocated for more lanes
    // serves any smaller batch: successive batches of a model selection (1, 20, 19, ... lanes) reuse one allocation
    if (f.V == V && h->lanes.B <= f.B && f.m == m && f.R == R) return TVS_OK;
    const int64_t B = (h->lanes.B + 31) / 32 * 32;              // capacity
    free_fit(f);
    const size_t vb = (size_t)V * B;
    int rc;
#define A_(ptr, n) if ((rc = fit_alloc(f, &(ptr), (size_t)(n))) != TVS_OK) { free_fit(f); return rc; }
    for (int i = 0; i < 2; ++i) {
        LaneSet& s = f.set[i];
        A_(s.orig, B) A_(s.z, vb) A_(s.zt, vb) A_(s.g, vb) A_(s.d, vb) A_(s.x, vb)
        A_(s.S, vb * m) A_(s.Y, vb * m) A_(s.rho, (size_t)B * m)
        A_(s.gamma, B) A_(s.F, B) A_(s.dg, B) A_(s.alpha, B) A_(s.llsum, B) A_(s.sumphi, B) A_(s.kkt, 2 * (size_t)B)
        A_(s.gram, (size_t)B * ((2 * m + 1) * (2 * m + 2) / 2))
        A_(s.ls, B) A_(s.iters, B) A_(s.head, B) A_(s.cnt, B) A_(s.done, B) A_(s.status, B)
    }
    f.half = B * 3 / 4 + 4;        // compaction keeps at most 75 % of a set; +3: compacted widths are padded to a multiple of 4
    for (int i = 0; i < 2; ++i) {
        A_(f.mbuf[i], (size_t)V * f.half) A_(f.nbuf[i], f.half) A_(f.cbuf[i], f.half)      // wbuf / wbuf16: on first compaction
    }
    A_(f.gt, vb) A_(f.gh, vb)
    // partial buffers: splits*width never exceeds (resident CTAs)*(lanes per tile) + width
    f.tpart_cap = (size_t)V * ((size_t)h->sm_count * 16 * 64 + (size_t)B);
    f.llpart_cap = f.tpart_cap / (size_t)V;
    A_(f.tpart, f.tpart_cap) A_(f.llpart, f.llpart_cap) A_(f.t_red, vb) A_(f.ll_red, B)
    A_(f.idx, B + 4) A_(f.nactive, 32)
    A_(f.gha, vb) A_(f.lam, B) A_(f.need, B)
    A_(f.theta_io, vb) A_(f.ll_out, B) A_(f.kkt_out, 2 * (size_t)B) A_(f.iters_out, B) A_(f.status_out, B)
#undef A_
    if (cudaMallocHost((void**)&f.h_nactive, 32 * sizeof(int32_t)) != cudaSuccess) { free_fit(f); set_error("cudaMallocHost failed"); return TVS_ENOMEM; }
    f.V = V; f.B = B; f.m = m; f.R = R;
    return TVS_OK;
}

template <typename T>
static void gather_rows(cudaStream_t st, const T* src, T* dst, int64_t rows, int64_t Bs, int64_t Bd, const int32_t* idx) {
    if (!src || !dst || rows <= 0 || Bd <= 0) return;
    dim3 grid((unsigned)((Bd + 127) / 128), (unsigned)std::min<int64_t>(rows, 4096));
    gather_rows_kernel<T><<<grid, 128, 0, st>>>(src, dst, rows, Bs, Bd, idx);
}

}  // namespace tvs

using namespace tvs;

// =============================================================================================== API
extern "C" {

int tvs_version(void) { return 100; }
const char* tvs_last_error(void) { return g_err; }

int tvs_device_count(int* n_out) {
    if (!n_out) { set_error("n_out is null"); return TVS_EINVAL; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *n_out = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    *n_out = n;
    return TVS_OK;
}

int tvs_create(tvs_handle* out, int device, int64_t R, int64_t V, int64_t nnz, const int32_t* row_ptr,
               const int32_t* col, const int32_t* val, const int64_t* L) {
    if (!out) { set_error("out is null"); return TVS_EINVAL; }
    *out = nullptr;
    if (R <= 0 || V <= 0 || nnz < 0 || !row_ptr || (nnz > 0 && (!col || !val)) || !L) {
        set_error("tvs_create: bad sizes or null arrays (R=%lld V=%lld nnz=%lld)", (long long)R, (long long)V, (long long)nnz);
        return TVS_EINVAL;
    }
    if (V > 65535 || R > 2000000000LL || nnz > 2000000000LL) { set_error("tvs_create: V>65535 or R/nnz>2e9 unsupported"); return TVS_ELIMIT; }
    if (row_ptr[0] != 0 || row_ptr[R] != nnz) { set_error("tvs_create: row_ptr[0]!=0 or row_ptr[R]!=nnz"); return TVS_EINVAL; }
    for (int64_t r = 0; r < R; ++r) {
        if (row_ptr[r + 1] < row_ptr[r]) { set_error("tvs_create: row_ptr not monotone at row %lld", (long long)r); return TVS_EINVAL; }
        for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            if (col[k] < 0 || col[k] >= V) { set_error("tvs_create: column %d out of range at nnz %lld", col[k], (long long)k); return TVS_EINVAL; }
            if (k > row_ptr[r] && col[k] <= col[k - 1]) { set_error("tvs_create: columns not strictly increasing in row %lld", (long long)r); return TVS_EINVAL; }
            if (val[k] < 1) { set_error("tvs_create: count < 1 at nnz %lld", (long long)k); return TVS_EINVAL; }
            if (val[k] > 65535) { set_error("tvs_create: count %d > 65535 at nnz %lld", val[k], (long long)k); return TVS_ELIMIT; }
        }
    }
    for (int64_t v = 0; v < V; ++v)
        if (L[v] <= 0) { set_error("tvs_create: variant size L[%lld] <= 0", (long long)v); return TVS_EINVAL; }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("tvs_create: no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return TVS_ECUDA;
    }
    if (device < 0 || device >= ndev) { set_error("tvs_create: device %d out of range (0..%d)", device, ndev - 1); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(device));

    tvs_problem* h = new (std::nothrow) tvs_problem();
    if (!h) { set_error("out of host memory"); return TVS_ENOMEM; }
    h->device = device;
    int rc = TVS_OK;
    auto fail = [&](int code) { tvs_destroy(h); return code; };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(TVS_ECUDA);
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TVS_ECUDA);
    cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1); cudaEventCreate(&h->evp0); cudaEventCreate(&h->evp1);

    Csr& c = h->csr;
    c.R = R; c.V = V; c.nnz = nnz;
    // rows per chunk (a multiple of the 16 warps, so that every warp owns the same number of rows: 136 rows cost 131 us
    // per full-width cfg2 pass where 144 cost 121): the largest of 160, 144, ... 96 that the widest staged kernel
    // (4 lanes per thread, 16 warps) holds in shared memory while the matrix still has >= 96 chunks to split over CTAs
    // (cfg2, V = 64 with the t accumulators in registers: 144 -- 80 rows 139.8 us, 112 131.6, 144 127.5 before the other
    // changes of the round); else 80 when that fits, else 64
    int rpc = env_int("TVS_RC", 0);
    if (rpc <= 0) {
        auto fits = [&](int rows) {
            int64_t cap = 1;
            for (int64_t r0 = 0; r0 < R; r0 += rows) cap = std::max<int64_t>(cap, row_ptr[std::min<int64_t>(R, r0 + rows)] - row_ptr[r0]);
            return pass3_layout((int)V, 128, rows, (int)cap, true, 16).total <= 225 * 1024;
        };
        rpc = 64;
        if (fits(80)) rpc = 80;
        for (int cand = 160; cand >= 96; cand -= 16)
            if (R / cand >= 96 && fits(cand)) { rpc = cand; break; }
    }
    rpc = std::max(kWarps2, std::min(rpc, 4096));
    c.rows_per_chunk = rpc;
    c.n_chunks = (int)((R + rpc - 1) / rpc);

    // host build: packed row-major entries, per-chunk column-major entries
    std::vector<uint32_t> ent((size_t)std::max<int64_t>(nnz, 1)), cent((size_t)std::max<int64_t>(nnz, 1));
    std::vector<int32_t> chunk_row0(c.n_chunks + 1), ccol_ptr(c.n_chunks + 1), ccol_id, ccol_start;
    std::vector<int32_t> cnt((size_t)V), pos((size_t)V);
    for (int64_t k = 0; k < nnz; ++k) ent[k] = (uint32_t)col[k] | ((uint32_t)val[k] << 16);
    int64_t cursor = 0;
    for (int ch = 0; ch < c.n_chunks; ++ch) {
        const int64_t r0 = (int64_t)ch * rpc, r1 = std::min<int64_t>(R, r0 + rpc);
        chunk_row0[ch] = (int32_t)r0;
        ccol_ptr[ch] = (int32_t)ccol_id.size();
        std::fill(cnt.begin(), cnt.end(), 0);
        for (int64_t k = row_ptr[r0]; k < row_ptr[r1]; ++k) cnt[col[k]]++;
        for (int64_t v = 0; v < V; ++v)
            if (cnt[v]) { ccol_id.push_back((int32_t)v); ccol_start.push_back((int32_t)cursor); pos[v] = (int32_t)cursor; cursor += cnt[v]; }
        for (int64_t r = r0; r < r1; ++r)
            for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k)
                cent[pos[col[k]]++] = (uint32_t)(r - r0) | ((uint32_t)val[k] << 16);
    }
    chunk_row0[c.n_chunks] = (int32_t)R;
    ccol_ptr[c.n_chunks] = (int32_t)ccol_id.size();
    c.max_chunk_nnz = 0;
    for (int ch = 0; ch < c.n_chunks; ++ch) {
        const int64_t r0 = (int64_t)ch * rpc, r1 = std::min<int64_t>(R, r0 + rpc);
        c.max_chunk_nnz = std::max(c.max_chunk_nnz, (int)(row_ptr[r1] - row_ptr[r0]));
    }
    ccol_start.push_back((int32_t)cursor);
    c.n_ccol = (int64_t)ccol_id.size();
    std::vector<double> invL((size_t)V), Ld((size_t)V);
    for (int64_t v = 0; v < V; ++v) { Ld[v] = (double)L[v]; invL[v] = 1.0 / Ld[v]; }

#define UP_(dst, vec, T)                                                                                  \
    if ((rc = dev_alloc(&c.dst, (vec).size())) != TVS_OK) return fail(rc);                                \
    if (cudaMemcpy(c.dst, (vec).data(), (vec).size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) \
        return fail(cuda_fail(cudaGetLastError(), "upload " #dst, __FILE__, __LINE__));
    if ((rc = dev_alloc(&c.row_ptr, (size_t)R + 1)) != TVS_OK) return fail(rc);
    if (cudaMemcpy(c.row_ptr, row_ptr, ((size_t)R + 1) * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "upload row_ptr", __FILE__, __LINE__));
    {   // log table: c_i = 1/m_i rounded to 24 bits (so that m c_i - 1 has few bits), l_i = -log(c_i)
        std::vector<double> tab(256);
        for (int i = 0; i < 128; ++i) {
            const double mi = 1.0 + (i + 0.5) / 128.0;
            const double ci = (double)(float)(1.0 / mi);
            tab[2 * i] = ci;
            tab[2 * i + 1] = (double)(-logl((long double)ci));
        }
        double* dt = nullptr;
        if ((rc = dev_alloc(&dt, tab.size())) != TVS_OK) return fail(rc);
        c.logtab = dt;
        if (cudaMemcpy(dt, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
            return fail(cuda_fail(cudaGetLastError(), "upload logtab", __FILE__, __LINE__));
    }
    std::vector<ChunkDesc> descv((size_t)c.n_chunks);
    for (int ch = 0; ch < c.n_chunks; ++ch) {
        ChunkDesc& d = descv[ch];
        d.row0 = chunk_row0[ch]; d.nrows = chunk_row0[ch + 1] - chunk_row0[ch];
        d.e0 = row_ptr[d.row0]; d.nent = row_ptr[d.row0 + d.nrows] - d.e0;
        d.c0 = ccol_ptr[ch]; d.ncol = ccol_ptr[ch + 1] - ccol_ptr[ch];
        d.ce0 = ccol_start[d.c0]; d.ncent = ccol_start[d.c0 + d.ncol] - d.ce0;
    }
    {
        ChunkDesc* dd = nullptr;
        if ((rc = dev_alloc(&dd, descv.size())) != TVS_OK) return fail(rc);
        c.desc = dd;
        if (cudaMemcpy(dd, descv.data(), descv.size() * sizeof(ChunkDesc), cudaMemcpyHostToDevice) != cudaSuccess)
            return fail(cuda_fail(cudaGetLastError(), "upload desc", __FILE__, __LINE__));
    }
    {   // pass3_kernel's entries: the count as the top 16 bits of its fp64 form (rebuilt without an fp64 instruction)
        auto hi16 = [](uint32_t v, bool* ok) {
            const double d = (double)v;
            uint64_t bits;
            memcpy(&bits, &d, 8);
            if (bits & 0x0000ffffffffffffull) *ok = false;
            return (uint32_t)(bits >> 48);
        };
        bool ok = true;
        std::vector<uint32_t> ent3(ent.size()), cent3(cent.size());
        for (size_t k = 0; k < (size_t)nnz; ++k) {
            ent3[k] = (ent[k] & 0xffffu) | (hi16(ent[k] >> 16, &ok) << 16);
            cent3[k] = (cent[k] & 0xffffu) | (hi16(cent[k] >> 16, &ok) << 16);
        }
        c.ent3_ok = ok;
        UP_(ent3, ent3, uint32_t) UP_(cent3, cent3, uint32_t)
    }
    UP_(ent, ent, uint32_t) UP_(cent, cent, uint32_t) UP_(chunk_row0, chunk_row0, int32_t) UP_(ccol_ptr, ccol_ptr, int32_t)
    if (ccol_id.empty()) ccol_id.push_back(0);
    UP_(ccol_id, ccol_id, int32_t) UP_(ccol_start, ccol_start, int32_t) UP_(invL, invL, double) UP_(Ld, Ld, double)
    // Hessian slabs of the Newton phase (tvs_newton.cuh)
    if (V <= kNewtonMaxV && env_int("TVS_NEWTON_LANES", 128) > 0) {
        int64_t np = 0;
        bool exact = true;
        for (int64_t r = 0; r < R; ++r) { const int64_t n = row_ptr[r + 1] - row_ptr[r]; np += n * (n + 1) / 2; }
        for (int64_t k = 0; k < nnz && exact; ++k) exact = val[k] < 4096;          // products stay below 2^24: exact in float
        if (exact && np <= (int64_t)48 << 20) {
            const int HB = std::max(32, std::min(env_int("TVS_HESS_ROWS", 1024), 16384));
            const int NPk = (int)(V * (V + 1) / 2), NG = (NPk + 31) / 32;
            const int nblk = (int)((R + HB - 1) / HB);
            std::vector<int64_t> sp((size_t)nblk * (NG + 1));
            std::vector<uint16_t> hrow, hslot((size_t)nblk * NG * 32, (uint16_t)0xFFFF);
            std::vector<float> hval;
            std::vector<std::vector<std::pair<uint16_t, float>>> lists((size_t)NPk);
            std::vector<int> order((size_t)NPk);
            int64_t steps = 0;
            for (int b = 0; b < nblk; ++b) {
                const int64_t rb0 = (int64_t)b * HB, rb1 = std::min<int64_t>(R, rb0 + HB);
                for (auto& l : lists) l.clear();
                for (int64_t r = rb0; r < rb1; ++r)
                    for (int64_t i = row_ptr[r]; i < row_ptr[r + 1]; ++i)
                        for (int64_t j = row_ptr[r]; j <= i; ++j)           // col[i] >= col[j]
                            lists[(size_t)(col[i] * (col[i] + 1) / 2 + col[j])].push_back({(uint16_t)(r - rb0), (float)val[i] * (float)val[j]});
                for (int i = 0; i < NPk; ++i) order[i] = i;
                std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return lists[x].size() > lists[y].size(); });
                for (int e = 0; e < NG; ++e) {
                    sp[(size_t)b * (NG + 1) + e] = steps;
                    const size_t len = (e * 32 < NPk) ? lists[order[e * 32]].size() : 0;
                    for (size_t k = 0; k < len; ++k)
                        for (int t = 0; t < 32; ++t) {
                            const int pos = e * 32 + t;
                            if (pos < NPk && k < lists[order[pos]].size()) { hrow.push_back(lists[order[pos]][k].first); hval.push_back(lists[order[pos]][k].second); }
                            else { hrow.push_back((uint16_t)HB); hval.push_back(0.f); }
                        }
                    for (int t = 0; t < 32; ++t) {
                        const int pos = e * 32 + t;
                        if (pos < NPk && !lists[order[pos]].empty()) hslot[((size_t)b * NG + e) * 32 + t] = (uint16_t)order[pos];
                    }
                    steps += (int64_t)len;
                }
                sp[(size_t)b * (NG + 1) + NG] = steps;
            }
            if (hrow.empty()) { hrow.push_back(0); hval.push_back(0.f); }
            c.h_HB = HB; c.h_nblk = nblk; c.h_NG = NG; c.h_steps = steps;
            UP_(slab_ptr, sp, int64_t) UP_(h_row, hrow, uint16_t) UP_(h_val, hval, float) UP_(h_slot, hslot, uint16_t)
            // entry-major lists inside blocks of WB rows for hess_wide_kernel
            const int WB = std::max(32, std::min(env_int("TVS_HESS_WIDE_ROWS", 128), 512));
            const int wnblk = (int)((R + WB - 1) / WB);
            const int NS = (NPk + kWideEnt - 1) / kWideEnt;
            std::vector<int32_t> wp((size_t)wnblk * (NS + 1));
            std::vector<uint2> wrec((size_t)std::max<int64_t>(np, 1));
            std::vector<std::vector<uint2>> sub((size_t)NS);
            int64_t q = 0;
            int wcap = 1;
            for (int b = 0; b < wnblk; ++b) {
                const int64_t rb0 = (int64_t)b * WB, rb1 = std::min<int64_t>(R, rb0 + WB);
                for (auto& l : sub) l.clear();
                for (int64_t r = rb0; r < rb1; ++r)                          // row-major inside every subset
                    for (int64_t i = row_ptr[r]; i < row_ptr[r + 1]; ++i)
                        for (int64_t j = row_ptr[r]; j <= i; ++j) {
                            const int e = col[i] * (col[i] + 1) / 2 + col[j];
                            const float pv = (float)val[i] * (float)val[j];
                            uint32_t bits;
                            memcpy(&bits, &pv, 4);
                            sub[(size_t)(e / kWideEnt)].push_back(make_uint2((uint32_t)(r - rb0) | ((uint32_t)(e % kWideEnt) << 16), bits));
                        }
                for (int sidx = 0; sidx < NS; ++sidx) {
                    wp[(size_t)b * (NS + 1) + sidx] = (int32_t)q;
                    for (auto& rec : sub[sidx]) wrec[q++] = rec;
                }
                wp[(size_t)b * (NS + 1) + NS] = (int32_t)q;
                for (int s0 = 0; s0 < NS; s0 += kWideWarps)
                    wcap = std::max(wcap, (int)(wp[(size_t)b * (NS + 1) + std::min(NS, s0 + kWideWarps)] - wp[(size_t)b * (NS + 1) + s0]));
            }
            c.w_HB = WB; c.w_nblk = wnblk; c.w_cap = wcap;
            // the same products once more in blocked form, grouped by H entry (major) with the read path as minor index:
            // the Hessian of every lane is then the blocked product G x (w / p^2), one launch of blk_kernel's backward mode
            if (env_int("TVS_HESS_BLK", 1) != 0) {
                bool small = true;
                for (int64_t k = 0; k < nnz && small; ++k) small = val[k] <= 46340;       // products below 2^31
                if (small) {
                    std::vector<int64_t> gp((size_t)NPk + 1, 0);
                    for (int64_t r = 0; r < R; ++r)
                        for (int64_t i = row_ptr[r]; i < row_ptr[r + 1]; ++i)
                            for (int64_t j = row_ptr[r]; j <= i; ++j) gp[(size_t)(col[i] * (col[i] + 1) / 2 + col[j]) + 1]++;
                    for (int e = 0; e < NPk; ++e) gp[(size_t)e + 1] += gp[(size_t)e];
                    std::vector<int32_t> gr((size_t)std::max<int64_t>(np, 1)), gv((size_t)std::max<int64_t>(np, 1));
                    std::vector<int64_t> at(gp.begin(), gp.end() - 1);
                    for (int64_t r = 0; r < R; ++r)                          // rows ascending inside every entry
                        for (int64_t i = row_ptr[r]; i < row_ptr[r + 1]; ++i)
                            for (int64_t j = row_ptr[r]; j <= i; ++j) {
                                const int64_t q = at[(size_t)(col[i] * (col[i] + 1) / 2 + col[j])]++;
                                gr[(size_t)q] = (int32_t)r; gv[(size_t)q] = val[i] * val[j];
                            }
                    BlkHost hg;
                    if (build_blk<int64_t>(NPk, R, gp.data(), gr.data(), gv.data(), &hg)) {
                        if ((rc = upload_blk(hg, &c.blk_h)) != TVS_OK) return fail(rc);
                        c.has_blk_h = true;
                    }
                }
            }
            UP_(wide_ptr, wp, int32_t)
            {
                uint2* dd = nullptr;
                if ((rc = dev_alloc(&dd, wrec.size())) != TVS_OK) return fail(rc);
                c.wide_rec = dd;
                if (cudaMemcpy(dd, wrec.data(), wrec.size() * sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess)
                    return fail(cuda_fail(cudaGetLastError(), "upload wide_rec", __FILE__, __LINE__));
            }
        }
    }
#undef UP_
    {   // dense copy of A for the DGEMM pass: only where the staged sparse kernels do not apply (their x / t tiles need
        // V <= ~400) and the matrix is dense enough for a full contraction to win (TVS_DENSE=1 forces, 0 forbids)
        const int want = env_int("TVS_DENSE", -1);
        const double density = (double)nnz / ((double)R * (double)V);
        const double min_density = env_int("TVS_DENSE_MIN_PCT", 22) / 100.0;   // measured crossover against blk_kernel ~21 % (profiles/r2_summary.md; ~10 % against the round-1 sparse kernel)
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const bool fits = (double)R * (double)V * 8.0 < 0.25 * (double)free_b;
        if (want != 0 && fits && nnz > 0 && (want == 1 || (V > 400 && density >= min_density))) {
            if ((rc = dev_alloc(&c.dense, (size_t)R * (size_t)V)) != TVS_OK) return fail(rc);
            cudaMemset(c.dense, 0, (size_t)R * (size_t)V * sizeof(double));
            densify_kernel<<<(unsigned)((R + 127) / 128), 128>>>(c.row_ptr, c.ent, R, V, c.dense);
            if (cudaDeviceSynchronize() != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "densify", __FILE__, __LINE__));
        }
    }
    {   // blocked copies of A and A^T for blk_kernel: candidate sets too wide for the staged kernel's shared-memory tiles
        const int want = env_int("TVS_BLK", -1);
        // (V > ~330), or whose staged kernel is left with one lane per thread (V > ~160: the blocked kernel is 1.46x faster on
        // wide lane sets there -- cfg3, 1,024 lanes: 1.21 ms against 1.76 ms per pass)
        if (nnz > 0 && want != 0 && (want == 1 || V > env_int("TVS_BLK_MIN_V", 160)) && c.dense == nullptr) {
            BlkHost hf, hb;
            const bool ok = build_blk<int32_t>(R, V, row_ptr, col, val, &hf) && build_blk_transposed(R, V, nnz, row_ptr, col, val, &hb);
            if (ok) {
                if ((rc = upload_blk(hf, &c.blk_f)) != TVS_OK || (rc = upload_blk(hb, &c.blk_b)) != TVS_OK) return fail(rc);
                c.has_blk = true;
            }
        }
    }
    *out = h;
    return TVS_OK;
}

int tvs_destroy(tvs_handle h) {
    if (!h) return TVS_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_fit(h->fs); free_lanes(h->lanes); free_csr(h->csr);
    if (h->cublas) cublasDestroy((cublasHandle_t)h->cublas);
    if (h->dense_p) cudaFree(h->dense_p);
    if (h->dense_ll) cudaFree(h->dense_ll);
    if (h->blk_rr) cudaFree(h->blk_rr);
    if (h->ll_scratch) cudaFree(h->ll_scratch);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evp0) cudaEventDestroy(h->evp0);
    if (h->evp1) cudaEventDestroy(h->evp1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return TVS_OK;
}

int tvs_dims(tvs_handle h, int64_t* R, int64_t* V, int64_t* nnz, int64_t* B) {
    if (!h) { set_error("null handle"); return TVS_EINVAL; }
    if (R) *R = h->csr.R;
    if (V) *V = h->csr.V;
    if (nnz) *nnz = h->csr.nnz;
    if (B) *B = h->lanes_set ? h->lanes.B : 0;
    return TVS_OK;
}

// Lane weights live on the device in one or both of two forms: int32 (`w`, has32) and 16-bit counts (`w16`, has16).  The
// pass kernels read the 16-bit form when it exists (half the traffic of the only HBM stream of a pass: a 1000-lane cfg2
// pass reads 36 MB instead of 72 MB and the solver state stays in L2 between the pass and the update kernel).
static int reserve_lanes(tvs_problem* h, int64_t B, bool with_mask, bool with_C, bool need32) {
    Lanes& l = h->lanes;
    const int64_t R = h->csr.R, V = h->csr.V;
    int rc;
    if (need32 && l.cap_w < R * B) { dev_free(l.w); l.cap_w = 0; if ((rc = dev_alloc(&l.w, (size_t)(R * B))) != TVS_OK) return rc; l.cap_w = R * B; }
    if (l.cap_w16 < R * B) { dev_free(l.w16); l.cap_w16 = 0; if ((rc = dev_alloc(&l.w16, (size_t)(R * B))) != TVS_OK) return rc; l.cap_w16 = R * B; }
    if (!l.flag && (rc = dev_alloc(&l.flag, 1)) != TVS_OK) return rc;
    if (with_mask) {
        if (l.cap_mask < V * B || !l.mask) { dev_free(l.mask); if ((rc = dev_alloc(&l.mask, (size_t)(V * B))) != TVS_OK) return rc; l.cap_mask = V * B; }
    } else { dev_free(l.mask); l.cap_mask = 0; }
    if (l.cap_b < B || !l.N) {
        dev_free(l.N); dev_free(l.C);
        if ((rc = dev_alloc(&l.N, (size_t)B)) != TVS_OK) return rc;
        l.cap_b = B;
    }
    if (with_C) { if (!l.C && (rc = dev_alloc(&l.C, (size_t)l.cap_b)) != TVS_OK) return rc; }
    else dev_free(l.C);
    l.B = B; l.has32 = false; l.has16 = false;
    return TVS_OK;
}

// the int32 form on demand (tvs_get_weights after a 16-bit upload)
static int ensure_w32(tvs_problem* h) {
    Lanes& l = h->lanes;
    if (l.has32) return TVS_OK;
    const int64_t n = h->csr.R * l.B;
    int rc;
    if (l.cap_w < n) { dev_free(l.w); l.cap_w = 0; if ((rc = dev_alloc(&l.w, (size_t)n)) != TVS_OK) return rc; l.cap_w = n; }
    widen_u16_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, h->stream>>>(l.w16, l.w, n);
    TVS_CUDA(cudaGetLastError());
    l.has32 = true;
    return TVS_OK;
}

// `from16`: the weights were uploaded as 16-bit counts (l.w16); otherwise l.w holds them and the 16-bit copy is made here.
static int finish_lanes(tvs_problem* h, const uint8_t* mask, const double* C, bool from16) {
    Lanes& l = h->lanes;
    const int64_t V = h->csr.V, B = l.B, n = h->csr.R * B;
    if (mask) TVS_CUDA(cudaMemcpyAsync(l.mask, mask, (size_t)(V * B), cudaMemcpyDefault, h->stream));
    if (C) TVS_CUDA(cudaMemcpyAsync(l.C, C, (size_t)B * sizeof(double), cudaMemcpyDefault, h->stream));
    const bool narrow = !from16 && env_int("TVS_W16", 1) != 0;
    if (narrow) {
        cudaMemsetAsync(l.flag, 0, sizeof(int32_t), h->stream);
        narrow_u16_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, h->stream>>>(l.w, l.w16, n, l.flag);
    }
    {   // N_b = sum_r w[r,b] (exact integer sums; the solver state is not allocated yet, so use a scratch buffer)
        if (l.cap_acc < B) {
            if (l.acc) cudaFree(l.acc);
            l.acc = nullptr; l.cap_acc = 0;
            TVS_CUDA(cudaMalloc((void**)&l.acc, (size_t)B * sizeof(unsigned long long)));
            l.cap_acc = B;
        }
        unsigned long long* acc = l.acc;
        cudaMemsetAsync(acc, 0, (size_t)B * sizeof(unsigned long long), h->stream);
        const unsigned tiles = (unsigned)((B + 31) / 32);
        const unsigned slices = (unsigned)std::max<int64_t>(1, std::min<int64_t>(h->csr.R / 64 + 1, (int64_t)(4 * h->sm_count) / tiles + 1));
        if (from16) lane_sum_kernel<uint16_t><<<dim3(tiles, slices), 256, 0, h->stream>>>(l.w16, h->csr.R, B, acc);
        else lane_sum_kernel<int32_t><<<dim3(tiles, slices), 256, 0, h->stream>>>(l.w, h->csr.R, B, acc);
        lane_sum_finish_kernel<<<(unsigned)((B + 255) / 256), 256, 0, h->stream>>>(acc, B, l.N);
    }
    TVS_CUDA(cudaGetLastError());
    l.has32 = !from16; l.has16 = from16;
    if (narrow) {          // the host needs the overflow flag to choose the form the pass kernels read
        int32_t overflow = 0;
        TVS_CUDA(cudaMemcpyAsync(&overflow, l.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        TVS_CUDA(cudaStreamSynchronize(h->stream));
        l.has16 = overflow == 0;
        if (overflow & 2) { h->lanes_set = false; set_error("negative weight in the lane matrix (weights are record counts: w >= 0)"); return TVS_EINVAL; }
    }
    // 16-bit uploads do not synchronise: the copy of the caller's buffer and the kernels above are queued on the handle's
    // stream, tvs_fit / tvs_loglik follow on the same stream (include/tvs.h: the buffer must stay valid until the next
    // call on the handle has returned)
    h->lanes_set = true;
    return TVS_OK;
}

int tvs_set_lanes(tvs_handle h, int64_t B, const int32_t* w, const uint8_t* mask, const double* C) {
    if (!h || B <= 0 || !w) { set_error("tvs_set_lanes: null handle/weights or B<=0"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    h->lanes_set = false;
    int rc = reserve_lanes(h, B, mask != nullptr, C != nullptr, true);
    if (rc != TVS_OK) return rc;
    TVS_CUDA(cudaMemcpyAsync(h->lanes.w, w, (size_t)(h->csr.R * B) * sizeof(int32_t), cudaMemcpyDefault, h->stream));
    return finish_lanes(h, mask, C, false);
}

int tvs_set_lanes_u16(tvs_handle h, int64_t B, const uint16_t* w, const uint8_t* mask, const double* C) {
    if (!h || B <= 0 || !w) { set_error("tvs_set_lanes_u16: null handle/weights or B<=0"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    h->lanes_set = false;
    int rc = reserve_lanes(h, B, mask != nullptr, C != nullptr, false);
    if (rc != TVS_OK) return rc;
    TVS_CUDA(cudaMemcpyAsync(h->lanes.w16, w, (size_t)(h->csr.R * B) * sizeof(uint16_t), cudaMemcpyDefault, h->stream));
    return finish_lanes(h, mask, C, true);
}

static int check_slots(const int32_t* ids, int64_t B, int64_t limit, const char* who, std::vector<int32_t>* out) {
    out->resize((size_t)B);
    TVS_CUDA(cudaMemcpy(out->data(), ids, (size_t)B * sizeof(int32_t), cudaMemcpyDefault));
    for (int64_t b = 0; b < B; ++b)
        if ((*out)[(size_t)b] < 0 || (*out)[(size_t)b] >= limit) { set_error("%s: column %d of lane %lld out of range (0..%lld)", who, (*out)[(size_t)b], (long long)b, (long long)limit - 1); return TVS_EINVAL; }
    return TVS_OK;
}

int tvs_set_lanes_indexed(tvs_handle h, int64_t B, int64_t n_cols, const int32_t* w_cols, const int32_t* col_of_lane,
                          const uint8_t* mask, const double* C) {
    if (!h || B <= 0 || n_cols <= 0 || !w_cols || !col_of_lane) { set_error("tvs_set_lanes_indexed: null argument or B/n_cols<=0"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    h->lanes_set = false;
    const int64_t R = h->csr.R;
    std::vector<int32_t> idx;
    int rc = check_slots(col_of_lane, B, n_cols, "tvs_set_lanes_indexed", &idx);
    if (rc != TVS_OK) return rc;
    if ((rc = reserve_lanes(h, B, mask != nullptr, C != nullptr, true)) != TVS_OK) return rc;
    int32_t *d_cols = nullptr, *d_idx = nullptr;
    if ((rc = dev_alloc(&d_cols, (size_t)(R * n_cols))) != TVS_OK) return rc;
    if ((rc = dev_alloc(&d_idx, (size_t)B)) != TVS_OK) { cudaFree(d_cols); return rc; }
    cudaMemcpyAsync(d_cols, w_cols, (size_t)(R * n_cols) * sizeof(int32_t), cudaMemcpyDefault, h->stream);
    cudaMemcpyAsync(d_idx, idx.data(), (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
    gather_rows(h->stream, (const int32_t*)d_cols, h->lanes.w, R, n_cols, B, d_idx);
    cudaError_t e = cudaGetLastError();
    rc = (e == cudaSuccess) ? finish_lanes(h, mask, C, false) : cuda_fail(e, "expand weight columns", __FILE__, __LINE__);
    cudaFree(d_cols); cudaFree(d_idx);
    return rc;
}

// ------------------------------------------------------------------------------------------ column store
int tvs_columns_reserve(tvs_handle h, int64_t n_cols) {
    if (!h || n_cols <= 0) { set_error("tvs_columns_reserve: null handle or n_cols<=0"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    Lanes& l = h->lanes;
    if (n_cols <= l.col_cap) return TVS_OK;
    const int64_t R = h->csr.R;
    int32_t* fresh = nullptr;
    uint16_t* fresh16 = nullptr;
    int rc = dev_alloc(&fresh, (size_t)(R * n_cols));
    if (rc != TVS_OK) return rc;
    if ((rc = dev_alloc(&fresh16, (size_t)(R * n_cols))) != TVS_OK) { cudaFree(fresh); return rc; }
    if (!l.flag && (rc = dev_alloc(&l.flag, 1)) != TVS_OK) { cudaFree(fresh); cudaFree(fresh16); return rc; }
    if (l.cols) {      // keep the registered columns: row pitch changes from col_cap to n_cols
        cudaMemcpy2DAsync(fresh, (size_t)n_cols * 4, l.cols, (size_t)l.col_cap * 4, (size_t)l.col_cap * 4, (size_t)R, cudaMemcpyDeviceToDevice, h->stream);
        cudaMemcpy2DAsync(fresh16, (size_t)n_cols * 2, l.cols16, (size_t)l.col_cap * 2, (size_t)l.col_cap * 2, (size_t)R, cudaMemcpyDeviceToDevice, h->stream);
        cudaStreamSynchronize(h->stream);
        cudaFree(l.cols); cudaFree(l.cols16);
    }
    l.cols = fresh; l.cols16 = fresh16; l.col_cap = n_cols;
    TVS_CUDA(cudaGetLastError());
    return TVS_OK;
}

static int columns_range_ok(tvs_problem* h, int64_t first, int64_t n, const char* who) {
    if (first < 0 || n <= 0 || first + n > h->lanes.col_cap) {
        set_error("%s: columns %lld..%lld outside the reserved store (%lld)", who, (long long)first, (long long)(first + n - 1), (long long)h->lanes.col_cap);
        return TVS_EINVAL;
    }
    return TVS_OK;
}

// 16-bit mirror of freshly registered columns; a count above 65,535 retires the mirror, a negative weight is an error
static int mirror_columns(tvs_problem* h, int64_t first, int64_t n, const char* who) {
    Lanes& l = h->lanes;
    const int64_t cnt = h->csr.R * n;
    cudaMemsetAsync(l.flag, 0, sizeof(int32_t), h->stream);
    narrow_columns_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(l.cols, l.cols16, l.col_cap, first, n, h->csr.R, l.flag);
    int32_t flag = 0;
    TVS_CUDA(cudaMemcpyAsync(&flag, l.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    TVS_CUDA(cudaStreamSynchronize(h->stream));
    if (flag & 2) { set_error("%s: negative weight (weights are record counts: w >= 0)", who); return TVS_EINVAL; }
    if (flag & 1) l.cols16_ok = false;
    return TVS_OK;
}

int tvs_columns_upload(tvs_handle h, int64_t first, int64_t n, const int32_t* w_cols) {
    if (!h || !w_cols) { set_error("tvs_columns_upload: null argument"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    int rc = columns_range_ok(h, first, n, "tvs_columns_upload");
    if (rc != TVS_OK) return rc;
    Lanes& l = h->lanes;
    TVS_CUDA(cudaMemcpy2DAsync(l.cols + first, (size_t)l.col_cap * 4, w_cols, (size_t)n * 4, (size_t)n * 4, (size_t)h->csr.R, cudaMemcpyDefault, h->stream));
    return mirror_columns(h, first, n, "tvs_columns_upload");
}

int tvs_columns_upload_u16(tvs_handle h, int64_t first, int64_t n, const uint16_t* w_cols) {
    if (!h || !w_cols) { set_error("tvs_columns_upload_u16: null argument"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    int rc = columns_range_ok(h, first, n, "tvs_columns_upload_u16");
    if (rc != TVS_OK) return rc;
    Lanes& l = h->lanes;
    const int64_t R = h->csr.R, cnt = R * n;
    uint16_t* tmp = nullptr;
    if ((rc = dev_alloc(&tmp, (size_t)cnt)) != TVS_OK) return rc;
    cudaMemcpyAsync(tmp, w_cols, (size_t)cnt * 2, cudaMemcpyDefault, h->stream);
    widen_columns_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(tmp, n, R, l.cols, l.col_cap, first);
    cudaMemcpy2DAsync(l.cols16 + first, (size_t)l.col_cap * 2, tmp, (size_t)n * 2, (size_t)n * 2, (size_t)R, cudaMemcpyDeviceToDevice, h->stream);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    return e == cudaSuccess ? TVS_OK : cuda_fail(e, "tvs_columns_upload_u16", __FILE__, __LINE__);
}

static int resample_into(tvs_problem* h, int32_t* dst, int64_t ld, int64_t n, const int32_t* n_obs, uint64_t seed, int keep_first, uint32_t id0);

int tvs_columns_resample(tvs_handle h, int64_t first, int64_t n, const int32_t* n_obs, uint64_t seed, uint64_t first_id) {
    if (!h || !n_obs) { set_error("tvs_columns_resample: null argument"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    int rc = columns_range_ok(h, first, n, "tvs_columns_resample");
    if (rc != TVS_OK) return rc;
    if (first_id + (uint64_t)n > 0xffffffffull) { set_error("tvs_columns_resample: column ids beyond 2^32"); return TVS_ELIMIT; }
    rc = resample_into(h, h->lanes.cols + first, h->lanes.col_cap, n, n_obs, seed, 0, (uint32_t)first_id);
    if (rc != TVS_OK) return rc;
    return mirror_columns(h, first, n, "tvs_columns_resample");
}

int tvs_set_lanes_columns(tvs_handle h, int64_t B, const int32_t* col_of_lane, const uint8_t* mask, const double* C) {
    if (!h || B <= 0 || !col_of_lane) { set_error("tvs_set_lanes_columns: null argument or B<=0"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    Lanes& l = h->lanes;
    if (!l.cols) { set_error("tvs_set_lanes_columns: no columns registered (tvs_columns_reserve / _upload)"); return TVS_ESTATE; }
    h->lanes_set = false;
    std::vector<int32_t> idx;
    int rc = check_slots(col_of_lane, B, l.col_cap, "tvs_set_lanes_columns", &idx);
    if (rc != TVS_OK) return rc;
    const bool use16 = l.cols16_ok && env_int("TVS_W16", 1) != 0;       // gather the 16-bit mirror: half the traffic, no flag to wait for
    if ((rc = reserve_lanes(h, B, mask != nullptr, C != nullptr, !use16)) != TVS_OK) return rc;
    if (l.cap_slot < B) { dev_free(l.slot); l.cap_slot = 0; if ((rc = dev_alloc(&l.slot, (size_t)B)) != TVS_OK) return rc; l.cap_slot = B; }
    TVS_CUDA(cudaMemcpyAsync(l.slot, idx.data(), (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    const int64_t R = h->csr.R;
    dim3 grid((unsigned)((B + 127) / 128), (unsigned)std::min<int64_t>(R, 4096));
    if (use16) gather_columns_kernel<uint16_t><<<grid, 128, 0, h->stream>>>(l.cols16, l.col_cap, l.slot, R, B, l.w16);
    else gather_columns_kernel<int32_t><<<grid, 128, 0, h->stream>>>(l.cols, l.col_cap, l.slot, R, B, l.w);
    TVS_CUDA(cudaGetLastError());
    // (the slot ids come from a pageable host vector: cudaMemcpyAsync returns once they are staged, so the vector may die here)
    return finish_lanes(h, mask, C, use16);
}

// n bootstrap columns into dst[r*ld + j] (j < n): column j ~ Multinomial(N, n_obs/N) from the Philox stream id0 + j
static int resample_into(tvs_problem* h, int32_t* dst, int64_t ld, int64_t n, const int32_t* n_obs, uint64_t seed, int keep_first, uint32_t id0) {
    const int64_t R = h->csr.R;
    std::vector<int32_t> nn((size_t)R);
    TVS_CUDA(cudaMemcpy(nn.data(), n_obs, (size_t)R * sizeof(int32_t), cudaMemcpyDefault));
    std::vector<int64_t> cum((size_t)R);
    int64_t tot = 0;
    for (int64_t r = 0; r < R; ++r) {
        if (nn[r] < 0) { set_error("resample: negative count at row %lld", (long long)r); return TVS_EINVAL; }
        tot += nn[r]; cum[r] = tot;
    }
    if (tot <= 0) { set_error("resample: no records"); return TVS_EINVAL; }
    int64_t* d_cum = nullptr; int32_t* d_n = nullptr;
    int rc;
    if ((rc = dev_alloc(&d_cum, (size_t)R)) != TVS_OK) return rc;
    if ((rc = dev_alloc(&d_n, (size_t)R)) != TVS_OK) { cudaFree(d_cum); return rc; }
    cudaMemcpyAsync(d_cum, cum.data(), (size_t)R * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_n, nn.data(), (size_t)R * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
    cudaMemset2DAsync(dst, (size_t)ld * 4, 0, (size_t)n * 4, (size_t)R, h->stream);
    const int threads = 256;
    const int bx = (int)std::min<int64_t>((tot + threads - 1) / threads, 64);
    for (int64_t j0 = 0; j0 < n; j0 += 32768) {
        const int64_t nj = std::min<int64_t>(32768, n - j0);
        resample_kernel<<<dim3(bx, (unsigned)nj), threads, 0, h->stream>>>(d_cum, R, tot, ld, seed, keep_first, d_n, dst + j0, id0 + (uint32_t)j0);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);          // the host vectors and the scratch arrays die here
    cudaFree(d_cum); cudaFree(d_n);
    return e == cudaSuccess ? TVS_OK : cuda_fail(e, "resample_kernel", __FILE__, __LINE__);
}

int tvs_resample(tvs_handle h, int64_t B, const int32_t* n_obs, uint64_t seed, int keep_lane0, const uint8_t* mask,
                 const double* C) {
    if (!h || B <= 0 || !n_obs) { set_error("tvs_resample: null handle/n_obs or B<=0"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    h->lanes_set = false;
    int rc = reserve_lanes(h, B, mask != nullptr, C != nullptr, true);
    if (rc != TVS_OK) return rc;
    if ((rc = resample_into(h, h->lanes.w, B, B, n_obs, seed, keep_lane0, 0u)) != TVS_OK) return rc;
    return finish_lanes(h, mask, C, false);
}

int tvs_get_weights(tvs_handle h, int32_t* w_out) {
    if (!h || !w_out) { set_error("tvs_get_weights: null argument"); return TVS_EINVAL; }
    if (!h->lanes_set) { set_error("tvs_get_weights: lanes not set"); return TVS_ESTATE; }
    TVS_CUDA(cudaSetDevice(h->device));
    int rc = ensure_w32(h);
    if (rc != TVS_OK) return rc;
    TVS_CUDA(cudaMemcpyAsync(w_out, h->lanes.w, (size_t)(h->csr.R * h->lanes.B) * sizeof(int32_t), cudaMemcpyDefault, h->stream));
    TVS_CUDA(cudaStreamSynchronize(h->stream));
    return TVS_OK;
}

int tvs_loglik(tvs_handle h, const double* theta, double* ll_out) {
    if (!h || !theta || !ll_out) { set_error("tvs_loglik: null argument"); return TVS_EINVAL; }
    if (!h->lanes_set) { set_error("tvs_loglik: call tvs_set_lanes first"); return TVS_ESTATE; }
    TVS_CUDA(cudaSetDevice(h->device));
    const int64_t V = h->csr.V, B = h->lanes.B;
    PassPlan pl;
    int rc = make_plan<false>(h, B, &pl);
    if (rc != TVS_OK) return rc;
    // one pooled scratch block per handle (was five cudaMalloc / cudaFree per call)
    const size_t need = (size_t)(2 * V * B) + (size_t)B * (size_t)(pl.splits + 2);
    if (h->ll_scratch_cap < need) {
        if (h->ll_scratch) cudaFree(h->ll_scratch);
        h->ll_scratch = nullptr; h->ll_scratch_cap = 0;
        TVS_CUDA(cudaMalloc((void**)&h->ll_scratch, need * sizeof(double)));
        h->ll_scratch_cap = need;
    }
    double* d_theta = h->ll_scratch; double* d_x = d_theta + V * B; double* d_den = d_x + V * B; double* d_out = d_den + B;
    double* d_llp = d_out + B;
    auto cleanup = [&]() {};
    cudaMemcpyAsync(d_theta, theta, (size_t)(V * B) * sizeof(double), cudaMemcpyDefault, h->stream);
    const unsigned gb = (unsigned)((B + 127) / 128);
    prep_theta_kernel<<<gb, 128, 0, h->stream>>>(d_theta, h->lanes.mask, h->csr.Ld, (int)V, B, d_x, d_den);
    PassArgs a = base_args(h);
    a.x = d_x; a.done = nullptr; a.tpart = nullptr; a.llpart = d_llp;
    rc = launch_pass<false>(h, pl, a);
    if (rc == TVS_OK) {
        finish_loglik_kernel<<<gb, 128, 0, h->stream>>>(d_llp, pl.splits, B, h->lanes.N, h->lanes.C, d_den, d_out);
        cudaMemcpyAsync(ll_out, d_out, (size_t)B * sizeof(double), cudaMemcpyDefault, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "tvs_loglik", __FILE__, __LINE__);
    }
    cleanup();
    return rc;
}

int tvs_fit(tvs_handle h, const double* theta0, const tvs_fit_options* opt_in, double* theta_out, double* ll_out,
            double* kkt_out, int32_t* iters_out, int32_t* status_out) {
    if (!h) { set_error("tvs_fit: null handle"); return TVS_EINVAL; }
    if (!h->lanes_set) { set_error("tvs_fit: call tvs_set_lanes or tvs_resample first"); return TVS_ESTATE; }
    TVS_CUDA(cudaSetDevice(h->device));
    tvs_fit_options opt{};
    if (opt_in) opt = *opt_in;
    if (!(opt.tol > 0.0)) opt.tol = 1e-9;
    if (opt.max_iter <= 0) opt.max_iter = 1000;
    if (opt.history <= 0) opt.history = 16;
    opt.history = std::min(opt.history, kMaxHistory);
    opt.history = (opt.history <= 4) ? 4 : (opt.history <= 8) ? 8 : (opt.history <= 12) ? 12 : 16;   // instantiations of the update kernel
    if (opt.check_every <= 0) opt.check_every = 4;
    const bool time_passes = (opt.reserved & 1) != 0 || env_int("TVS_TIME_PASSES", 0) != 0;
    const bool no_compact = (opt.reserved & 2) != 0 || env_int("TVS_NO_COMPACT", 0) != 0;

    const int64_t V = h->csr.V, B0 = h->lanes.B, R = h->csr.R;
    const int m = opt.history;
    int rc;
    if ((rc = ensure_fit_state(h, m)) != TVS_OK) return rc;
    FitState& f = h->fs;
    cudaStream_t st = h->stream;
    const size_t vb0 = (size_t)V * B0;

    if (!h->attrs_set) {                                           // once per handle (the attribute is per device)
        TVS_CUDA(cudaFuncSetAttribute(lbfgs_update_vl2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Vl2Layout<4>::doubles * 8)));
        TVS_CUDA(cudaFuncSetAttribute(lbfgs_update_vl2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Vl2Layout<8>::doubles * 8)));
        TVS_CUDA(cudaFuncSetAttribute(lbfgs_update_vl2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Vl2Layout<16>::doubles * 8)));
        TVS_CUDA(cudaFuncSetAttribute(lbfgs_update_vl2_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Vl2Layout<12>::doubles * 8)));
        h->attrs_set = true;
    }
    TVS_CUDA(cudaEventRecord(h->ev0, st));
    int cur = 0;
    {   // ---- reset set 0 (full width, views the caller's lanes)
        LaneSet& s = f.set[0];
        s.mask = h->lanes.mask; s.N = h->lanes.N; s.C = h->lanes.C;
        s.w16 = h->lanes.has16 ? h->lanes.w16 : nullptr; s.w = h->lanes.has16 ? nullptr : h->lanes.w;   // the 16-bit form when the lanes carry it
        // The 2- and 4-lane pass kernels load lane pairs / quads as vectors and need a width that is a multiple of their
        // lane count (a 2,063-lane batch of a selection round ran 1.65x slower per lane than a 2,118-lane one).  A batch
        // whose lane count is not a multiple of 4 is therefore re-laid once at the next multiple: its inputs are copied
        // with the row stride A0, the 1..3 extra lanes are zeros, flagged done, without an original id -- the padding
        // lanes a compaction makes.
        const bool pad0 = B0 % 4 != 0 && B0 >= env_int("TVS_PAD_MIN_LANES", 6) && env_int("TVS_PAD", 1) != 0;
        const int64_t A0 = pad0 ? ((B0 + 3) & ~(int64_t)3) : B0;
        if (pad0) {
            const size_t elem = s.w16 ? 2 : 4;
            auto grow = [&](void** ptr, size_t* cap, size_t bytes) -> int {
                if (*cap >= bytes) return TVS_OK;
                if (*ptr) cudaFree(*ptr);
                *ptr = nullptr; *cap = 0;
                TVS_CUDA(cudaMalloc(ptr, bytes));
                *cap = bytes;
                return TVS_OK;
            };
            if ((rc = grow(&f.pad_w, &f.pad_w_cap, (size_t)R * A0 * elem)) != TVS_OK) return rc;
            if (s.mask && (rc = grow((void**)&f.pad_mask, &f.pad_mask_cap, (size_t)V * A0)) != TVS_OK) return rc;
            if (f.pad_b_cap < (size_t)A0) {
                if (f.pad_N) cudaFree(f.pad_N);
                if (f.pad_C) cudaFree(f.pad_C);
                f.pad_N = f.pad_C = nullptr; f.pad_b_cap = 0;
                TVS_CUDA(cudaMalloc((void**)&f.pad_N, (size_t)A0 * 8));
                TVS_CUDA(cudaMalloc((void**)&f.pad_C, (size_t)A0 * 8));
                f.pad_b_cap = (size_t)A0;
            }
            iota_kernel<<<(unsigned)((B0 + 255) / 256), 256, 0, st>>>(f.idx, B0);
            cudaMemsetAsync(f.idx + B0, 0xFF, (size_t)(A0 - B0) * 4, st);           // -1: a column of zeros
            GatherArgs ga{};
            int64_t row0 = 0;
            auto add = [&](const void* src, void* dst, int64_t rows, int el) {
                if (!src || !dst || rows <= 0) return;
                ga.t[ga.n].src = src; ga.t[ga.n].dst = dst; ga.t[ga.n].rows = rows; ga.t[ga.n].elem = el; ga.t[ga.n].row0 = (int32_t)row0;
                row0 += rows; ++ga.n;
            };
            add(s.w16 ? (const void*)s.w16 : (const void*)s.w, f.pad_w, R, (int)elem);
            add(s.mask, f.pad_mask, V, 1); add(s.N, f.pad_N, 1, 8); add(s.C, f.pad_C, 1, 8);
            ga.total_rows = row0; ga.Bs = B0; ga.Bd = A0; ga.idx = f.idx;
            dim3 grid((unsigned)((A0 + 127) / 128), (unsigned)std::min<int64_t>(row0, 8192));
            gather_all_kernel<<<grid, 128, 0, st>>>(ga);
            if (s.w16) s.w16 = (const uint16_t*)f.pad_w; else s.w = (const int32_t*)f.pad_w;
            if (s.mask) s.mask = f.pad_mask;
            s.N = f.pad_N;
            if (s.C) s.C = f.pad_C;
        }
        s.width = A0;
        const size_t vbA = (size_t)V * A0;
        FillArgs fa{};
        auto fill = [&](void* ptr, size_t bytes, uint32_t v) { fa.t[fa.n].p = ptr; fa.t[fa.n].bytes = bytes; fa.t[fa.n].fill = v; ++fa.n; };
        fill(s.z, vbA * 8, 0); fill(s.g, vbA * 8, 0); fill(s.d, vbA * 8, 0); fill(f.gt, vbA * 8, 0);
        fill(s.F, A0 * 8, 0); fill(s.dg, A0 * 8, 0); fill(s.alpha, A0 * 8, 0); fill(s.llsum, A0 * 8, 0); fill(s.kkt, 2 * A0 * 8, 0);
        fill(s.sumphi, A0 * 8, 0); fill(s.gamma, A0 * 8, 0); fill(s.rho, (size_t)A0 * m * 8, 0);
        fill(s.gram, (size_t)A0 * ((2 * m + 1) * (2 * m + 2) / 2) * 8, 0);
        fill(s.S, vbA * m * 8, 0); fill(s.Y, vbA * m * 8, 0);
        fill(s.ls, A0 * 4, 0); fill(s.iters, A0 * 4, 0xFFFFFFFFu); fill(s.status, A0 * 4, 0);
        fill(s.head, A0 * 4, 0); fill(s.cnt, A0 * 4, 0); fill(s.done, A0 * 4, 0);
        {
            const size_t big = vbA * m * 8 / 16;                  // 16-byte words of the largest array
            const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((big + 255) / 256, (size_t)h->sm_count * 4));
            fill_many_kernel<<<dim3(gx, (unsigned)fa.n), 256, 0, st>>>(fa);
        }
        iota_kernel<<<(unsigned)((B0 + 255) / 256), 256, 0, st>>>(s.orig, B0);
        if (A0 > B0) {
            cudaMemsetAsync(s.orig + B0, 0xFF, (size_t)(A0 - B0) * 4, st);
            cudaMemsetAsync(s.done + B0, 1, (size_t)(A0 - B0) * 4, st);
        }
        const double* d_theta0 = nullptr;
        if (theta0) { cudaMemcpyAsync(f.theta_io, theta0, vb0 * 8, cudaMemcpyDefault, st); d_theta0 = f.theta_io; }
        init_z_kernel<<<(unsigned)((A0 + 127) / 128), 128, 0, st>>>(d_theta0, s.mask, h->csr.Ld, h->csr.invL, (int)V, A0, B0, s.zt, s.x);
    }

    std::vector<cudaEvent_t> pev, uev; // per-pass / per-update events when timing is requested
    std::vector<int64_t> pass_width;   // width of every pass (TVS_DUMP_PASSES)
    std::vector<cudaEvent_t> nev;      // Newton phase: (before hess, after hess, after solve) per iteration
    int64_t passes = 0, launches = 3, lane_passes = 0, n_polls = 0, compactions = 0;
    const int64_t max_passes = (int64_t)opt.max_iter * 2 + 64;
    cudaError_t cerr = cudaSuccess;
    bool all_done = false;

    auto finalize = [&](const LaneSet& s, int only_done) {
        FinArgs fa{};
        fa.z = s.z; fa.invL = h->csr.invL; fa.llsum = s.llsum; fa.sumphi = s.sumphi; fa.N = s.N; fa.C = s.C; fa.kkt = s.kkt;
        fa.iters = s.iters; fa.status = s.status; fa.done = s.done; fa.orig = s.orig;
        fa.theta_out = theta_out ? f.theta_io : nullptr; fa.ll_out = ll_out ? f.ll_out : nullptr;
        fa.kkt_out = kkt_out ? f.kkt_out : nullptr; fa.iters_out = iters_out ? f.iters_out : nullptr;
        fa.status_out = status_out ? f.status_out : nullptr;
        fa.V = (int)V; fa.B = s.width; fa.B0 = B0; fa.only_done = only_done;
        finalize_fit_kernel<<<(unsigned)((s.width + 127) / 128), 128, 0, st>>>(fa);
        ++launches;
    };

    // ---- damped Newton phase (tvs_newton.cuh): runs the lanes of a small set to convergence
    // Newton takes over once at most this many lanes are still iterating: 512 with the blocked Hessian product (one launch,
    // 0.6 ms per 1,000 lanes: cfg2 11.7 -> 11.1 ms per 1000-lane fit), 128 with the round-1 Hessian kernels
    const int newton_lanes = (h->csr.slab_ptr != nullptr && V <= kNewtonMaxV) ? env_int("TVS_NEWTON_LANES", h->csr.has_blk_h ? 512 : 128) : 0;
    int64_t newton_iters = 0;
    const bool newton_trace = env_int("TVS_NEWTON_TRACE", 0) != 0;
    bool newton_off = false;                                    // set when the shape has no staged pass kernel
    // One SEGMENT of the Newton phase on lane set `s`: iterates until every lane has finished (all_done) or so many
    // have that a compaction pays (`*active_out` <= width / 2); the caller compacts and calls again.
    auto newton_phase = [&](LaneSet& s, bool* ran, int64_t* active_out) -> int {
        *ran = false;
        const int64_t W = s.width;
        const int NP = (int)(V * (V + 1) / 2);
        const Csr& c = h->csr;
        // Hessian as a blocked product (blk_kernel, backward mode) when its matrix was built; else the round-1 kernels
        BlkOne hb;
        const bool hblk = c.has_blk_h && env_int("TVS_HESS_BLK", 1) != 0 && W >= env_int("TVS_HESS_BLK_MIN", 16) &&
                          plan_blk_backward(h, W, c.blk_h, &hb) == TVS_OK;
        const bool wide = !hblk && c.wide_ptr != nullptr && W >= env_int("TVS_HESS_WIDE_MIN", 256) && W % 4 == 0;
        // slab kernel: one CTA per (lane, range of row blocks), all CTAs resident at once
        const size_t hess_smem = ((size_t)NP + c.h_HB + 1 + c.h_NG + 1) * 8 + (size_t)c.h_NG * 32 * 2;
        const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(220 * 1024) / (hess_smem + 1024)));
        int S_h = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)h->sm_count * ctas_per_sm / W, c.h_nblk));
        const int blocks_per_split = (c.h_nblk + S_h - 1) / S_h;
        S_h = (c.h_nblk + blocks_per_split - 1) / blocks_per_split;
        const size_t wide_smem = (size_t)2 * c.w_HB * 32 * sizeof(float) + (size_t)2 * c.w_cap * 8 + (size_t)(2 * (kWideWarps + 1) + 2) * 4 +
                                 (size_t)kWideWarps * kWideEnt * 32 * sizeof(float);
        const size_t need_cap = hblk ? (size_t)hb.splits * (size_t)W * (size_t)NP : wide ? (size_t)W * (size_t)NP : (size_t)S_h * (size_t)W * (size_t)NP;
        if (f.Hpart_cap < need_cap) {
            if (f.Hpart) cudaFree(f.Hpart);
            f.Hpart = nullptr; f.Hpart_cap = 0;
            TVS_CUDA(cudaMalloc((void**)&f.Hpart, need_cap * sizeof(double)));
            f.Hpart_cap = need_cap;
        }
        if (!wide && f.cr_cap < (size_t)R * (size_t)W) {
            if (f.cr) cudaFree(f.cr);
            f.cr = nullptr; f.cr_cap = 0;
            TVS_CUDA(cudaMalloc((void**)&f.cr, (size_t)R * (size_t)W * sizeof(double)));
            f.cr_cap = (size_t)R * (size_t)W;
        }
        if (wide && f.crf_cap < (size_t)R * (size_t)W) {
            if (f.crf) cudaFree(f.crf);
            f.crf = nullptr; f.crf_cap = 0;
            TVS_CUDA(cudaMalloc((void**)&f.crf, (size_t)R * (size_t)W * sizeof(float)));
            f.crf_cap = (size_t)R * (size_t)W;
        }
        TVS_CUDA(cudaFuncSetAttribute(hess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hess_smem));
        TVS_CUDA(cudaFuncSetAttribute(hess_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem));
        const size_t solve_smem = ((size_t)V * (V + 1) + 2 * (V + 2) + 3 * V) * 8;
        TVS_CUDA(cudaFuncSetAttribute(solve_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem));
        TVS_CUDA(cudaFuncSetAttribute(solve_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem));
        PassPlan pl;
        int prc = make_plan<true>(h, W, &pl);
        if (prc != TVS_OK) return prc;
        if (pl.v3_warps == 0) { newton_off = true; return TVS_OK; }   // no staged pass kernel for this shape: stay with L-BFGS
        *ran = true;
        while (pl.splits > 1 && (size_t)pl.splits * (size_t)W * (size_t)V > f.tpart_cap) --pl.splits;
        const bool pre_reduce = pl.splits > env_int("TVS_FUSE_SPLITS", 24);
        newton_init_kernel<<<(unsigned)((W + 127) / 128), 128, 0, st>>>(s.z, s.g, s.mask, h->csr.invL, (int)V, W, f.gha, s.x, f.lam, f.need, s.ls);
        ++launches;
        PassArgs pa = base_args(h);
        pa.w = s.w; pa.w16 = s.w16; pa.B = W; pa.x = s.x; pa.done = s.done; pa.tpart = f.tpart; pa.llpart = f.llpart;
        pa.cr = wide ? nullptr : f.cr; pa.crf = wide ? f.crf : nullptr;
        {   // refresh pass at the accepted point: c_r = w_r / p_r^2 for the first Hessian
            int lrc = launch_pass<true>(h, pl, pa);
            if (lrc != TVS_OK) return lrc;
            ++launches;
        }
        HessArgs ha{};
        ha.slab_ptr = c.slab_ptr; ha.h_row = c.h_row; ha.h_val = c.h_val; ha.h_slot = c.h_slot; ha.cr = f.cr;
        ha.done = s.done; ha.need = f.need; ha.Hpart = f.Hpart;
        ha.W = W; ha.R = R; ha.V = (int)V; ha.HB = c.h_HB; ha.nblk = c.h_nblk; ha.NG = c.h_NG; ha.blocks_per_split = blocks_per_split;
        HessWideArgs hw{};
        hw.wide_ptr = c.wide_ptr; hw.wide_rec = (const uint2*)c.wide_rec; hw.crf = f.crf; hw.H = f.Hpart;
        hw.W = W; hw.R = R; hw.V = (int)V; hw.HB = c.w_HB; hw.nblk = c.w_nblk; hw.cap = c.w_cap;
        const dim3 wide_grid((unsigned)((W + 31) / 32), (unsigned)((NP + kWideSet - 1) / kWideSet));
        BlkArgs hbk{};                                           // blocked Hessian product: H[(u,v), lane] = sum_r G[(u,v), r] * cr[r, lane]
        hbk.m.seg_off = c.blk_h.seg_off; hbk.m.blob = (const uint4*)c.blk_h.blob; hbk.m.n_chunks = c.blk_h.n_chunks;
        hbk.m.n_panels = c.blk_h.n_panels; hbk.m.max_seg_bytes = c.blk_h.max_seg_bytes; hbk.m.dealt = c.blk_h.dealt;
        hbk.src = f.cr; hbk.n_minor = R; hbk.n_major = NP; hbk.W = W; hbk.done = s.done; hbk.need = f.need;
        hbk.tpart = f.Hpart; hbk.stages = hb.stages; hbk.logc = pa.logc;
        const dim3 hblk_grid((unsigned)c.blk_h.n_chunks, (unsigned)hb.tiles, (unsigned)hb.splits);
        SolveArgs sa{};
        sa.Hpart = f.Hpart; sa.S = hblk ? hb.splits : wide ? 1 : S_h;
        sa.h_entry = (wide || hblk) ? W : 1; sa.h_lane = (wide || hblk) ? 1 : NP;
        sa.h_split = hblk ? (int64_t)W * NP : wide ? 0 : (int64_t)W * NP;
        sa.z = s.z; sa.g = s.g; sa.gha = f.gha; sa.invL = h->csr.invL; sa.N = s.N; sa.mask = s.mask;
        sa.done = s.done; sa.lam = f.lam; sa.d = s.d; sa.zt = s.zt; sa.x = s.x; sa.dg = s.dg; sa.alpha = s.alpha; sa.W = W; sa.V = (int)V;
        sa.sumphi = s.sumphi; sa.tol_dual = (double)env_int("TVS_DUAL_MULT", 100) * opt.tol; sa.ls = s.ls;
        NewtonUpdArgs na{};
        na.invL = h->csr.invL; na.N = s.N; na.mask = s.mask; na.z = s.z; na.zt = s.zt; na.g = s.g; na.gt = f.gt; na.gh = f.gh;
        na.gha = f.gha; na.d = s.d; na.F = s.F; na.dg = s.dg; na.alpha = s.alpha; na.llsum = s.llsum; na.kkt = s.kkt;
        na.sumphi = s.sumphi; na.lam = f.lam; na.ls = s.ls; na.iters = s.iters; na.status = s.status; na.done = s.done;
        na.need = f.need; na.V = (int)V; na.B = W; na.max_iter = opt.max_iter; na.tol = opt.tol;
        na.tol_dual = (double)env_int("TVS_DUAL_MULT", 100) * opt.tol;
        const int max_newton = env_int("TVS_NEWTON_MAXIT", 300);
        for (int it = 0; it < max_newton && passes < max_passes; ++it) {
            const int slot = (int)(n_polls % 16);
            cudaEvent_t eh0 = nullptr, eh1 = nullptr, eh2 = nullptr;
            if (time_passes) { cudaEventCreate(&eh0); cudaEventCreate(&eh1); cudaEventCreate(&eh2); cudaEventRecord(eh0, st); }
            if (hblk) {
                if (hb.K == 2) blk_launch_one<2, true>(hbk, 1, hblk_grid, hb.smem, st);
                else if (hb.bulk) blk_launch_one<1, true>(hbk, 1, hblk_grid, hb.smem, st);
                else blk_launch_one<1, false>(hbk, 1, hblk_grid, hb.smem, st);
            } else if (wide) hess_wide_kernel<<<wide_grid, 32 * kWideWarps, wide_smem, st>>>(hw);
            else hess_kernel<<<dim3((unsigned)W, (unsigned)S_h), 32 * kHessWarps, hess_smem, st>>>(ha);
            if (time_passes) cudaEventRecord(eh1, st);
            if (V <= 64) solve_kernel<64><<<(unsigned)W, kSolveThreads, solve_smem, st>>>(sa);
            else solve_kernel<128><<<(unsigned)W, kSolveThreads, solve_smem, st>>>(sa);
            if (time_passes) { cudaEventRecord(eh2, st); nev.push_back(eh0); nev.push_back(eh1); nev.push_back(eh2); }
            if (time_passes) { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); pev.push_back(e0); pev.push_back(e1); cudaEventRecord(e0, st); }
            int lrc = launch_pass<true>(h, pl, pa);
            if (lrc != TVS_OK) return lrc;
            if (time_passes) cudaEventRecord(pev.back(), st);
            if (pre_reduce) {
                reduce_partials_kernel<<<(unsigned)(((int64_t)(V + 1) * W + 255) / 256), 256, 0, st>>>(f.tpart, f.llpart, pl.splits, pl.splits, (int)V, W,
                                                                                                      f.t_red, f.ll_red);
                na.tpart = f.t_red; na.llpart = f.ll_red; na.S = 1;
                ++launches;
            } else { na.tpart = f.tpart; na.llpart = f.llpart; na.S = pl.splits; }
            na.nactive = f.nactive + slot;
            cudaMemsetAsync(f.nactive + slot, 0, 4, st);
            newton_update_kernel<<<(unsigned)((W + kU2Lanes - 1) / kU2Lanes), kU2Threads, 0, st>>>(na);
            if (time_passes) { cudaEvent_t e2; cudaEventCreate(&e2); uev.push_back(pev.back()); uev.push_back(e2); cudaEventRecord(e2, st); pass_width.push_back(-W); }
            ++passes; launches += 5; lane_passes += W; ++newton_iters;
            cudaMemcpyAsync(f.h_nactive + slot, f.nactive + slot, 4, cudaMemcpyDeviceToHost, st);
            cerr = cudaStreamSynchronize(st);
            if (cerr != cudaSuccess) return TVS_OK;       // reported by the caller through cerr
            ++n_polls;
            *active_out = f.h_nactive[slot];
            if (newton_trace && W <= 8) {                         // developer trace: per-lane damping and residuals
                double lam[8], kkt[16], F[8]; int32_t ls[8], dn[8];
                cudaMemcpy(lam, f.lam, W * 8, cudaMemcpyDeviceToHost);
                cudaMemcpy(kkt, s.kkt, W * 16, cudaMemcpyDeviceToHost);
                cudaMemcpy(F, s.F, W * 8, cudaMemcpyDeviceToHost);
                cudaMemcpy(ls, s.ls, W * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(dn, s.done, W * 4, cudaMemcpyDeviceToHost);
                for (int64_t b = 0; b < W; ++b)
                    if (!dn[b] || it < 2)
                        fprintf(stderr, "newton it %d lane %lld done %d ls %d lam %.3e comp %.3e dual %.3e F %.17g\n", it,
                                (long long)b, dn[b], ls[b], lam[b], kkt[b], kkt[W + b], F[b]);
            }
            if (*active_out == 0) { all_done = true; break; }
            if (!no_compact && *active_out * 2 <= W && W > 32 && *active_out + 3 <= f.half) break;   // the caller compacts
        }
        return TVS_OK;
    };

    bool in_newton = false;
    int64_t newton_active = 0;
    // compact when at most this fraction of the set's lanes is still iterating (one gather launch per compaction; the
    // buffers are sized for 75 %).  Measured, fit time at 50 / 60-65 / 70 / 75 %: cfg2 1,000 lanes 11.64 / 11.53 / 11.52 / 11.54 ms,
    // cfg3 2,048 lanes 438 / 430 / 424 / 427 ms, cfg4 1,248 lanes 4.31 / 4.14 / - / 4.07 s (tools/compact_sweep.py)
    const double compact_frac = std::min(0.75, std::max(0.1, env_int("TVS_COMPACT_PCT", 70) / 100.0));
    while (!all_done && passes < max_passes && rc == TVS_OK) {
        LaneSet& s = f.set[cur];
        const int64_t W = s.width;
        int64_t active = newton_active;                        // Newton mode: the count of the last segment
        if (!in_newton) {
            // ---- launch geometry for the current width
            PassPlan pl;
            {
                rc = make_plan<true>(h, W, &pl);
                if (rc != TVS_OK) break;
                int& tsplits = pl.blk ? pl.bsplits : pl.splits;          // splits of the t partials
                while (tsplits > 1 && (size_t)tsplits * (size_t)W * (size_t)V > f.tpart_cap) --tsplits;
                if ((size_t)tsplits * (size_t)W * (size_t)V > f.tpart_cap) { set_error("partial buffer too small"); rc = TVS_ENOMEM; break; }
                while (pl.splits > 1 && (size_t)pl.splits * (size_t)W > f.llpart_cap) --pl.splits;
            }
            PassArgs pa = base_args(h);
            pa.w = s.w; pa.w16 = s.w16; pa.B = W; pa.x = s.x; pa.done = s.done; pa.tpart = f.tpart; pa.llpart = f.llpart;
            UpdArgs ua{};
            ua.invL = h->csr.invL; ua.N = s.N; ua.mask = s.mask;
            ua.z = s.z; ua.zt = s.zt; ua.g = s.g; ua.gt = f.gt; ua.gh = f.gh; ua.d = s.d; ua.x = s.x; ua.Sh = s.S; ua.Yh = s.Y;
            ua.rho = s.rho; ua.gamma = s.gamma; ua.F = s.F; ua.dg = s.dg; ua.alpha = s.alpha; ua.llsum = s.llsum; ua.kkt = s.kkt;
            ua.sumphi = s.sumphi; ua.ls = s.ls; ua.iters = s.iters; ua.status = s.status; ua.head = s.head; ua.cnt = s.cnt;
            ua.done = s.done; ua.V = (int)V; ua.B = W; ua.m = m; ua.max_iter = opt.max_iter; ua.tol = opt.tol;
            ua.tol_dual = (double)env_int("TVS_DUAL_MULT", 100) * opt.tol;
    
            // ---- a batch of passes, then one synchronous poll of the device-side "still iterating" counter
            const int slot = (int)(n_polls % 16);
            // the first poll comes later: no lane of a cold start converges within a dozen L-BFGS iterations, and a poll
            // costs a stream synchronisation plus the relaunch gap (kernels of finished lane sets exit at once anyway)
            const int group = (passes == 0 && !theta0) ? std::max(opt.check_every, env_int("TVS_FIRST_GROUP", 12)) : opt.check_every;
            for (int k = 0; k < group && passes < max_passes; ++k) {
                if (!pl.xs) { cudaMemsetAsync(f.tpart, 0, (size_t)V * W * pl.splits * 8, st); ++launches; }
                if (time_passes) { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); pev.push_back(e0); pev.push_back(e1); cudaEventRecord(e0, st); }
                rc = launch_pass<true>(h, pl, pa);
                if (rc != TVS_OK) break;
                if (time_passes) cudaEventRecord(pev.back(), st);
                const bool last = (k == group - 1) || (passes + 1 == max_passes);
                const int t_splits = pl.blk ? pl.bsplits : pl.splits;               // blk: t and ll partials have their own split counts
                const bool pre_reduce = std::max(t_splits, pl.splits) > env_int("TVS_FUSE_SPLITS", 24);   // many splits: sum them with full parallelism first
            if (pre_reduce) {
                reduce_partials_kernel<<<(unsigned)(((int64_t)(V + 1) * W + 255) / 256), 256, 0, st>>>(f.tpart, f.llpart, t_splits, pl.splits, (int)V, W,
                                                                                                      f.t_red, f.ll_red);
                ua.tpart = f.t_red; ua.llpart = f.ll_red; ua.S = 1; ua.S_ll = 1;
                ++launches;
            } else {
                ua.tpart = f.tpart; ua.llpart = f.llpart; ua.S = t_splits; ua.S_ll = pl.splits;
            }
                if (pl.blk) ++launches;                                              // forward + backward
            // the poll counter of a group is zeroed by the group's first update launch (or a memset when the group is one pass)
            ua.nactive = last ? f.nactive + slot : f.nactive + 31;   // word 31 = scratch between polls
            ua.nreset = (k == 0 && !last) ? f.nactive + slot : nullptr;
            if (k == 0 && last) { cudaMemsetAsync(f.nactive + slot, 0, 4, st); ++launches; }
            const unsigned ub2 = (unsigned)((W + kU2Lanes - 1) / kU2Lanes);
            if (m == 4) lbfgs_update_vl2_kernel<4><<<ub2, kU2Threads, Vl2Layout<4>::doubles * 8, st>>>(ua, s.gram);
            else if (m == 8) lbfgs_update_vl2_kernel<8><<<ub2, kU2Threads, Vl2Layout<8>::doubles * 8, st>>>(ua, s.gram);
            else if (m == 12) lbfgs_update_vl2_kernel<12><<<ub2, kU2Threads, Vl2Layout<12>::doubles * 8, st>>>(ua, s.gram);
            else lbfgs_update_vl2_kernel<16><<<ub2, kU2Threads, Vl2Layout<16>::doubles * 8, st>>>(ua, s.gram);
            if (time_passes) { cudaEvent_t e2; cudaEventCreate(&e2); uev.push_back(pev.back()); uev.push_back(e2); cudaEventRecord(e2, st); }
                ++passes; launches += 2; lane_passes += W;
                if (time_passes) pass_width.push_back(W);
            }
            if (rc != TVS_OK) break;
            cudaMemcpyAsync(f.h_nactive + slot, f.nactive + slot, 4, cudaMemcpyDeviceToHost, st);
            cerr = cudaStreamSynchronize(st);
            if (cerr != cudaSuccess) break;
            ++n_polls;
            active = f.h_nactive[slot];
            if (active == 0) { all_done = true; break; }
        }
        // ---- compaction: keep only the lanes that are still iterating
        const bool to_newton = in_newton || (newton_lanes > 0 && !newton_off && active <= newton_lanes);
        if (!no_compact && ((double)active <= compact_frac * (double)W || (to_newton && W > newton_lanes)) && W > 32 && active + 3 <= f.half) {
            finalize(s, 1);                                      // results of the finished lanes, at their original ids
            LaneSet& n = f.set[cur ^ 1];
            build_index_kernel<<<1, 1024, 0, st>>>(s.done, W, f.idx, f.nactive + 30);
            const int bi = (int)(compactions & 1);
            // new width: survivors + up to 3 padding lanes (index -1 -> zeros, flagged done, no original id)
            const int64_t A = (active + 3) & ~(int64_t)3;
            {   // compaction buffer of the weights, in the form the lane set carries (allocated on first use)
                cudaError_t ae = cudaSuccess;
                if (s.w16 && !f.wbuf16[bi]) ae = cudaMalloc((void**)&f.wbuf16[bi], (size_t)R * f.half * sizeof(uint16_t));
                if (!s.w16 && !f.wbuf[bi]) ae = cudaMalloc((void**)&f.wbuf[bi], (size_t)R * f.half * sizeof(int32_t));
                if (ae != cudaSuccess) { rc = cuda_fail(ae, "compaction buffer", __FILE__, __LINE__); break; }
            }
            n.w = s.w16 ? nullptr : f.wbuf[bi]; n.w16 = s.w16 ? f.wbuf16[bi] : nullptr;
            n.mask = s.mask ? f.mbuf[bi] : nullptr; n.N = f.nbuf[bi]; n.C = s.C ? f.cbuf[bi] : nullptr;
            {   // one launch gathers every per-lane matrix (the L-BFGS history only while that phase is still running)
                GatherArgs ga{};
                int64_t row0 = 0;
                auto add = [&](const void* src, void* dst, int64_t rows, int elem) {
                    if (!src || !dst || rows <= 0) return;
                    ga.t[ga.n].src = src; ga.t[ga.n].dst = dst; ga.t[ga.n].rows = rows; ga.t[ga.n].elem = elem; ga.t[ga.n].row0 = (int32_t)row0;
                    row0 += rows; ++ga.n;
                };
                const int64_t ng = (int64_t)((2 * m + 1) * (2 * m + 2) / 2);
                if (s.w16) add(s.w16, f.wbuf16[bi], R, 2); else add(s.w, f.wbuf[bi], R, 4);
                add(s.mask, f.mbuf[bi], V, 1); add(s.N, f.nbuf[bi], 1, 8); add(s.C, f.cbuf[bi], 1, 8);
                add(s.orig, n.orig, 1, 4); add(s.z, n.z, V, 8); add(s.zt, n.zt, V, 8); add(s.g, n.g, V, 8); add(s.d, n.d, V, 8); add(s.x, n.x, V, 8);
                if (!in_newton && !to_newton) {
                    add(s.S, n.S, (int64_t)m * V, 8); add(s.Y, n.Y, (int64_t)m * V, 8); add(s.rho, n.rho, m, 8); add(s.gamma, n.gamma, 1, 8);
                    add(s.gram, n.gram, ng, 8); add(s.head, n.head, 1, 4); add(s.cnt, n.cnt, 1, 4);
                }
                add(s.F, n.F, 1, 8); add(s.dg, n.dg, 1, 8); add(s.alpha, n.alpha, 1, 8); add(s.llsum, n.llsum, 1, 8); add(s.sumphi, n.sumphi, 1, 8);
                add(s.kkt, n.kkt, 2, 8); add(s.ls, n.ls, 1, 4); add(s.iters, n.iters, 1, 4);
                ga.total_rows = row0; ga.Bs = W; ga.Bd = A; ga.idx = f.idx;
                dim3 grid((unsigned)((A + 127) / 128), (unsigned)std::min<int64_t>(row0, 8192));
                gather_all_kernel<<<grid, 128, 0, st>>>(ga);
            }
            cudaMemsetAsync(n.done, 0, A * 4, st); cudaMemsetAsync(n.status, 0, A * 4, st);
            if (A > active) {
                cudaMemsetAsync(n.done + active, 1, (A - active) * 4, st);
                cudaMemsetAsync(n.orig + active, 0xFF, (A - active) * 4, st);
            }
            n.width = A;
            launches += 3;
            ++compactions;
            cur ^= 1;
        }
        if (to_newton && !newton_off && (in_newton || f.set[cur].width <= std::max(newton_lanes, 32) + 3)) {
            bool ran = false;
            rc = newton_phase(f.set[cur], &ran, &newton_active);
            if (rc != TVS_OK || cerr != cudaSuccess) break;
            if (!ran) {   // Newton is not available for this shape after all: L-BFGS goes on, from an empty history
                LaneSet& q = f.set[cur];                       // (a compaction on the way here did not carry the history over)
                cudaMemsetAsync(q.cnt, 0, q.width * 4, st); cudaMemsetAsync(q.head, 0, q.width * 4, st);
                cudaMemsetAsync(q.gram, 0, (size_t)q.width * ((2 * m + 1) * (2 * m + 2) / 2) * 8, st);
            }
            in_newton = ran;                                    // next trip: compact the survivors, then another segment
            if (ran && !all_done && newton_active * 2 > f.set[cur].width) break;   // pass / iteration budget exhausted
        }
    }
    if (rc == TVS_OK && cerr == cudaSuccess) {
        finalize(f.set[cur], 0);
        if (theta_out) cudaMemcpyAsync(theta_out, f.theta_io, vb0 * 8, cudaMemcpyDefault, st);
        if (ll_out) cudaMemcpyAsync(ll_out, f.ll_out, (size_t)B0 * 8, cudaMemcpyDefault, st);
        if (kkt_out) cudaMemcpyAsync(kkt_out, f.kkt_out, 2 * (size_t)B0 * 8, cudaMemcpyDefault, st);
        if (iters_out) cudaMemcpyAsync(iters_out, f.iters_out, (size_t)B0 * 4, cudaMemcpyDefault, st);
        if (status_out) cudaMemcpyAsync(status_out, f.status_out, (size_t)B0 * 4, cudaMemcpyDefault, st);
        cudaEventRecord(h->ev1, st);
        cerr = cudaStreamSynchronize(st);
    } else {
        cudaStreamSynchronize(st);
    }
    h->stats = tvs_fit_stats{};
    h->stats.passes = passes; h->stats.kernel_launches = launches; h->stats.lane_passes = lane_passes;
    if (rc == TVS_OK && cerr == cudaSuccess) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        h->stats.total_ms = ms;
        double pm = 0.0;
        for (size_t i = 0; i + 1 < pev.size(); i += 2) { float q = 0.f; if (cudaEventElapsedTime(&q, pev[i], pev[i + 1]) == cudaSuccess) pm += q; }
        h->stats.pass_ms = pm;
        double um = 0.0;
        for (size_t i = 0; i + 1 < uev.size(); i += 2) { float q = 0.f; if (cudaEventElapsedTime(&q, uev[i], uev[i + 1]) == cudaSuccess) um += q; }
        h->stats.update_ms = um;
        h->stats.compactions = compactions;
        if (time_passes && env_int("TVS_DUMP_PASSES", 0)) {
            for (size_t i = 0; i + 1 < pev.size(); i += 2) {
                float q = 0.f, u = 0.f;
                cudaEventElapsedTime(&q, pev[i], pev[i + 1]);
                cudaEventElapsedTime(&u, uev[i], uev[i + 1]);
                fprintf(stderr, "pass %zu width %lld pass_us %.1f update_us %.1f\n", i / 2, (long long)pass_width[i / 2], q * 1e3, u * 1e3);
            }
            for (size_t i = 0; i + 2 < nev.size(); i += 3) {
                float q = 0.f, u = 0.f;
                cudaEventElapsedTime(&q, nev[i], nev[i + 1]);
                cudaEventElapsedTime(&u, nev[i + 1], nev[i + 2]);
                fprintf(stderr, "newton %zu hess_us %.1f solve_us %.1f\n", i / 3, q * 1e3, u * 1e3);
            }
        }
    }
    for (size_t i = 1; i < uev.size(); i += 2) cudaEventDestroy(uev[i]);
    for (auto e : pev) cudaEventDestroy(e);
    for (auto e : nev) cudaEventDestroy(e);
    if (rc != TVS_OK) return rc;
    if (cerr != cudaSuccess) return cuda_fail(cerr, "tvs_fit", __FILE__, __LINE__);
    return TVS_OK;
}

int tvs_last_fit_stats(tvs_handle h, tvs_fit_stats* out) {
    if (!h || !out) { set_error("tvs_last_fit_stats: null argument"); return TVS_EINVAL; }
    *out = h->stats;
    return TVS_OK;
}

int tvs_selftest_math(int device, int64_t n, const double* p, double* log_out, double* rcp_out) {
    if (n <= 0 || !p || !log_out || !rcp_out) { set_error("tvs_selftest_math: bad argument"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(device));
    double *d_p = nullptr, *d_l = nullptr, *d_r = nullptr;
    int rc;
    if ((rc = dev_alloc(&d_p, (size_t)n)) != TVS_OK || (rc = dev_alloc(&d_l, (size_t)n)) != TVS_OK || (rc = dev_alloc(&d_r, (size_t)n)) != TVS_OK) {
        cudaFree(d_p); cudaFree(d_l); cudaFree(d_r); return rc;
    }
    cudaMemcpy(d_p, p, (size_t)n * 8, cudaMemcpyDefault);
    selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_p, n, d_l, d_r, make_log_coef());
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) { cudaMemcpy(log_out, d_l, (size_t)n * 8, cudaMemcpyDefault); cudaMemcpy(rcp_out, d_r, (size_t)n * 8, cudaMemcpyDefault); }
    cudaFree(d_p); cudaFree(d_l); cudaFree(d_r);
    if (e != cudaSuccess) return cuda_fail(e, "selftest_math_kernel", __FILE__, __LINE__);
    return TVS_OK;
}

int tvs_selftest_table_log(tvs_handle h, int64_t n, const double* p, double* log_out) {
    if (!h || n <= 0 || !p || !log_out) { set_error("tvs_selftest_table_log: bad argument"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    double *d_p = nullptr, *d_l = nullptr;
    int rc;
    if ((rc = dev_alloc(&d_p, (size_t)n)) != TVS_OK || (rc = dev_alloc(&d_l, (size_t)n)) != TVS_OK) { cudaFree(d_p); cudaFree(d_l); return rc; }
    cudaMemcpy(d_p, p, (size_t)n * 8, cudaMemcpyDefault);
    selftest_table_log_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_p, n, d_l, (const double2*)h->csr.logtab, make_log_coef());
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) cudaMemcpy(log_out, d_l, (size_t)n * 8, cudaMemcpyDefault);
    cudaFree(d_p); cudaFree(d_l);
    if (e != cudaSuccess) return cuda_fail(e, "selftest_table_log_kernel", __FILE__, __LINE__);
    return TVS_OK;
}

int tvs_measure_fp64_peak(int device, double* tflops_out) {
    if (!tflops_out) { set_error("tflops_out is null"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TVS_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double* d = nullptr;
    TVS_CUDA(cudaMalloc((void**)&d, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0 + rep);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "fp64_peak_kernel", __FILE__, __LINE__); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops_out = best;
    return TVS_OK;
}

int tvs_measure_smem_peak(int device, double* gbs_out) {
    if (!gbs_out) { set_error("gbs_out is null"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TVS_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 2, threads = 512, iters = 1 << 14;
    double* d = nullptr;
    TVS_CUDA(cudaMalloc((void**)&d, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        smem_peak_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "smem_peak_kernel", __FILE__, __LINE__); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 16.0 * 8.0 * (double)iters * blocks * threads;
        best = std::max(best, bytes / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *gbs_out = best;
    return TVS_OK;
}

}  // extern "C"
