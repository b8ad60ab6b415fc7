// libtvs.so -- per-replicate multinomial bin statistics on the device (include/tvs_pool.h).
//
// replaces, for a batch of replicates: Traversome.sample_sub_paths (traversome.py:1636-1699, the regrouping part),
// generate_multinomial_bin_stats (:1305-1384), _identify_bins (:1386-1462), the id-range lookup (:1797-1842),
// get_averaged_possible_alignment_start_points (:1492-1521) / get_num_of_possible_alignment_start_points (:1523-1579) and
// the X < 1 pruning of PathMultinomialModel.get_like_formula (ModelGenerator.py:138-151).
//
// Everything is expressed over the POOL sorted once by (alignment length, record id): a replicate is a multiplicity
// vector c[pos] over that order (a histogram of its draws), so "the replicate's records sorted by length" is the pool order
// itself, an id range of the replicate is a difference of the prefix sums PC[pos] = sum c, and the start-point sum of a bin
//     sum_{records of the range} starts(len, read path)
// -- starts is piecewise linear in len -- is a closed form in PC and PL[pos] = sum c * len with two binary searches for
// the break points, instead of the reference's O(records x read paths of the range) Python loop.  All of it in int64:
// num_possible_X = exact integer sum / record count, one IEEE division, bit-identical to the reference's float.
//
// Kernels (one launch each per batch of replicates):
//   pool_count_kernel    c, per-row counts            thread per draw, integer atomics (order-independent)
//   pool_scan_kernel     PC, PL                       one CTA per replicate, blocked scan
//   pool_ranges_kernel   _identify_bins               one thread per replicate walks the statically sorted range points
//   pool_kept_kernel     ranges that hold a record    thread per (replicate, range): position interval, kept flag
//   pool_rows_kernel     bins -> lane                 thread per (replicate, read path): groups its records by range,
//                                                     num_matched, num_possible_X (closed form), pruning, w / C / N partials
//   pool_reduce_kernel   C, N                         one CTA per replicate, fixed summation order
#include "tvs_internal.h"
#include "../../include/tvs_pool.h"

#include <algorithm>
#include <cmath>
#include <new>
#include <numeric>
#include <vector>

struct tvs_record_pool {
    int device = 0;
    int64_t N = 0, R = 0;
    cudaStream_t stream = nullptr;
    // static device arrays
    int32_t* pos_of = nullptr;      // [N] position of pool record i in the (length, id) order
    int32_t* row_of = nullptr;      // [N] read path of pool record i
    int64_t* len_sorted = nullptr;  // [N]
    int32_t* rp_ptr = nullptr;      // [R+1] positions of the records of every read path ...
    int32_t* rp_pos = nullptr;      // [N]   ... ascending (= by length)
    uint8_t* multi = nullptr; int64_t *inner = nullptr, *full = nullptr, *lroom = nullptr, *rroom = nullptr;   // [R]
    uint8_t* legal = nullptr;       // [R] shortest possible alignment <= longest (:1391-1400 "deleting illegal path")
    int64_t* pt_val = nullptr; int32_t* pt_row = nullptr; uint8_t* pt_type = nullptr;   // [NPt] range points sorted by value
    int64_t NPt = 0;
    // scratch (grown on demand)
    void* scratch = nullptr; size_t scratch_bytes = 0;
};

namespace tvs {

constexpr int64_t kNoMax = INT64_MIN;

__global__ void pool_count_kernel(const int32_t* __restrict__ samples, int64_t n_draw, int64_t B, int64_t N, int64_t R,
                                  const int32_t* __restrict__ pos_of, const int32_t* __restrict__ row_of,
                                  const uint8_t* __restrict__ masked, int32_t* __restrict__ c, int32_t* __restrict__ rowcnt,
                                  int32_t* __restrict__ bad) {
    const int64_t b = blockIdx.y;
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < n_draw; d += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = samples[b * n_draw + d];
        if (i < 0) continue;                                   // no draw (jackknife pad)
        if (i >= N) { atomicOr(bad, 1); continue; }
        const int32_t r = row_of[i];
        if (masked != nullptr && masked[r]) continue;          // read_paths_masked: the record is not part of the replicate
        atomicAdd(c + b * N + pos_of[i], 1);
        atomicAdd(rowcnt + b * R + r, 1);
    }
}

// PC[p] = sum_{q<p} c[q], PL[p] = sum_{q<p} c[q] * len[q]   (p = 0..N)
__global__ void __launch_bounds__(256) pool_scan_kernel(const int32_t* __restrict__ c, const int64_t* __restrict__ len_sorted, int64_t N,
                                                        int64_t* __restrict__ PC, int64_t* __restrict__ PL) {
    __shared__ int64_t sc[256], sl[256];
    const int64_t b = blockIdx.x;
    const int t = threadIdx.x;
    const int64_t seg = (N + 255) / 256;
    const int64_t p0 = ((int64_t)t * seg < N) ? (int64_t)t * seg : N, p1 = (p0 + seg < N) ? p0 + seg : N;
    const int32_t* cb = c + b * N;
    int64_t a = 0, l = 0;
    for (int64_t p = p0; p < p1; ++p) { a += cb[p]; l += (int64_t)cb[p] * len_sorted[p]; }
    sc[t] = a; sl[t] = l;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {                  // inclusive scan of the segment totals
        const int64_t va = (t >= off) ? sc[t - off] : 0, vl = (t >= off) ? sl[t - off] : 0;
        __syncthreads();
        sc[t] += va; sl[t] += vl;
        __syncthreads();
    }
    int64_t ra = sc[t] - a, rl = sl[t] - l;                    // exclusive prefix of this segment
    int64_t* pc = PC + b * (N + 1);
    int64_t* pl = PL + b * (N + 1);
    for (int64_t p = p0; p < p1; ++p) { pc[p] = ra; pl[p] = rl; ra += cb[p]; rl += (int64_t)cb[p] * len_sorted[p]; }
    if (t == 255) { pc[N] = sc[255]; pl[N] = sl[255]; }
}

// _identify_bins (traversome.py:1386-1462) of every replicate: the points (shortest / longest possible alignment of the
// replicate's legal read paths) are walked in increasing order; a start point opens a range (closing an open one just
// before it), an end point closes the current range and opens the next one unless the following point continues it.
__global__ void pool_ranges_kernel(int64_t B, int64_t R, int64_t NPt, const int64_t* __restrict__ pt_val, const int32_t* __restrict__ pt_row,
                                   const uint8_t* __restrict__ pt_type, const int32_t* __restrict__ rowcnt, int64_t MR,
                                   int64_t* __restrict__ rmin, int64_t* __restrict__ rmax, int32_t* __restrict__ nr_out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int32_t* rc = rowcnt + b * R;
    int64_t* mn = rmin + b * MR;
    int64_t* mx = rmax + b * MR;
    int nr = 0;
    bool last_type = false;
    int64_t i = 0;
    while (i < NPt && rc[pt_row[i]] == 0) ++i;
    while (i < NPt) {
        const int64_t val = pt_val[i];
        bool is_start = false, is_end = false;
        int64_t j = i;
        while (j < NPt && pt_val[j] == val) {
            if (rc[pt_row[j]] > 0) { const int ty = pt_type[j]; is_start = is_start || ty < 2; is_end = is_end || ty > 0; }
            ++j;
        }
        int64_t k = j;
        while (k < NPt && rc[pt_row[k]] == 0) ++k;             // next point of a read path the replicate holds
        const bool has_next = k < NPt;
        if (is_start) {
            if (nr > 0 && last_type) mx[nr - 1] = val - 1;
            mn[nr] = val; mx[nr] = kNoMax; ++nr;
            last_type = true;
        }
        if (is_end) {
            mx[nr - 1] = val;
            last_type = false;
            if (has_next && !(pt_val[k] == val + 1 && is_start)) { mn[nr] = val + 1; mx[nr] = kNoMax; ++nr; last_type = true; }
        }
        i = k;
    }
    nr_out[b] = nr;
}

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t v) {   // first p with a[p] >= v
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}
__device__ __forceinline__ int64_t upper_bound_i64(const int64_t* a, int64_t n, int64_t v) {   // first p with a[p] > v
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (a[m] <= v) lo = m + 1; else hi = m; }
    return lo;
}

// position interval [pa, pb) of every range and whether the replicate has a record in it (the two-pointer filter of
// :1440-1462 and the id range of :1809-1842: left_id = PC[pa], right_id = PC[pb] - 1)
__global__ void pool_kept_kernel(int64_t B, int64_t N, int64_t MR, const int32_t* __restrict__ nr, const int64_t* __restrict__ rmin,
                                 const int64_t* __restrict__ rmax, const int64_t* __restrict__ len_sorted, const int64_t* __restrict__ PC,
                                 int32_t* __restrict__ pa_out, int32_t* __restrict__ pb_out) {
    const int64_t b = blockIdx.y;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nr[b]) return;
    const int64_t lo = rmin[b * MR + j], hi = rmax[b * MR + j];
    int64_t pa = 0, pb = 0;
    if (hi != kNoMax && hi >= lo) {
        pa = lower_bound_i64(len_sorted, N, lo);
        pb = upper_bound_i64(len_sorted, N, hi);
        const int64_t* pc = PC + b * (N + 1);
        if (pb <= pa || pc[pb] - pc[pa] <= 0) pb = pa;         // no record of the replicate in the range: not kept
    }
    pa_out[b * MR + j] = (int32_t)pa;
    pb_out[b * MR + j] = (int32_t)pb;
}

struct PoolRowArgs {
    int64_t B, N, R, MR;
    const int32_t* c; const int32_t* rowcnt; const int64_t* PC; const int64_t* PL; const int64_t* len_sorted;
    const int32_t* rp_ptr; const int32_t* rp_pos;
    const uint8_t* multi; const int64_t* inner; const int64_t* full; const int64_t* lroom; const int64_t* rroom; const uint8_t* legal;
    const int32_t* nr; const int64_t* rmin; const int64_t* rmax; const int32_t* pa; const int32_t* pb;
    int32_t* w_out; uint8_t* obs_out; double* Cpart; int64_t* Npart;
};

// sum over the records at positions [pa, pb) of c * starts(len, row)   (get_num_of_possible_alignment_start_points :1523-1579)
__device__ __forceinline__ int64_t start_point_sum(const PoolRowArgs& a, const int64_t* pc, const int64_t* pl, int64_t pa, int64_t pb, int64_t r) {
    const int64_t SC = pc[pb] - pc[pa], SL = pl[pb] - pl[pa];
    if (!a.multi[r]) return (a.full[r] + 1) * SC - SL;         // full_len - len + 1
    const int64_t k = a.inner[r] + 1;                          // m = len - (inner_len + 2) + 1
    int64_t s = SL - k * SC;
    const int64_t rooms[2] = {a.lroom[r], a.rroom[r]};
#pragma unroll
    for (int q = 0; q < 2; ++q) {                              // - max(m - room, 0): the records longer than room + k
        int64_t p = upper_bound_i64(a.len_sorted, a.N, rooms[q] + k);
        p = p < pa ? pa : (p > pb ? pb : p);
        s -= (pl[pb] - pl[p]) - (k + rooms[q]) * (pc[pb] - pc[p]);
    }
    return s;
}

__global__ void __launch_bounds__(128) pool_rows_kernel(const PoolRowArgs a) {
    const int64_t b = blockIdx.y;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    int32_t w = 0;
    double C = 0.0;
    int64_t Nn = 0;
    const bool present = a.rowcnt[b * a.R + r] > 0 && a.legal[r];
    if (present) {
        const int32_t* cb = a.c + b * a.N;
        const int64_t* pc = a.PC + b * (a.N + 1);
        const int64_t* pl = a.PL + b * (a.N + 1);
        const int64_t* mn = a.rmin + b * a.MR;
        const int32_t* pa = a.pa + b * a.MR;
        const int32_t* pb = a.pb + b * a.MR;
        const int nr = a.nr[b];
        int cur = -1;
        int64_t n = 0;
        auto flush = [&]() {
            if (cur >= 0 && n > 0) {
                const int64_t S = start_point_sum(a, pc, pl, pa[cur], pb[cur], r);
                const double X = (double)S / (double)(pc[pb[cur]] - pc[pa[cur]]);   // integer sum / float(record count), :1521
                if (!(X < 1.0)) { w += (int32_t)n; C += log(X) * (double)n; Nn += n; }   // ModelGenerator.py:144 prunes X < 1
            }
        };
        for (int32_t q = a.rp_ptr[r]; q < a.rp_ptr[r + 1]; ++q) {
            const int32_t p = a.rp_pos[q];
            const int32_t cnt = cb[p];
            if (cnt == 0) continue;
            // the range that holds position p: the last one starting at or before its length
            const int64_t L = a.len_sorted[p];
            int lo = 0, hi = nr;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (mn[m] <= L) lo = m + 1; else hi = m; }
            const int j = lo - 1;
            const bool in = j >= 0 && p >= pa[j] && p < pb[j];
            if (!in) continue;                                  // cannot happen for a legal read path; kept for safety
            if (j != cur) { flush(); cur = j; n = 0; }
            n += cnt;
        }
        flush();
    }
    if (a.w_out) a.w_out[r * a.B + b] = w;
    if (a.obs_out) a.obs_out[r * a.B + b] = present ? 1 : 0;
    a.Cpart[b * a.R + r] = C;
    a.Npart[b * a.R + r] = Nn;
}

__global__ void __launch_bounds__(256) pool_reduce_kernel(int64_t R, const double* __restrict__ Cpart, const int64_t* __restrict__ Npart,
                                                          double* __restrict__ C_out, int64_t* __restrict__ N_out) {
    __shared__ double sc[256];
    __shared__ int64_t sn[256];
    const int64_t b = blockIdx.x;
    const int t = threadIdx.x;
    double c = 0.0;
    int64_t n = 0;
    for (int64_t r = t; r < R; r += 256) { c += Cpart[b * R + r]; n += Npart[b * R + r]; }
    sc[t] = c; sn[t] = n;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (t < off) { sc[t] += sc[t + off]; sn[t] += sn[t + off]; }
        __syncthreads();
    }
    if (t == 0) { if (C_out) C_out[b] = sc[0]; if (N_out) N_out[b] = sn[0]; }
}

}  // namespace tvs

using namespace tvs;

extern "C" {

int tvs_pool_create(tvs_pool_handle* out, int device, int64_t n_records, int64_t n_rows, const int64_t* record_ids,
                    const int32_t* row_of, const int64_t* align_len, const uint8_t* multi, const int64_t* inner_len,
                    const int64_t* full_len, const int64_t* left_room, const int64_t* right_room) {
    if (!out) { set_error("tvs_pool_create: out is null"); return TVS_EINVAL; }
    *out = nullptr;
    if (n_records <= 0 || n_rows <= 0 || !record_ids || !row_of || !align_len || !multi || !inner_len || !full_len || !left_room || !right_room) {
        set_error("tvs_pool_create: bad sizes or null arrays (n_records=%lld n_rows=%lld)", (long long)n_records, (long long)n_rows);
        return TVS_EINVAL;
    }
    if (n_records > 2000000000LL || n_rows > 2000000000LL) { set_error("tvs_pool_create: more than 2e9 records / read paths"); return TVS_ELIMIT; }
    const int64_t N = n_records, R = n_rows;
    for (int64_t i = 0; i < N; ++i) {
        if (row_of[i] < 0 || row_of[i] >= R) { set_error("tvs_pool_create: row_of[%lld]=%d out of range", (long long)i, row_of[i]); return TVS_EINVAL; }
        if (i > 0 && record_ids[i] <= record_ids[i - 1]) { set_error("tvs_pool_create: record_ids must be strictly increasing"); return TVS_EINVAL; }
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { set_error("tvs_pool_create: no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e)); return TVS_ECUDA; }
    if (device < 0 || device >= ndev) { set_error("tvs_pool_create: device %d out of range", device); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(device));

    // ---- static tables (host): the (length, id) order, records per read path, range points
    std::vector<int32_t> order((size_t)N);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return align_len[x] < align_len[y]; });   // ids ascending inside a length
    std::vector<int32_t> pos_of((size_t)N), rp_ptr((size_t)R + 1, 0), rp_pos((size_t)N);
    std::vector<int64_t> len_sorted((size_t)N);
    for (int64_t p = 0; p < N; ++p) { pos_of[(size_t)order[(size_t)p]] = (int32_t)p; len_sorted[(size_t)p] = align_len[order[(size_t)p]]; }
    for (int64_t i = 0; i < N; ++i) rp_ptr[(size_t)row_of[i] + 1]++;
    for (int64_t r = 0; r < R; ++r) rp_ptr[(size_t)r + 1] += rp_ptr[(size_t)r];
    {
        std::vector<int32_t> at(rp_ptr.begin(), rp_ptr.end() - 1);
        for (int64_t p = 0; p < N; ++p) rp_pos[(size_t)at[(size_t)row_of[order[(size_t)p]]]++] = (int32_t)p;
    }
    std::vector<uint8_t> legal((size_t)R);
    struct Pt { int64_t val; int32_t row; uint8_t type; };
    std::vector<Pt> pts;
    for (int64_t r = 0; r < R; ++r) {
        const int64_t lo = multi[r] ? inner_len[r] + 2 : 1, hi = full_len[r];      // :1391-1392
        legal[(size_t)r] = lo <= hi;
        if (lo > hi) continue;
        if (lo == hi) pts.push_back({lo, (int32_t)r, 1});
        else { pts.push_back({lo, (int32_t)r, 0}); pts.push_back({hi, (int32_t)r, 2}); }
    }
    std::stable_sort(pts.begin(), pts.end(), [](const Pt& x, const Pt& y) { return x.val < y.val; });
    std::vector<int64_t> pt_val(pts.size() + 1);
    std::vector<int32_t> pt_row(pts.size() + 1);
    std::vector<uint8_t> pt_type(pts.size() + 1);
    for (size_t i = 0; i < pts.size(); ++i) { pt_val[i] = pts[i].val; pt_row[i] = pts[i].row; pt_type[i] = pts[i].type; }

    tvs_record_pool* p = new (std::nothrow) tvs_record_pool();
    if (!p) { set_error("out of host memory"); return TVS_ENOMEM; }
    p->device = device; p->N = N; p->R = R; p->NPt = (int64_t)pts.size();
    auto fail = [&](int code) { tvs_pool_destroy(p); return code; };
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TVS_ECUDA);
#define UPP_(dst, src, n, T)                                                                                   \
    if (cudaMalloc((void**)&p->dst, std::max<size_t>(1, (size_t)(n)) * sizeof(T)) != cudaSuccess) return fail(TVS_ENOMEM); \
    if (cudaMemcpy(p->dst, src, (size_t)(n) * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return fail(TVS_ECUDA);
    UPP_(pos_of, pos_of.data(), N, int32_t) UPP_(row_of, row_of, N, int32_t) UPP_(len_sorted, len_sorted.data(), N, int64_t)
    UPP_(rp_ptr, rp_ptr.data(), R + 1, int32_t) UPP_(rp_pos, rp_pos.data(), N, int32_t)
    UPP_(multi, multi, R, uint8_t) UPP_(inner, inner_len, R, int64_t) UPP_(full, full_len, R, int64_t)
    UPP_(lroom, left_room, R, int64_t) UPP_(rroom, right_room, R, int64_t) UPP_(legal, legal.data(), R, uint8_t)
    UPP_(pt_val, pt_val.data(), pts.size(), int64_t) UPP_(pt_row, pt_row.data(), pts.size(), int32_t) UPP_(pt_type, pt_type.data(), pts.size(), uint8_t)
#undef UPP_
    *out = p;
    return TVS_OK;
}

int tvs_pool_destroy(tvs_pool_handle p) {
    if (!p) return TVS_OK;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (void* q : {(void*)p->pos_of, (void*)p->row_of, (void*)p->len_sorted, (void*)p->rp_ptr, (void*)p->rp_pos, (void*)p->multi, (void*)p->inner,
                    (void*)p->full, (void*)p->lroom, (void*)p->rroom, (void*)p->legal, (void*)p->pt_val, (void*)p->pt_row, (void*)p->pt_type, p->scratch})
        if (q) cudaFree(q);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return TVS_OK;
}

int tvs_pool_lanes(tvs_pool_handle p, int64_t B, int64_t n_draw, const int32_t* samples, const uint8_t* masked_rows,
                   int32_t* w_out, double* C_out, int64_t* N_out, uint8_t* observed_out) {
    if (!p || B <= 0 || n_draw <= 0 || !samples) { set_error("tvs_pool_lanes: null handle/samples or B/n_draw<=0"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(p->device));
    const int64_t N = p->N, R = p->R, MR = 2 * p->NPt + 2;
    cudaStream_t st = p->stream;
    // replicates per chunk: scratch of at most ~2 GB
    const size_t per_rep = (size_t)N * 4 + (size_t)R * 4 + (size_t)(N + 1) * 16 + (size_t)MR * 24 + 4 + (size_t)R * 16 + (size_t)n_draw * 4 +
                           (size_t)R * 5 + 16;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)(((size_t)2 << 30) / per_rep)));
    const size_t need = per_rep * (size_t)chunk + (size_t)R + 1024;
    if (p->scratch_bytes < need) {
        if (p->scratch) cudaFree(p->scratch);
        p->scratch = nullptr; p->scratch_bytes = 0;
        TVS_CUDA(cudaMalloc(&p->scratch, need));
        p->scratch_bytes = need;
    }
    // carve (8-byte aligned pieces first)
    unsigned char* q = (unsigned char*)p->scratch;
    auto take = [&](size_t bytes) { unsigned char* r = q; q += (bytes + 15) & ~(size_t)15; return r; };
    int64_t* PC = (int64_t*)take((size_t)chunk * (N + 1) * 8);
    int64_t* PL = (int64_t*)take((size_t)chunk * (N + 1) * 8);
    int64_t* rmin = (int64_t*)take((size_t)chunk * MR * 8);
    int64_t* rmax = (int64_t*)take((size_t)chunk * MR * 8);
    double* Cpart = (double*)take((size_t)chunk * R * 8);
    int64_t* Npart = (int64_t*)take((size_t)chunk * R * 8);
    double* dC = (double*)take((size_t)chunk * 8);
    int64_t* dN = (int64_t*)take((size_t)chunk * 8);
    int32_t* c = (int32_t*)take((size_t)chunk * N * 4);
    int32_t* rowcnt = (int32_t*)take((size_t)chunk * R * 4);
    int32_t* pa = (int32_t*)take((size_t)chunk * MR * 4);
    int32_t* pb = (int32_t*)take((size_t)chunk * MR * 4);
    int32_t* nr = (int32_t*)take((size_t)chunk * 4);
    int32_t* d_samples = (int32_t*)take((size_t)chunk * n_draw * 4);
    int32_t* dw = (int32_t*)take((size_t)chunk * R * 4);
    uint8_t* dobs = (uint8_t*)take((size_t)chunk * R);
    uint8_t* d_masked = (uint8_t*)take((size_t)R);
    int32_t* bad = (int32_t*)take(16);
    if ((size_t)(q - (unsigned char*)p->scratch) > p->scratch_bytes) {   // the carve uses 16-byte rounding: grow once more if needed
        const size_t need2 = (size_t)(q - (unsigned char*)p->scratch) + 1024;
        cudaFree(p->scratch); p->scratch = nullptr; p->scratch_bytes = 0;
        TVS_CUDA(cudaMalloc(&p->scratch, need2));
        p->scratch_bytes = need2;
        return tvs_pool_lanes(p, B, n_draw, samples, masked_rows, w_out, C_out, N_out, observed_out);
    }
    if (masked_rows) TVS_CUDA(cudaMemcpyAsync(d_masked, masked_rows, (size_t)R, cudaMemcpyDefault, st));
    cudaMemsetAsync(bad, 0, 4, st);
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, B - b0);
        TVS_CUDA(cudaMemcpyAsync(d_samples, samples + b0 * n_draw, (size_t)nb * n_draw * 4, cudaMemcpyDefault, st));
        cudaMemsetAsync(c, 0, (size_t)nb * N * 4, st);
        cudaMemsetAsync(rowcnt, 0, (size_t)nb * R * 4, st);
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_draw + 255) / 256, 1024));
        pool_count_kernel<<<dim3(gx, (unsigned)nb), 256, 0, st>>>(d_samples, n_draw, nb, N, R, p->pos_of, p->row_of, masked_rows ? d_masked : nullptr,
                                                                  c, rowcnt, bad);
        pool_scan_kernel<<<(unsigned)nb, 256, 0, st>>>(c, p->len_sorted, N, PC, PL);
        pool_ranges_kernel<<<(unsigned)((nb + 31) / 32), 32, 0, st>>>(nb, R, p->NPt, p->pt_val, p->pt_row, p->pt_type, rowcnt, MR, rmin, rmax, nr);
        pool_kept_kernel<<<dim3((unsigned)((MR + 127) / 128), (unsigned)nb), 128, 0, st>>>(nb, N, MR, nr, rmin, rmax, p->len_sorted, PC, pa, pb);
        PoolRowArgs ra{};
        ra.B = nb; ra.N = N; ra.R = R; ra.MR = MR; ra.c = c; ra.rowcnt = rowcnt; ra.PC = PC; ra.PL = PL; ra.len_sorted = p->len_sorted;
        ra.rp_ptr = p->rp_ptr; ra.rp_pos = p->rp_pos; ra.multi = p->multi; ra.inner = p->inner; ra.full = p->full; ra.lroom = p->lroom;
        ra.rroom = p->rroom; ra.legal = p->legal; ra.nr = nr; ra.rmin = rmin; ra.rmax = rmax; ra.pa = pa; ra.pb = pb;
        ra.w_out = dw; ra.obs_out = dobs; ra.Cpart = Cpart; ra.Npart = Npart;
        pool_rows_kernel<<<dim3((unsigned)((R + 127) / 128), (unsigned)nb), 128, 0, st>>>(ra);
        pool_reduce_kernel<<<(unsigned)nb, 256, 0, st>>>(R, Cpart, Npart, dC, dN);
        TVS_CUDA(cudaGetLastError());
        // outputs of this chunk: lanes b0 .. b0+nb of the caller's lane-minor [R x B] matrices
        if (w_out) TVS_CUDA(cudaMemcpy2DAsync(w_out + b0, (size_t)B * 4, dw, (size_t)nb * 4, (size_t)nb * 4, (size_t)R, cudaMemcpyDefault, st));
        if (observed_out) TVS_CUDA(cudaMemcpy2DAsync(observed_out + b0, (size_t)B, dobs, (size_t)nb, (size_t)nb, (size_t)R, cudaMemcpyDefault, st));
        if (C_out) TVS_CUDA(cudaMemcpyAsync(C_out + b0, dC, (size_t)nb * 8, cudaMemcpyDefault, st));
        if (N_out) TVS_CUDA(cudaMemcpyAsync(N_out + b0, dN, (size_t)nb * 8, cudaMemcpyDefault, st));
        TVS_CUDA(cudaStreamSynchronize(st));
    }
    int32_t hbad = 0;
    TVS_CUDA(cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost));
    if (hbad) { set_error("tvs_pool_lanes: a drawn record index is outside the pool"); return TVS_EINVAL; }
    return TVS_OK;
}

}  // extern "C"
