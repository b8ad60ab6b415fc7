// libtvs.so -- read-path x variant occurrence counting (include/tvs_subpaths.h).
//
// Replaces VariantSubPathsGenerator.gen_subpaths (traversome/utils.py:416-501): one thread per (variant, start
// position) grows the window contig by contig, keeps a running 64-bit hash of the tokens, and looks every prefix up
// in an open-addressing table of the read paths' key sequences (exact token comparison on a tag match, so hash
// collisions cannot change a count).  Every hit becomes one record {(read path, variant), rank of the window in the
// reference's enumeration order}: a first sweep counts the hits, a second writes them; the records are radix-sorted by
// (read path, variant) (cub::DeviceRadixSort) and each run is reduced to (occurrences, smallest rank).  The result is
// the CSR that tvs_create takes, independent of the order in which threads found the hits.  Memory is proportional to
// the number of hits, not to variants x read paths (the first version kept a dense [V x R] scratch: 3.3 GB and 80 % of
// the device time at 2,048 x 200k, and out of reach at 10^5 x 10^5).
#include "tvs_internal.h"
#include "../../include/tvs_subpaths.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>

namespace tvs {
namespace {

constexpr uint64_t kHashSeed = 0xcbf29ce484222325ull;
constexpr uint64_t kHashMul = 0x100000001b3ull;
constexpr int kMaxWindow = 1 << 16;       // contigs per window; beyond this the graph has zero-length cycles

__device__ __forceinline__ uint64_t hash_step(uint64_t h, int32_t tok) {
    return (h ^ (uint64_t)(uint32_t)(tok + 1)) * kHashMul;
}
// tag stored in the table: avalanche of the running hash, never 0 (0 = empty slot)
__device__ __forceinline__ uint64_t hash_tag(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h ? h : 1ull;
}

struct SubpathArgs {
    int64_t n_variants, n_starts, n_rows, n_keys;
    const int64_t* var_ptr;
    const int32_t* var_tok;
    const int32_t* var_step;
    const uint8_t* var_circular;
    const int64_t* key_ptr;
    const int32_t* key_tok;
    const int32_t* key_row_circ;
    const int32_t* key_row_lin;
    int64_t max_internal_len;
    uint64_t* tags;                    // [cap]
    int32_t* slots;                    // [cap] key index
    uint64_t cap_mask;
    unsigned long long* n_hits;        // [1] number of hits (count sweep) / write cursor (emit sweep)
    uint64_t* hit_key;                 // [hits] read path << 32 | variant   (null in the count sweep)
    uint32_t* hit_rank;                // [hits] start * rank_stride + (rank_stride - 1 - length)
    unsigned long long* n_windows;     // [1]
    int* max_window;                   // [1]
    int* overflow;                     // [1]
    int rank_stride;                   // max_window + 1
};

__device__ __forceinline__ int64_t variant_of(const int64_t* __restrict__ var_ptr, int64_t n_variants, int64_t pos) {
    int64_t lo = 0, hi = n_variants;          // largest v with var_ptr[v] <= pos
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (var_ptr[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void sp_insert_kernel(SubpathArgs a) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_keys) return;
    uint64_t h = kHashSeed;
    for (int64_t i = a.key_ptr[k]; i < a.key_ptr[k + 1]; ++i) h = hash_step(h, a.key_tok[i]);
    uint64_t tag = hash_tag(h);
    uint64_t s = tag & a.cap_mask;
    for (;;) {
        unsigned long long old = atomicCAS((unsigned long long*)&a.tags[s], 0ull, (unsigned long long)tag);
        if (old == 0ull) { a.slots[s] = (int32_t)k; return; }
        s = (s + 1) & a.cap_mask;             // occupied (same tag or not): keys are distinct, keep probing
    }
}

// One thread per (variant, start): grows the window (utils.py:441-457 circular / :471-486 linear) and looks every
// prefix up (utils.py:459-470 / :488-498).
template <bool EMIT>
__global__ void sp_window_kernel(SubpathArgs a) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= a.n_starts) return;
    unsigned long long my_hits = 0;
    const int64_t v = variant_of(a.var_ptr, a.n_variants, pos);
    const int64_t base = a.var_ptr[v];
    const int64_t n = a.var_ptr[v + 1] - base;
    const int64_t start = pos - base;
    const bool circ = a.var_circular[v] != 0;
    const int32_t* __restrict__ key_row = circ ? a.key_row_circ : a.key_row_lin;

    int len = 1;
    int64_t j = start;
    int64_t internal = 0;
    uint64_t h = hash_step(kHashSeed, a.var_tok[base + start]);
    for (;;) {
        // look the current prefix (tokens start .. start+len-1, cyclic) up
        const uint64_t tag = hash_tag(h);
        uint64_t s = tag & a.cap_mask;
        for (;;) {
            const uint64_t t = a.tags[s];
            if (t == 0ull) break;
            if (t == tag) {
                const int32_t k = a.slots[s];
                const int64_t k0 = a.key_ptr[k];
                const int32_t row = key_row[k];
                if (row >= 0 && a.key_ptr[k + 1] - k0 == len) {
                    bool same = true;
                    int64_t q = start;
                    for (int i = 0; i < len; ++i) {
                        if (a.key_tok[k0 + i] != a.var_tok[base + q]) { same = false; break; }
                        if (++q == n) q = 0;
                    }
                    if (same) {
                        if (EMIT) {
                            const unsigned long long at = atomicAdd(a.n_hits, 1ull);
                            a.hit_key[at] = ((uint64_t)(uint32_t)row << 32) | (uint64_t)(uint32_t)v;
                            a.hit_rank[at] = (uint32_t)(start * a.rank_stride + (a.rank_stride - 1 - len));
                        } else ++my_hits;
                        break;
                    }
                }
            }
            s = (s + 1) & a.cap_mask;
        }
        // extend by one contig while the accumulated length allows it
        if (internal > a.max_internal_len) break;
        ++j;
        if (j == n) {
            if (!circ) break;
            j = 0;
        }
        if (len >= kMaxWindow) break;
        internal += a.var_step[base + j];
        ++len;
        h = hash_step(h, a.var_tok[base + j]);
    }
    if (!EMIT && my_hits) atomicAdd(a.n_hits, my_hits);
}

// Statistics pre-pass: longest window over all starts (sizes the 32-bit rank) and the number of windows.
__global__ void sp_stats_kernel(SubpathArgs a) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    bool over = false;
    if (pos < a.n_starts) {
        const int64_t v = variant_of(a.var_ptr, a.n_variants, pos);
        const int64_t base = a.var_ptr[v];
        const int64_t n = a.var_ptr[v + 1] - base;
        const bool circ = a.var_circular[v] != 0;
        int64_t j = pos - base;
        int64_t internal = 0;
        len = 1;
        for (;;) {
            if (internal > a.max_internal_len) break;
            ++j;
            if (j == n) {
                if (!circ) break;
                j = 0;
            }
            if (len >= kMaxWindow) { over = true; break; }
            internal += a.var_step[base + j];
            ++len;
        }
    }
    __shared__ int s_max[32];
    __shared__ unsigned long long s_sum[32];
    __shared__ int s_over;
    if (threadIdx.x == 0) s_over = 0;
    __syncthreads();
    if (over) s_over = 1;
    int m = len;
    unsigned long long w = (unsigned long long)len;
    for (int o = 16; o > 0; o >>= 1) {
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        w += __shfl_xor_sync(0xffffffffu, w, o);
    }
    if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5] = m; s_sum[threadIdx.x >> 5] = w; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int i = 1; i < nw; ++i) { m = max(m, s_max[i]); w += s_sum[i]; }
        atomicMax(a.max_window, m);
        atomicAdd(a.n_windows, w);
        if (s_over) atomicExch(a.overflow, 1);
    }
}

// sorted hit records -> run heads: head[i] = 1 where a new (read path, variant) pair starts
__global__ void sp_heads_kernel(const uint64_t* __restrict__ key, int64_t n, int32_t* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// one thread per run head: occurrences = run length, first = smallest rank of the run; also the per-row entry counts
__global__ void sp_runs_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ rank, const int32_t* __restrict__ head,
                               const int32_t* __restrict__ pos, int64_t n, int32_t* __restrict__ col, int32_t* __restrict__ val,
                               int64_t* __restrict__ first, unsigned long long* __restrict__ row_nnz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || head[i] == 0) return;
    const uint64_t k = key[i];
    uint32_t best = rank[i];
    int32_t c = 1;
    for (int64_t j = i + 1; j < n && key[j] == k; ++j) { best = min(best, rank[j]); ++c; }
    const int32_t o = pos[i];
    col[o] = (int32_t)(uint32_t)(k & 0xffffffffull);
    val[o] = c;
    first[o] = (int64_t)best;
    atomicAdd(&row_nnz[k >> 32], 1ull);
}

template <class T>
int upload(T** dst, const T* src, size_t n) {
    *dst = nullptr;
    TVS_CUDA(cudaMalloc((void**)dst, (n ? n : 1) * sizeof(T)));
    if (n) TVS_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyDefault));
    return TVS_OK;
}

}  // namespace
}  // namespace tvs

struct tvs_subpath_counts {
    int device = 0;
    int64_t n_rows = 0, nnz = 0;
    int64_t* row_ptr = nullptr;   // device
    int32_t* col = nullptr;
    int32_t* val = nullptr;
    int64_t* rank = nullptr;
};

namespace tvs {
namespace {

struct Scratch {
    int64_t* var_ptr = nullptr; int32_t* var_tok = nullptr; int32_t* var_step = nullptr; uint8_t* var_circ = nullptr;
    int64_t* key_ptr = nullptr; int32_t* key_tok = nullptr; int32_t* key_row_circ = nullptr; int32_t* key_row_lin = nullptr;
    uint64_t* tags = nullptr; int32_t* slots = nullptr;
    unsigned long long* n_hits = nullptr;
    uint64_t* key_a = nullptr; uint64_t* key_b = nullptr; uint32_t* rank_a = nullptr; uint32_t* rank_b = nullptr;
    int32_t* head = nullptr; int32_t* pos = nullptr; void* cub_tmp = nullptr;
    unsigned long long* n_windows = nullptr; int* max_window = nullptr; int* overflow = nullptr;
    unsigned long long* row_nnz = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Scratch() {
        cudaFree(var_ptr); cudaFree(var_tok); cudaFree(var_step); cudaFree(var_circ);
        cudaFree(key_ptr); cudaFree(key_tok); cudaFree(key_row_circ); cudaFree(key_row_lin);
        cudaFree(tags); cudaFree(slots); cudaFree(n_hits);
        cudaFree(key_a); cudaFree(key_b); cudaFree(rank_a); cudaFree(rank_b); cudaFree(head); cudaFree(pos); cudaFree(cub_tmp);
        cudaFree(n_windows); cudaFree(max_window); cudaFree(overflow); cudaFree(row_nnz);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

int count_impl(tvs_subpath_counts* res, int64_t n_variants, const int64_t* var_ptr, const int32_t* var_tok,
               const int32_t* var_step, const uint8_t* var_circular, int64_t n_rows, int64_t n_keys,
               const int64_t* key_ptr, const int32_t* key_tok, const int32_t* key_row_circ, const int32_t* key_row_lin,
               int64_t max_internal_len, int64_t n_tok, int64_t n_key_tok, int64_t* n_windows_out, double* kernel_ms_out) {
    Scratch s;
    int rc;
    if ((rc = upload(&s.var_ptr, var_ptr, (size_t)n_variants + 1))) return rc;
    if ((rc = upload(&s.var_tok, var_tok, (size_t)n_tok))) return rc;
    if ((rc = upload(&s.var_step, var_step, (size_t)n_tok))) return rc;
    if ((rc = upload(&s.var_circ, var_circular, (size_t)n_variants))) return rc;
    if ((rc = upload(&s.key_ptr, key_ptr, (size_t)n_keys + 1))) return rc;
    if ((rc = upload(&s.key_tok, key_tok, (size_t)n_key_tok))) return rc;
    if ((rc = upload(&s.key_row_circ, key_row_circ, (size_t)n_keys))) return rc;
    if ((rc = upload(&s.key_row_lin, key_row_lin, (size_t)n_keys))) return rc;

    uint64_t cap = 64;
    while (cap < (uint64_t)n_keys * 2 + 2) cap <<= 1;
    TVS_CUDA(cudaMalloc((void**)&s.tags, cap * sizeof(uint64_t)));
    TVS_CUDA(cudaMalloc((void**)&s.slots, cap * sizeof(int32_t)));
    TVS_CUDA(cudaMalloc((void**)&s.n_hits, sizeof(unsigned long long)));
    TVS_CUDA(cudaMalloc((void**)&s.n_windows, sizeof(unsigned long long)));
    TVS_CUDA(cudaMalloc((void**)&s.max_window, sizeof(int)));
    TVS_CUDA(cudaMalloc((void**)&s.overflow, sizeof(int)));
    TVS_CUDA(cudaMalloc((void**)&s.row_nnz, ((size_t)n_rows + 1) * sizeof(unsigned long long)));
    TVS_CUDA(cudaEventCreate(&s.e0));
    TVS_CUDA(cudaEventCreate(&s.e1));

    SubpathArgs a{};
    a.n_variants = n_variants; a.n_starts = n_tok; a.n_rows = n_rows; a.n_keys = n_keys;
    a.var_ptr = s.var_ptr; a.var_tok = s.var_tok; a.var_step = s.var_step; a.var_circular = s.var_circ;
    a.key_ptr = s.key_ptr; a.key_tok = s.key_tok; a.key_row_circ = s.key_row_circ; a.key_row_lin = s.key_row_lin;
    a.max_internal_len = max_internal_len;
    a.tags = s.tags; a.slots = s.slots; a.cap_mask = cap - 1;
    a.n_hits = s.n_hits; a.hit_key = nullptr; a.hit_rank = nullptr;
    a.n_windows = s.n_windows; a.max_window = s.max_window; a.overflow = s.overflow;

    const int T = 256;
    const unsigned grid_starts = (unsigned)((n_tok + T - 1) / T);
    TVS_CUDA(cudaEventRecord(s.e0));
    TVS_CUDA(cudaMemsetAsync(s.tags, 0, cap * sizeof(uint64_t)));
    TVS_CUDA(cudaMemsetAsync(s.n_hits, 0, sizeof(unsigned long long)));
    TVS_CUDA(cudaMemsetAsync(s.n_windows, 0, sizeof(unsigned long long)));
    TVS_CUDA(cudaMemsetAsync(s.max_window, 0, sizeof(int)));
    TVS_CUDA(cudaMemsetAsync(s.overflow, 0, sizeof(int)));
    TVS_CUDA(cudaMemsetAsync(s.row_nnz, 0, ((size_t)n_rows + 1) * sizeof(unsigned long long)));
    if (n_keys) sp_insert_kernel<<<(unsigned)((n_keys + T - 1) / T), T>>>(a);
    if (n_tok) sp_stats_kernel<<<grid_starts, T>>>(a);
    TVS_CUDA(cudaGetLastError());
    int h_max = 0, h_over = 0;
    unsigned long long h_windows = 0;
    TVS_CUDA(cudaMemcpy(&h_max, s.max_window, sizeof(int), cudaMemcpyDeviceToHost));
    TVS_CUDA(cudaMemcpy(&h_over, s.overflow, sizeof(int), cudaMemcpyDeviceToHost));
    TVS_CUDA(cudaMemcpy(&h_windows, s.n_windows, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (h_over) {
        set_error("tvs_subpaths_count: a window grew past %d contigs (cycle of zero-length steps?)", kMaxWindow);
        return TVS_ELIMIT;
    }
    int64_t longest_variant = 0;
    for (int64_t v = 0; v < n_variants; ++v) longest_variant = std::max<int64_t>(longest_variant, var_ptr[v + 1] - var_ptr[v]);
    if ((double)longest_variant * (double)(h_max + 1) >= 4294967295.0) {
        set_error("tvs_subpaths_count: variant of %lld contigs x window of %d contigs overflows the 32-bit rank",
                  (long long)longest_variant, h_max);
        return TVS_ELIMIT;
    }
    a.rank_stride = h_max + 1;
    // ---- hit records.  Every window yields at most one hit, so with few windows (< 64 M: < 2 GB of records and sort
    //      buffers) the list is sized for all of them and written in ONE sweep; otherwise a first sweep counts the hits.
    unsigned long long n_hits = 0;
    const bool any = n_tok && n_keys && n_rows;
    const bool one_sweep = any && h_windows <= (64ull << 20);
    if (any && !one_sweep) {
        sp_window_kernel<false><<<grid_starts, T>>>(a);
        TVS_CUDA(cudaGetLastError());
        TVS_CUDA(cudaMemcpy(&n_hits, s.n_hits, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    const unsigned long long cap_hits = one_sweep ? h_windows : n_hits;
    if (cap_hits > 2000000000ull) { set_error("tvs_subpaths_count: %llu hits exceed the 32-bit record index", cap_hits); return TVS_ELIMIT; }
    int64_t H = (int64_t)cap_hits;
    int64_t nnz = 0;
    TVS_CUDA(cudaMalloc((void**)&res->row_ptr, ((size_t)n_rows + 1) * sizeof(int64_t)));
    if (H > 0) {
        TVS_CUDA(cudaMalloc((void**)&s.key_a, (size_t)H * sizeof(uint64_t)));
        TVS_CUDA(cudaMalloc((void**)&s.key_b, (size_t)H * sizeof(uint64_t)));
        TVS_CUDA(cudaMalloc((void**)&s.rank_a, (size_t)H * sizeof(uint32_t)));
        TVS_CUDA(cudaMalloc((void**)&s.rank_b, (size_t)H * sizeof(uint32_t)));
        TVS_CUDA(cudaMalloc((void**)&s.head, (size_t)H * sizeof(int32_t)));
        TVS_CUDA(cudaMalloc((void**)&s.pos, (size_t)H * sizeof(int32_t)));
        TVS_CUDA(cudaMemsetAsync(s.n_hits, 0, sizeof(unsigned long long)));
        a.hit_key = s.key_a; a.hit_rank = s.rank_a;
        sp_window_kernel<true><<<grid_starts, T>>>(a);
        TVS_CUDA(cudaGetLastError());
        if (one_sweep) {
            TVS_CUDA(cudaMemcpy(&n_hits, s.n_hits, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            H = (int64_t)n_hits;
        }
    }
    if (H > 0) {
        // ---- sort by (read path, variant); key bits actually in use: 32 for the variant + those of n_rows
        int row_bits = 1;
        while (((int64_t)1 << row_bits) < n_rows) ++row_bits;
        size_t tmp_sort = 0, tmp_scan = 0;
        cub::DoubleBuffer<uint64_t> dk(s.key_a, s.key_b);
        cub::DoubleBuffer<uint32_t> dv(s.rank_a, s.rank_b);
        TVS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, dk, dv, (int)H, 0, 32 + row_bits));
        TVS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, s.head, s.pos, (int)H));
        size_t tmp_rows = 0;
        TVS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_rows, s.row_nnz, (unsigned long long*)res->row_ptr, (int)n_rows + 1));
        const size_t tmp_bytes = std::max(tmp_sort, std::max(tmp_scan, tmp_rows));
        TVS_CUDA(cudaMalloc(&s.cub_tmp, tmp_bytes ? tmp_bytes : 1));
        size_t tb = tmp_bytes;
        TVS_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_tmp, tb, dk, dv, (int)H, 0, 32 + row_bits));
        const uint64_t* keys = dk.Current();
        const uint32_t* ranks = dv.Current();
        sp_heads_kernel<<<(unsigned)((H + T - 1) / T), T>>>(keys, H, s.head);
        tb = tmp_bytes;
        TVS_CUDA(cub::DeviceScan::ExclusiveSum(s.cub_tmp, tb, s.head, s.pos, (int)H));
        int32_t last_pos = 0, last_head = 0;
        TVS_CUDA(cudaMemcpy(&last_pos, s.pos + (H - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
        TVS_CUDA(cudaMemcpy(&last_head, s.head + (H - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
        nnz = (int64_t)last_pos + last_head;
        TVS_CUDA(cudaMalloc((void**)&res->col, (size_t)nnz * sizeof(int32_t)));
        TVS_CUDA(cudaMalloc((void**)&res->val, (size_t)nnz * sizeof(int32_t)));
        TVS_CUDA(cudaMalloc((void**)&res->rank, (size_t)nnz * sizeof(int64_t)));
        sp_runs_kernel<<<(unsigned)((H + T - 1) / T), T>>>(keys, ranks, s.head, s.pos, H, res->col, res->val, res->rank, s.row_nnz);
        TVS_CUDA(cudaGetLastError());
        tb = tmp_bytes;
        TVS_CUDA(cub::DeviceScan::ExclusiveSum(s.cub_tmp, tb, s.row_nnz, (unsigned long long*)res->row_ptr, (int)n_rows + 1));
    } else {
        TVS_CUDA(cudaMemsetAsync(res->row_ptr, 0, ((size_t)n_rows + 1) * sizeof(int64_t)));
        TVS_CUDA(cudaMalloc((void**)&res->col, sizeof(int32_t)));
        TVS_CUDA(cudaMalloc((void**)&res->val, sizeof(int32_t)));
        TVS_CUDA(cudaMalloc((void**)&res->rank, sizeof(int64_t)));
    }
    res->nnz = nnz;
    TVS_CUDA(cudaEventRecord(s.e1));
    TVS_CUDA(cudaEventSynchronize(s.e1));
    float ms = 0.f;
    TVS_CUDA(cudaEventElapsedTime(&ms, s.e0, s.e1));
    if (n_windows_out) *n_windows_out = (int64_t)h_windows;
    if (kernel_ms_out) *kernel_ms_out = (double)ms;
    return TVS_OK;
}

}  // namespace
}  // namespace tvs

extern "C" {

int tvs_subpaths_count(tvs_subpaths_handle* out, int device, int64_t n_variants, const int64_t* var_ptr,
                       const int32_t* var_tok, const int32_t* var_step, const uint8_t* var_circular, int64_t n_rows,
                       int64_t n_keys, const int64_t* key_ptr, const int32_t* key_tok, const int32_t* key_row_circular,
                       const int32_t* key_row_linear, int64_t max_internal_len, int64_t* nnz_out, int64_t* n_windows_out,
                       double* kernel_ms_out) {
    using namespace tvs;
    if (!out) { set_error("tvs_subpaths_count: out is null"); return TVS_EINVAL; }
    *out = nullptr;
    if (n_variants < 0 || n_rows < 0 || n_keys < 0 || !var_ptr || !key_ptr || (!var_circular && n_variants)) {
        set_error("tvs_subpaths_count: bad sizes or null arrays");
        return TVS_EINVAL;
    }
    // the pointer tables are read on the host for validation: they must be host arrays
    if (var_ptr[0] != 0 || key_ptr[0] != 0) { set_error("tvs_subpaths_count: var_ptr[0] or key_ptr[0] != 0"); return TVS_EINVAL; }
    for (int64_t v = 0; v < n_variants; ++v)
        if (var_ptr[v + 1] <= var_ptr[v]) { set_error("tvs_subpaths_count: variant %lld is empty or var_ptr not monotone", (long long)v); return TVS_EINVAL; }
    for (int64_t k = 0; k < n_keys; ++k)
        if (key_ptr[k + 1] <= key_ptr[k]) { set_error("tvs_subpaths_count: key %lld is empty or key_ptr not monotone", (long long)k); return TVS_EINVAL; }
    const int64_t n_tok = var_ptr[n_variants], n_key_tok = key_ptr[n_keys];
    if ((n_tok && (!var_tok || !var_step)) || (n_keys && (!key_tok || !key_row_circular || !key_row_linear))) {
        set_error("tvs_subpaths_count: null token arrays");
        return TVS_EINVAL;
    }
    if (n_keys > 1000000000LL || n_rows > 2000000000LL || n_variants > 2000000000LL) {
        set_error("tvs_subpaths_count: sizes beyond 32-bit indices");
        return TVS_ELIMIT;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("tvs_subpaths_count: no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return TVS_ECUDA;
    }
    if (device < 0 || device >= ndev) { set_error("tvs_subpaths_count: device %d out of range", device); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(device));
    tvs_subpath_counts* res = new (std::nothrow) tvs_subpath_counts();
    if (!res) { set_error("out of host memory"); return TVS_ENOMEM; }
    res->device = device;
    res->n_rows = n_rows;
    int rc = count_impl(res, n_variants, var_ptr, var_tok, var_step, var_circular, n_rows, n_keys, key_ptr, key_tok,
                        key_row_circular, key_row_linear, max_internal_len, n_tok, n_key_tok, n_windows_out, kernel_ms_out);
    if (rc != TVS_OK) { tvs_subpaths_destroy(res); return rc; }
    if (nnz_out) *nnz_out = res->nnz;
    *out = res;
    return TVS_OK;
}

int tvs_subpaths_fetch(tvs_subpaths_handle h, int64_t* row_ptr, int32_t* col, int32_t* val, int64_t* first) {
    using namespace tvs;
    if (!h) { set_error("tvs_subpaths_fetch: null handle"); return TVS_EINVAL; }
    TVS_CUDA(cudaSetDevice(h->device));
    if (row_ptr) TVS_CUDA(cudaMemcpy(row_ptr, h->row_ptr, ((size_t)h->n_rows + 1) * sizeof(int64_t), cudaMemcpyDefault));
    if (h->nnz) {
        if (col) TVS_CUDA(cudaMemcpy(col, h->col, (size_t)h->nnz * sizeof(int32_t), cudaMemcpyDefault));
        if (val) TVS_CUDA(cudaMemcpy(val, h->val, (size_t)h->nnz * sizeof(int32_t), cudaMemcpyDefault));
        if (first) TVS_CUDA(cudaMemcpy(first, h->rank, (size_t)h->nnz * sizeof(int64_t), cudaMemcpyDefault));
    }
    return TVS_OK;
}

int tvs_subpaths_destroy(tvs_subpaths_handle h) {
    if (!h) return TVS_OK;
    cudaSetDevice(h->device);
    cudaFree(h->row_ptr); cudaFree(h->col); cudaFree(h->val); cudaFree(h->rank);
    delete h;
    return TVS_OK;
}

}  // extern "C"
