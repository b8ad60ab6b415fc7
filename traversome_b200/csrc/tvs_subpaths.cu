// libtvs.so -- read-path x variant occurrence counting (include/tvs_subpaths.h).
//
// Replaces VariantSubPathsGenerator.gen_subpaths (traversome/utils.py:416-501): one thread per (variant, start
// position) grows the window contig by contig, keeps a running 64-bit hash of the tokens, and looks every prefix up
// in an open-addressing table of the read paths' key sequences (exact token comparison on a tag match, so hash
// collisions cannot change a count).  Hits go to a dense variant-major [V x R] pair of arrays with integer atomics
// (add for the count, min for the rank of the first hit): order-independent, hence deterministic.  Two coalesced
// sweeps then turn the dense arrays into the CSR that tvs_create takes.
#include "tvs_internal.h"
#include "../../include/tvs_subpaths.h"

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
    int32_t* cnt;                      // [V x R]
    uint32_t* first;                   // [V x R]
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
__global__ void sp_window_kernel(SubpathArgs a) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= a.n_starts) return;
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
                        const int64_t cell = v * a.n_rows + row;
                        atomicAdd(&a.cnt[cell], 1);
                        atomicMin(&a.first[cell], (uint32_t)(start * a.rank_stride + (a.rank_stride - 1 - len)));
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

// dense [V x R] -> per-row non-zero counts; thread per read path, coalesced over r
__global__ void sp_row_nnz_kernel(const int32_t* __restrict__ cnt, int64_t V, int64_t R, int64_t* __restrict__ row_nnz) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int64_t n = 0;
    for (int64_t v = 0; v < V; ++v) n += cnt[v * R + r] != 0;
    row_nnz[r] = n;
}

// exclusive scan of row_nnz into row_ptr[R+1] by ONE CTA of 1024 threads, 8 consecutive items per thread and round
// (R is at most a few 10^5..10^6: a few dozen rounds)
constexpr int kScanItems = 8;
__global__ void sp_scan_kernel(const int64_t* __restrict__ row_nnz, int64_t R, int64_t* __restrict__ row_ptr) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) s_carry = 0;
    __syncthreads();
    const int64_t tile = (int64_t)blockDim.x * kScanItems;
    for (int64_t r0 = 0; r0 < R; r0 += tile) {
        const int64_t base = r0 + (int64_t)t * kScanItems;
        int64_t x[kScanItems];
        int64_t sum = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            x[i] = base + i < R ? row_nnz[base + i] : 0;
            sum += x[i];
        }
        int64_t incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int64_t w = s_warp[lane];
            for (int o = 1; o < 32; o <<= 1) {
                int64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;                       // inclusive over warps
        }
        __syncthreads();
        int64_t run = s_carry + (wid ? s_warp[wid - 1] : 0) + incl - sum;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            if (base + i < R) row_ptr[base + i] = run;
            run += x[i];
        }
        __syncthreads();
        if (t == blockDim.x - 1) s_carry = run;
        __syncthreads();
    }
    if (t == 0) row_ptr[R] = s_carry;
}

// dense -> CSR entries; thread per read path, loads batched 8 variants at a time so that the sweep is bandwidth-bound
__global__ void sp_fill_kernel(const int32_t* __restrict__ cnt, const uint32_t* __restrict__ first, int64_t V, int64_t R,
                               const int64_t* __restrict__ row_ptr, int32_t* __restrict__ col, int32_t* __restrict__ val,
                               int64_t* __restrict__ rank) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int64_t k = row_ptr[r];
    if (row_ptr[r + 1] == k) return;                    // read path no variant contains
    for (int64_t v0 = 0; v0 < V; v0 += 8) {
        int32_t c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = v0 + i < V ? cnt[(v0 + i) * R + r] : 0;
        uint32_t f[8];                                    // predicated loads, issued together (not one miss per hit)
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = c[i] != 0 ? first[(v0 + i) * R + r] : 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (c[i] != 0) {
                col[k] = (int32_t)(v0 + i);
                val[k] = c[i];
                rank[k] = (int64_t)f[i];
                ++k;
            }
        }
    }
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
    int32_t* cnt = nullptr; uint32_t* first = nullptr;
    unsigned long long* n_windows = nullptr; int* max_window = nullptr; int* overflow = nullptr;
    int64_t* row_nnz = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Scratch() {
        cudaFree(var_ptr); cudaFree(var_tok); cudaFree(var_step); cudaFree(var_circ);
        cudaFree(key_ptr); cudaFree(key_tok); cudaFree(key_row_circ); cudaFree(key_row_lin);
        cudaFree(tags); cudaFree(slots); cudaFree(cnt); cudaFree(first);
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
    const size_t cells = (size_t)n_variants * (size_t)n_rows;
    TVS_CUDA(cudaMalloc((void**)&s.tags, cap * sizeof(uint64_t)));
    TVS_CUDA(cudaMalloc((void**)&s.slots, cap * sizeof(int32_t)));
    TVS_CUDA(cudaMalloc((void**)&s.cnt, (cells ? cells : 1) * sizeof(int32_t)));
    TVS_CUDA(cudaMalloc((void**)&s.first, (cells ? cells : 1) * sizeof(uint32_t)));
    TVS_CUDA(cudaMalloc((void**)&s.n_windows, sizeof(unsigned long long)));
    TVS_CUDA(cudaMalloc((void**)&s.max_window, sizeof(int)));
    TVS_CUDA(cudaMalloc((void**)&s.overflow, sizeof(int)));
    TVS_CUDA(cudaMalloc((void**)&s.row_nnz, ((size_t)n_rows + 1) * sizeof(int64_t)));
    TVS_CUDA(cudaEventCreate(&s.e0));
    TVS_CUDA(cudaEventCreate(&s.e1));

    SubpathArgs a{};
    a.n_variants = n_variants; a.n_starts = n_tok; a.n_rows = n_rows; a.n_keys = n_keys;
    a.var_ptr = s.var_ptr; a.var_tok = s.var_tok; a.var_step = s.var_step; a.var_circular = s.var_circ;
    a.key_ptr = s.key_ptr; a.key_tok = s.key_tok; a.key_row_circ = s.key_row_circ; a.key_row_lin = s.key_row_lin;
    a.max_internal_len = max_internal_len;
    a.tags = s.tags; a.slots = s.slots; a.cap_mask = cap - 1;
    a.cnt = s.cnt; a.first = s.first;
    a.n_windows = s.n_windows; a.max_window = s.max_window; a.overflow = s.overflow;

    const int T = 256;
    TVS_CUDA(cudaEventRecord(s.e0));
    TVS_CUDA(cudaMemsetAsync(s.tags, 0, cap * sizeof(uint64_t)));
    TVS_CUDA(cudaMemsetAsync(s.cnt, 0, (cells ? cells : 1) * sizeof(int32_t)));
    TVS_CUDA(cudaMemsetAsync(s.first, 0xff, (cells ? cells : 1) * sizeof(uint32_t)));
    TVS_CUDA(cudaMemsetAsync(s.n_windows, 0, sizeof(unsigned long long)));
    TVS_CUDA(cudaMemsetAsync(s.max_window, 0, sizeof(int)));
    TVS_CUDA(cudaMemsetAsync(s.overflow, 0, sizeof(int)));
    if (n_keys) sp_insert_kernel<<<(unsigned)((n_keys + T - 1) / T), T>>>(a);
    if (n_tok) sp_stats_kernel<<<(unsigned)((n_tok + T - 1) / T), T>>>(a);
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
    if (n_tok && n_keys && cells) sp_window_kernel<<<(unsigned)((n_tok + T - 1) / T), T>>>(a);
    TVS_CUDA(cudaGetLastError());

    // dense -> CSR by read path
    TVS_CUDA(cudaMalloc((void**)&res->row_ptr, ((size_t)n_rows + 1) * sizeof(int64_t)));
    if (n_rows) sp_row_nnz_kernel<<<(unsigned)((n_rows + T - 1) / T), T>>>(s.cnt, n_variants, n_rows, s.row_nnz);
    sp_scan_kernel<<<1, 1024>>>(s.row_nnz, n_rows, res->row_ptr);
    TVS_CUDA(cudaGetLastError());
    int64_t nnz = 0;
    TVS_CUDA(cudaMemcpy(&nnz, res->row_ptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost));
    res->nnz = nnz;
    TVS_CUDA(cudaMalloc((void**)&res->col, (size_t)(nnz ? nnz : 1) * sizeof(int32_t)));
    TVS_CUDA(cudaMalloc((void**)&res->val, (size_t)(nnz ? nnz : 1) * sizeof(int32_t)));
    TVS_CUDA(cudaMalloc((void**)&res->rank, (size_t)(nnz ? nnz : 1) * sizeof(int64_t)));
    if (n_rows && nnz)
        sp_fill_kernel<<<(unsigned)((n_rows + T - 1) / T), T>>>(s.cnt, s.first, n_variants, n_rows, res->row_ptr, res->col,
                                                               res->val, res->rank);
    TVS_CUDA(cudaGetLastError());
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
    size_t free_b = 0, total_b = 0;
    TVS_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if ((double)n_variants * (double)n_rows * 8.0 > 0.8 * (double)free_b) {
        set_error("tvs_subpaths_count: %lld variants x %lld read paths needs %.1f GB of scratch, %.1f GB free",
                  (long long)n_variants, (long long)n_rows, (double)n_variants * n_rows * 8e-9, free_b * 1e-9);
        return TVS_ENOMEM;
    }
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
