// Developer microbenchmark (not part of libtvs.so): forward sparse product p = A x for 32-lane warps, V = 64, with the
// dense operand x[v][lane] (a) in shared memory, read per non-zero (what pass3_kernel does) and (b) in REGISTERS,
// selected by a warp-uniform switch on the column (indirect branch per non-zero, no shared-memory operand).
// Question for round 2: is the shared-memory operand bandwidth (the bound of the likelihood pass) avoidable?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o regx_probe tools/micro/regx_probe.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int V = 64;

__global__ void __launch_bounds__(256) smem_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                   const double* __restrict__ val, const double* __restrict__ x,
                                                   double* __restrict__ p, int R, int B, int rows_per_cta) {
    __shared__ double xs[V * 32];
    const int lane_tile = blockIdx.x, t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lane = lane_tile * 32 + t;
    for (int i = threadIdx.x; i < V * 32; i += 256) xs[i] = x[(size_t)(i >> 5) * B + lane_tile * 32 + (i & 31)];
    __syncthreads();
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
    for (int r = r0 + w; r < r1; r += 8) {
        double acc = 0.0;
        const int k1 = row_ptr[r + 1];
        for (int k = row_ptr[r]; k < k1; ++k) acc = fma(val[k], xs[col[k] * 32 + t], acc);
        p[(size_t)r * B + lane] = acc;
    }
}

#define CASE(c) case c: acc = fma(v, x##c, acc); break;
#define DECL(c) const double x##c = x[(size_t)c * B + lane];

__global__ void __launch_bounds__(256) reg_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                  const double* __restrict__ val, const double* __restrict__ x,
                                                  double* __restrict__ p, int R, int B, int rows_per_cta) {
    const int lane_tile = blockIdx.x, t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int lane = lane_tile * 32 + t;
    DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6) DECL(7) DECL(8) DECL(9) DECL(10) DECL(11) DECL(12) DECL(13) DECL(14) DECL(15)
    DECL(16) DECL(17) DECL(18) DECL(19) DECL(20) DECL(21) DECL(22) DECL(23) DECL(24) DECL(25) DECL(26) DECL(27) DECL(28) DECL(29) DECL(30) DECL(31)
    DECL(32) DECL(33) DECL(34) DECL(35) DECL(36) DECL(37) DECL(38) DECL(39) DECL(40) DECL(41) DECL(42) DECL(43) DECL(44) DECL(45) DECL(46) DECL(47)
    DECL(48) DECL(49) DECL(50) DECL(51) DECL(52) DECL(53) DECL(54) DECL(55) DECL(56) DECL(57) DECL(58) DECL(59) DECL(60) DECL(61) DECL(62) DECL(63)
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
    for (int r = r0 + w; r < r1; r += 8) {
        double acc = 0.0;
        const int k1 = row_ptr[r + 1];
        for (int k = row_ptr[r]; k < k1; ++k) {
            const double v = val[k];
            switch (col[k]) {
                CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15)
                CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22) CASE(23) CASE(24) CASE(25) CASE(26) CASE(27) CASE(28) CASE(29) CASE(30) CASE(31)
                CASE(32) CASE(33) CASE(34) CASE(35) CASE(36) CASE(37) CASE(38) CASE(39) CASE(40) CASE(41) CASE(42) CASE(43) CASE(44) CASE(45) CASE(46) CASE(47)
                CASE(48) CASE(49) CASE(50) CASE(51) CASE(52) CASE(53) CASE(54) CASE(55) CASE(56) CASE(57) CASE(58) CASE(59) CASE(60) CASE(61) CASE(62) CASE(63)
            }
        }
        p[(size_t)r * B + lane] = acc;
    }
}

int main() {
    const int R = 18032, B = 1024;
    std::vector<int> row_ptr(R + 1, 0), col;
    std::vector<double> val;
    srand(1);
    for (int r = 0; r < R; ++r) {
        int n = 0;
        for (int v = 0; v < V; ++v)
            if (rand() % 10 == 0 || (v == V - 1 && n == 0)) { col.push_back(v); val.push_back(1 + rand() % 3); ++n; }
        row_ptr[r + 1] = (int)col.size();
    }
    const int nnz = (int)col.size();
    std::vector<double> x((size_t)V * B);
    for (auto& e : x) e = (rand() % 1000 + 1) * 1e-3;
    int *d_rp, *d_col; double *d_val, *d_x, *d_p1, *d_p2;
    cudaMalloc(&d_rp, (R + 1) * 4); cudaMalloc(&d_col, nnz * 4); cudaMalloc(&d_val, nnz * 8);
    cudaMalloc(&d_x, x.size() * 8); cudaMalloc(&d_p1, (size_t)R * B * 8); cudaMalloc(&d_p2, (size_t)R * B * 8);
    cudaMemcpy(d_rp, row_ptr.data(), (R + 1) * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_col, col.data(), nnz * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_val, val.data(), nnz * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_x, x.data(), x.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int splits : {19, 37, 74}) {
        const int rows_per_cta = (R + splits - 1) / splits;
        dim3 grid(B / 32, splits);
        float ms[2];
        for (int which = 0; which < 2; ++which) {
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                for (int i = 0; i < 10; ++i) {
                    if (which == 0) smem_kernel<<<grid, 256>>>(d_rp, d_col, d_val, d_x, d_p1, R, B, rows_per_cta);
                    else reg_kernel<<<grid, 256>>>(d_rp, d_col, d_val, d_x, d_p2, R, B, rows_per_cta);
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[which], e0, e1);
            }
        }
        std::vector<double> p1((size_t)R * B), p2((size_t)R * B);
        cudaMemcpy(p1.data(), d_p1, p1.size() * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(p2.data(), d_p2, p2.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < p1.size(); ++i) bad += p1[i] != p2[i];
        printf("splits %d (CTAs %d): smem %.1f us, registers+switch %.1f us per launch, nnz %d, lanes %d, mismatches %zu, err %s\n",
               splits, (B / 32) * splits, ms[0] * 100, ms[1] * 100, nnz, B, bad, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
