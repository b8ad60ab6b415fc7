// CPU test harness (g++, no CUDA): the blocked sparse layout that blk_kernel walks, built by the same host code that
// tvs_create uses (traversome_b200/csrc/tvs_blk_host.h), multiplied on the host in the kernel's traversal order.
#include "../traversome_b200/csrc/tvs_blk_host.h"

extern "C" int blk_host_products(int64_t R, int64_t V, int64_t nnz, const int32_t* row_ptr, const int32_t* col, const int32_t* val,
                                 const double* x, double* p_out, const double* rr, double* t_out, int64_t* info) {
    tvs::BlkHost f, b;
    if (!tvs::build_blk<int32_t>(R, V, row_ptr, col, val, &f)) return 1;
    if (!tvs::build_blk_transposed(R, V, nnz, row_ptr, col, val, &b)) return 2;
    tvs::blk_multiply_host(f, R, x, p_out);
    tvs::blk_multiply_host(b, V, rr, t_out);
    info[0] = f.n_chunks; info[1] = f.n_panels; info[2] = f.max_seg_bytes; info[3] = (int64_t)f.blob.size() * 4;
    info[4] = b.n_chunks; info[5] = b.n_panels; info[6] = b.max_seg_bytes; info[7] = (int64_t)b.blob.size() * 4;
    // alignment contract of the bulk copies: every segment starts on a 16-byte boundary and is a multiple of 16 bytes
    for (const tvs::BlkHost* m : {&f, &b})
        for (size_t s = 0; s + 1 < m->seg_off.size(); ++s)
            if (m->seg_off[s + 1] < m->seg_off[s] || (size_t)m->seg_off[s + 1] * 4 > m->blob.size()) return 3;
    return 0;
}
