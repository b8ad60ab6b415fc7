// Host-side construction of the blocked sparse matrices that blk_kernel (tvs_blk.cuh) consumes.  Plain C++ (no CUDA):
// included by tvs_api.cu and by the CPU unit test harness tests/blk_host_harness.cpp.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace tvs {

constexpr int kBlkCH = 240;                         // major indices per chunk = 15 consumer warps x 16 accumulator rows
constexpr int kBlkCW = 15;                          // consumer warps of blk_kernel (tvs_blk.cuh: kBlkWarps)
constexpr int kBlkRW = kBlkCH / kBlkCW;             // accumulator rows per warp

// Which major index sits in accumulator row `i` of consumer warp `wi` of chunk `c`: the major indices are dealt out round
// robin, first over the chunks, then over the warps of a chunk, so that every chunk and every warp holds the same number
// (+-1).  With contiguous ranges 256 variants were a chunk of 240 and a chunk of 16: half of the backward CTAs of cfg3 kept
// one warp of 15 busy while they streamed every panel (26 % of the launch's stall samples on the ring's wait).
// A matrix of many chunks keeps contiguous ranges (dealt = false: chunk c holds c * 240 ... + 239, warp wi rows wi * 16 ...):
// there the last chunk hardly matters and neighbouring rows of the weights / rr stay neighbours in memory (cfg4: dealing
// cost 2 % per pass).  blk_dealt() is the rule.
static inline bool blk_dealt(int64_t n_major) { return (n_major + kBlkCH - 1) / kBlkCH <= 4 && n_major % kBlkCH != 0; }
static inline int64_t blk_major_of(int c, int wi, int i, int n_chunks, bool dealt) {
    return dealt ? ((int64_t)i * kBlkCW + wi) * n_chunks + c : (int64_t)c * kBlkCH + wi * kBlkRW + i;
}
// the inverse: slot (= wi * rows per warp + i) of the j-th major index of a dealt chunk
static inline int blk_slot_of(int64_t j) { return (int)(j % kBlkCW) * kBlkRW + (int)(j / kBlkCW); }
constexpr int kBlkPN = 128;                         // minor indices per panel (rows of the operand tile)
constexpr int kBlkPtrBytes = ((kBlkCH + 1) * 2 + 15) & ~15;

// ---- blocked matrices for blk_kernel (tvs_blk.cuh).  `ptr/idx/val`: a major-sorted sparse matrix (CSR of A for the
// forward product, CSR of A^T for the backward one).  Segment (chunk, panel) = u16 entry pointers of the chunk's 256
// major indices + packed entries (minor_local * 512 | top 16 bits of the fp64 count).  A count that needs more than 5
// significant bits is split into several entries of the same position (each exactly representable in 16 bits).  The entry
// list of every major index is padded to an even length with zero entries (count +0.0).
struct BlkHost {
    std::vector<uint32_t> seg_off;     // 16-byte units
    std::vector<uint32_t> blob;        // 32-bit words
    int n_chunks = 0, n_panels = 0, max_seg_bytes = 0;
    bool dealt = false;                // major indices dealt round robin (blk_major_of)
};

static inline int blk_terms(uint32_t c, uint32_t* out) {
    int n = 0;
    while (c) {
        const int msb = 31 - __builtin_clz(c);
        const int sh = msb > 4 ? msb - 4 : 0;
        const uint32_t tval = (c >> sh) << sh;
        out[n++] = tval;
        c -= tval;
    }
    return n;
}
static inline uint32_t blk_hi16(uint32_t tval) {
    const double d = (double)tval;
    uint64_t bits;
    memcpy(&bits, &d, 8);
    return (uint32_t)(bits >> 48);
}

template <class PtrT>
inline bool build_blk(int64_t n_major, int64_t n_minor, const PtrT* ptr, const int32_t* idx, const int32_t* val, BlkHost* out) {
    const int CH = kBlkCH, PN = kBlkPN;
    const int n_chunks = (int)((n_major + CH - 1) / CH), n_panels = (int)((n_minor + PN - 1) / PN);
    if ((int64_t)n_chunks * n_panels > ((int64_t)1 << 28)) return false;
    out->n_chunks = n_chunks; out->n_panels = n_panels; out->max_seg_bytes = 0;
    const bool dealt = out->dealt = blk_dealt(n_major);
    out->seg_off.assign((size_t)n_chunks * n_panels + 1, 0);
    out->blob.clear();
    std::vector<int32_t> cnt((size_t)n_panels * (CH + 1));
    std::vector<int64_t> base((size_t)n_panels);
    const int ptr_words = kBlkPtrBytes / 4;
    uint32_t terms[8];
    for (int ch = 0; ch < n_chunks; ++ch) {
        // the chunk's major indices: dealt -- ch, ch + n_chunks, ch + 2 n_chunks, ... (j-th of them in slot blk_slot_of(j));
        // else the contiguous range ch * CH ... (j-th in slot j)
        const int64_t n_here = dealt ? (n_major > ch ? (n_major - ch + n_chunks - 1) / n_chunks : 0)
                                     : std::max<int64_t>(0, std::min<int64_t>(n_major, (int64_t)(ch + 1) * CH) - (int64_t)ch * CH);
        std::fill(cnt.begin(), cnt.end(), 0);
        for (int64_t j = 0; j < n_here; ++j) {
            const int64_t m = dealt ? j * n_chunks + ch : (int64_t)ch * CH + j;
            const int slot = dealt ? blk_slot_of(j) : (int)j;
            for (int64_t k = (int64_t)ptr[m]; k < (int64_t)ptr[m + 1]; ++k)
                cnt[(size_t)(idx[k] / PN) * (CH + 1) + slot + 1] += blk_terms((uint32_t)val[k], terms);
        }
        // per panel: prefix sums over the major index (every list padded to an even length: the kernel reads entries in
        // 8-byte pairs; a padding entry is 0 = offset 0 with count +0.0), segment sizes
        for (int pn = 0; pn < n_panels; ++pn) {
            int32_t* cp = &cnt[(size_t)pn * (CH + 1)];
            for (int i = 1; i <= CH; ++i) cp[i] = cp[i - 1] + ((cp[i] + 1) & ~1);
            const int64_t n_ent = cp[CH];
            const size_t seg = (size_t)ch * n_panels + pn;
            out->seg_off[seg] = (uint32_t)(out->blob.size() / 4);
            if (n_ent == 0) { base[pn] = -1; continue; }
            if (n_ent > 65535) return false;                                 // u16 pointers
            const size_t words = (size_t)ptr_words + (size_t)((n_ent + 3) & ~(int64_t)3);
            if (out->blob.size() + words > ((size_t)1 << 33)) return false;    // 16-byte offsets in 32 bits
            base[pn] = (int64_t)out->blob.size();
            out->blob.resize(out->blob.size() + words, 0u);
            uint16_t* p16 = reinterpret_cast<uint16_t*>(&out->blob[(size_t)base[pn]]);
            for (int i = 0; i <= CH; ++i) p16[i] = (uint16_t)cp[i];
            out->max_seg_bytes = std::max(out->max_seg_bytes, (int)(words * 4));
        }
        // fill: entries of a major index in increasing minor order (cp[i] = running write position)
        for (int64_t j = 0; j < n_here; ++j) {
            const int64_t m = dealt ? j * n_chunks + ch : (int64_t)ch * CH + j;
            const int slot = dealt ? blk_slot_of(j) : (int)j;
            for (int64_t k = (int64_t)ptr[m]; k < (int64_t)ptr[m + 1]; ++k) {
                const int pn = idx[k] / PN;
                int32_t* cp = &cnt[(size_t)pn * (CH + 1)];
                const int nt = blk_terms((uint32_t)val[k], terms);
                uint32_t* ent = &out->blob[(size_t)base[pn] + ptr_words];
                for (int q = 0; q < nt; ++q)
                    ent[cp[slot]++] = (uint32_t)((idx[k] - pn * PN) * 512) | (blk_hi16(terms[q]) << 16);
            }
        }
    }
    out->seg_off[(size_t)n_chunks * n_panels] = (uint32_t)(out->blob.size() / 4);
    if (out->blob.empty()) out->blob.resize(4, 0u);
    return true;
}

// blocked copy of the TRANSPOSE of a CSR matrix (major = column): counting sort, rows stay increasing inside a column
inline bool build_blk_transposed(int64_t R, int64_t V, int64_t nnz, const int32_t* row_ptr, const int32_t* col, const int32_t* val,
                                 BlkHost* out) {
    std::vector<int64_t> cp((size_t)V + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) cp[(size_t)col[k] + 1]++;
    for (int64_t v = 0; v < V; ++v) cp[(size_t)v + 1] += cp[(size_t)v];
    std::vector<int32_t> tr((size_t)std::max<int64_t>(nnz, 1)), tv((size_t)std::max<int64_t>(nnz, 1));
    std::vector<int64_t> at(cp.begin(), cp.end() - 1);
    for (int64_t r = 0; r < R; ++r)
        for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
            const int64_t q = at[(size_t)col[k]]++;
            tr[(size_t)q] = (int32_t)r; tv[(size_t)q] = val[k];
        }
    return build_blk<int64_t>(V, R, cp.data(), tr.data(), tv.data(), out);
}

// The product out[major] = sum f * x[minor], walking the blob exactly as blk_kernel does (chunk, panel, major index,
// entries in stored order; the count rebuilt from its top 16 bits).  Test helper: the CPU suite checks the layout with it.
inline void blk_multiply_host(const BlkHost& b, int64_t n_major, const double* x, double* out) {
    for (int64_t m = 0; m < n_major; ++m) out[m] = 0.0;
    for (int ch = 0; ch < b.n_chunks; ++ch)
        for (int pn = 0; pn < b.n_panels; ++pn) {
            const size_t seg = (size_t)ch * b.n_panels + pn;
            const uint32_t o0 = b.seg_off[seg], o1 = b.seg_off[seg + 1];
            if (o0 == o1) continue;
            const uint32_t* words = &b.blob[(size_t)o0 * 4];
            const uint16_t* p16 = reinterpret_cast<const uint16_t*>(words);
            const uint32_t* ent = words + kBlkPtrBytes / 4;
            for (int i = 0; i < kBlkCH; ++i) {
                const int64_t m = blk_major_of(ch, i / kBlkRW, i % kBlkRW, b.n_chunks, b.dealt);
                for (int k = p16[i]; k < p16[i + 1]; ++k) {
                    const uint32_t e = ent[k];
                    const uint64_t bits = (uint64_t)(e & 0xffff0000u) << 32;
                    double f;
                    memcpy(&f, &bits, 8);
                    const int64_t minor = (int64_t)pn * kBlkPN + (e & 0xffffu) / 512;
                    if (m < n_major) out[m] += f * x[minor];
                }
            }
        }
}

}  // namespace tvs
