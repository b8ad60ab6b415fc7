// tvs_gaf.cpp -- alignment-file parser in array form (include/tvs_gaf.h).  Host code.
//
// The text is cut into blocks at line boundaries, every block is parsed by one thread into its own column vectors and
// string pools, and the blocks are concatenated in file order.  The per-line rules restate, field by field, what the
// reference's constructors do (GraphAlignRecords.py:33-69 GAFRecord, :89-148 SPATSVRecord, utils.py:1065-1082
// gaf_str_to_path, :151-196 the two parse workers); anything they would raise on is an error here, with the line number.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <unordered_map>
#include <string_view>

#include "../../include/tvs_gaf.h"
#include "tvs_internal.h"

namespace {

struct Block {
    // input
    const char* begin = nullptr; const char* end = nullptr;
    // output
    int64_t n_lines = 0;
    std::vector<int64_t> line, query_len, q_start, q_end, p_len, p_start, p_end, num_match, align_len, mapq, tag_off, tag_end;
    std::vector<int8_t> strand;
    std::vector<double> identity;
    std::vector<int64_t> n_steps, name_len, cigar_len;       // per record
    std::vector<int8_t> step_fw;
    std::vector<int64_t> seg_len, extra;                      // per step
    std::string names, segs, cigars;
    int64_t err_line = -1; std::string err;
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// Python int(str): optional surrounding whitespace, optional sign, decimal digits with single underscores between them
bool parse_int(const char* a, const char* b, int64_t* out) {
    while (a < b && is_space(*a)) ++a;
    while (b > a && is_space(b[-1])) --b;
    if (a == b) return false;
    bool neg = false;
    if (*a == '+' || *a == '-') { neg = *a == '-'; ++a; }
    if (a == b || *a < '0' || *a > '9') return false;
    uint64_t v = 0;
    bool prev_us = false;
    for (; a < b; ++a) {
        if (*a == '_') { if (prev_us) return false; prev_us = true; continue; }
        if (*a < '0' || *a > '9') return false;
        prev_us = false;
        if (v > (uint64_t)std::numeric_limits<int64_t>::max() / 10) return false;   // beyond int64: not a coordinate
        v = v * 10 + (uint64_t)(*a - '0');
        if (v > (uint64_t)std::numeric_limits<int64_t>::max()) return false;
    }
    if (prev_us) return false;
    *out = neg ? -(int64_t)v : (int64_t)v;
    return true;
}

// Python float(str) for the forms alignment tools write (decimal, exponent, inf, nan; surrounding whitespace)
bool parse_float(const char* a, const char* b, double* out) {
    while (a < b && is_space(*a)) ++a;
    while (b > a && is_space(b[-1])) --b;
    if (a == b || b - a > 63) return false;
    char buf[64];
    memcpy(buf, a, (size_t)(b - a));
    buf[b - a] = 0;
    for (const char* p = buf; *p; ++p)
        if (*p == 'x' || *p == 'X' || *p == 'p' || *p == 'P') return false;        // hex floats: strtod takes them, Python does not
    char* endp = nullptr;
    errno = 0;
    const double v = strtod(buf, &endp);
    if (endp == buf || *endp != 0) return false;
    *out = v;
    return true;
}

struct Fields {
    const char* a[16]; const char* b[16]; int n = 0;          // the first 12 (GAF) / 9 (SPA) columns
    const char* rest = nullptr;                                // start of column 13 (GAF tags), or null
};

void fail(Block& blk, int64_t line, const std::string& msg) {
    if (blk.err_line < 0) { blk.err_line = line; blk.err = msg; }
}

// gaf_str_to_path (utils.py:1065-1082): re.findall(r".[^\s><]*"), '>' forward, '<' reverse, no prefix forward; name up to ':'
void parse_gaf_path(Block& blk, const char* a, const char* b, int64_t* n_steps) {
    int64_t n = 0;
    const char* p = a;
    while (p < b) {
        const char* s = p++;
        while (p < b && !is_space(*p) && *p != '>' && *p != '<') ++p;
        bool fw = true;
        const char* name = s;
        if (*s == '>') name = s + 1;
        else if (*s == '<') { name = s + 1; fw = false; }
        const char* colon = (const char*)memchr(name, ':', (size_t)(p - name));
        const char* name_end = colon ? colon : p;
        blk.step_fw.push_back(fw ? 1 : 0);
        blk.seg_len.push_back((int64_t)(name_end - name));
        blk.extra.push_back(0);
        blk.segs.append(name, (size_t)(name_end - name));
        ++n;
    }
    *n_steps = n;
}

void parse_gaf_line(Block& blk, const char* text0, const char* a, const char* b, int64_t line, double min_identity, bool keep_cigar) {
    Fields f;
    const char* p = a;
    while (f.n < 12) {
        const char* t = (const char*)memchr(p, '\t', (size_t)(b - p));
        f.a[f.n] = p; f.b[f.n] = t ? t : b; ++f.n;
        if (!t) { p = b; break; }
        p = t + 1;
        if (f.n == 12) f.rest = p;
    }
    if (f.n < 12) { fail(blk, line, "fewer than 12 tab-separated columns"); return; }
    int64_t v[12] = {0};
    static const int int_cols[] = {1, 2, 3, 6, 7, 8, 9, 10, 11};
    for (int c : int_cols)
        if (!parse_int(f.a[c], f.b[c], &v[c])) { fail(blk, line, "column " + std::to_string(c + 1) + " is not an integer"); return; }
    if (f.b[4] - f.a[4] != 1 || (*f.a[4] != '+' && *f.a[4] != '-')) { fail(blk, line, "strand (column 5) is neither '+' nor '-'"); return; }
    // optional tags: "xx:T:value" (split(":", maxsplit=2) must give three parts); i / f / Z are kept by the reference
    bool have_id = false;
    double id_val = 0.0;
    const char* cg_a = nullptr; const char* cg_b = nullptr;
    if (f.rest) {
        const char* q = f.rest;
        while (true) {
            const char* t = (const char*)memchr(q, '\t', (size_t)(b - q));
            const char* e = t ? t : b;
            const char* c1 = (const char*)memchr(q, ':', (size_t)(e - q));
            const char* c2 = c1 ? (const char*)memchr(c1 + 1, ':', (size_t)(e - c1 - 1)) : nullptr;
            if (!c2) { fail(blk, line, "optional field without tag:type:value"); return; }
            const bool one_char_type = (c2 - c1) == 2;
            const char ty = one_char_type ? c1[1] : 0;
            const bool is_id = (c1 - q) == 2 && q[0] == 'i' && q[1] == 'd';
            const bool is_cg = (c1 - q) == 2 && q[0] == 'c' && q[1] == 'g';
            if (ty == 'i') {
                int64_t iv;
                if (!parse_int(c2 + 1, e, &iv)) { fail(blk, line, "integer tag with a non-integer value"); return; }
                if (is_id) { have_id = true; id_val = (double)iv; }
                if (is_cg) { cg_a = nullptr; cg_b = nullptr; }
            } else if (ty == 'f') {
                double fv;
                if (!parse_float(c2 + 1, e, &fv)) { fail(blk, line, "float tag with a non-numeric value"); return; }
                if (is_id) { have_id = true; id_val = fv; }
                if (is_cg) { cg_a = nullptr; cg_b = nullptr; }
            } else if (ty == 'Z') {
                if (is_id) { fail(blk, line, "id tag of type Z cannot be compared with the identity threshold"); return; }
                if (is_cg) { cg_a = c2 + 1; cg_b = e; }
            }                                                  // other types: not stored by the reference
            if (!t) break;
            q = t + 1;
        }
    }
    if (v[10] == 0) { fail(blk, line, "alignment block length (column 11) is 0: num_match / align_len"); return; }
    const double identity = have_id ? id_val : (double)v[9] / (double)v[10];
    if (!(identity >= min_identity)) return;                   // GraphAlignRecords.py:169 (a NaN identity is dropped there too)
    blk.line.push_back(line);
    blk.query_len.push_back(v[1]); blk.q_start.push_back(v[2]); blk.q_end.push_back(v[3]);
    blk.strand.push_back(*f.a[4] == '+' ? 1 : 0);
    blk.p_len.push_back(v[6]); blk.p_start.push_back(v[7]); blk.p_end.push_back(v[8]);
    blk.num_match.push_back(v[9]); blk.align_len.push_back(v[10]); blk.mapq.push_back(v[11]);
    blk.identity.push_back(identity);
    blk.tag_off.push_back(f.rest ? (int64_t)(f.rest - text0) : -1);
    blk.tag_end.push_back((int64_t)(b - text0));
    blk.name_len.push_back((int64_t)(f.b[0] - f.a[0]));
    blk.names.append(f.a[0], (size_t)(f.b[0] - f.a[0]));
    int64_t ns = 0;
    parse_gaf_path(blk, f.a[5], f.b[5], &ns);
    blk.n_steps.push_back(ns);
    if (keep_cigar && cg_a) { blk.cigar_len.push_back((int64_t)(cg_b - cg_a)); blk.cigars.append(cg_a, (size_t)(cg_b - cg_a)); }
    else blk.cigar_len.push_back(0);
}

// SPATSVRecord (GraphAlignRecords.py:109-131) + _tsv_parse_worker (:178-196)
void parse_spa_line(Block& blk, const char* text0, const char* a, const char* b, int64_t line) {
    Fields f;
    const char* p = a;
    while (f.n < 9) {
        const char* t = (const char*)memchr(p, '\t', (size_t)(b - p));
        f.a[f.n] = p; f.b[f.n] = t ? t : b; ++f.n;
        if (!t) break;
        p = t + 1;
    }
    if (f.n < 2) { fail(blk, line, "fewer than 2 tab-separated columns"); return; }
    if (memchr(f.a[1], ',', (size_t)(f.b[1] - f.a[1]))) return;      // several non-overlapping sub-paths: skipped (:190-191)
    if (f.n < 8) { fail(blk, line, "fewer than 8 tab-separated columns"); return; }
    int64_t qs, qe, ps, ql;
    if (!parse_int(f.a[1], f.b[1], &qs) || !parse_int(f.a[2], f.b[2], &qe) || !parse_int(f.a[3], f.b[3], &ps) || !parse_int(f.a[5], f.b[5], &ql)) {
        fail(blk, line, "a coordinate column is not an integer"); return;
    }
    // path "44+,24+,22+": name = all but the last character, strand = the last
    const size_t step0 = blk.step_fw.size();
    int64_t ns = 0;
    for (const char* s = f.a[6];;) {
        const char* c = (const char*)memchr(s, ',', (size_t)(f.b[6] - s));
        const char* e = c ? c : f.b[6];
        if (e == s || (e[-1] != '+' && e[-1] != '-')) { fail(blk, line, "path step without a '+' / '-' suffix"); return; }
        blk.step_fw.push_back(e[-1] == '+' ? 1 : 0);
        blk.seg_len.push_back((int64_t)(e - 1 - s));
        blk.segs.append(s, (size_t)(e - 1 - s));
        blk.extra.push_back(0);
        ++ns;
        if (!c) break;
        s = c + 1;
    }
    int64_t nl = 0, sum = 0;
    for (const char* s = f.a[7];;) {
        const char* c = (const char*)memchr(s, ',', (size_t)(f.b[7] - s));
        const char* e = c ? c : f.b[7];
        int64_t lv;
        if (!parse_int(s, e, &lv)) { fail(blk, line, "aligned length on a path step is not an integer"); return; }
        if (nl < ns) blk.extra[step0 + (size_t)nl] = lv;
        sum += lv; ++nl;
        if (!c) break;
        s = c + 1;
    }
    // (the reference does not compare the two list lengths; extra keeps what pairs up, the sum takes every value)
    const int64_t q_align = (qs > qe ? qs - qe : qe - qs) + 1;
    blk.line.push_back(line);
    blk.query_len.push_back(ql); blk.q_start.push_back(qs); blk.q_end.push_back(qe);
    blk.strand.push_back(qe >= qs ? 1 : 0);
    blk.p_len.push_back(-1); blk.p_start.push_back(ps); blk.p_end.push_back(ps + sum);
    blk.num_match.push_back(-1); blk.align_len.push_back(std::max(sum, q_align)); blk.mapq.push_back(-1);
    blk.identity.push_back(std::numeric_limits<double>::quiet_NaN());
    blk.tag_off.push_back(-1); blk.tag_end.push_back((int64_t)(b - text0));
    blk.name_len.push_back((int64_t)(f.b[0] - f.a[0]));
    blk.names.append(f.a[0], (size_t)(f.b[0] - f.a[0]));
    blk.n_steps.push_back(ns);
    blk.cigar_len.push_back(0);
}

void parse_block(Block& blk, const char* text0, bool spa, double min_identity, bool keep_cigar) {
    const char* p = blk.begin;
    int64_t line = 0;
    while (p < blk.end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(blk.end - p));
        const char* e = nl ? nl : blk.end;
        const char* a = p; const char* b = e;
        while (a < b && is_space(*a)) ++a;                     // str.strip()
        while (b > a && is_space(b[-1])) --b;
        if (spa) parse_spa_line(blk, text0, a, b, line);
        else parse_gaf_line(blk, text0, a, b, line, min_identity, keep_cigar);
        ++line;
        if (blk.err_line >= 0) break;
        p = nl ? nl + 1 : blk.end;
    }
    blk.n_lines = line;
}

}  // namespace

struct tvs_gaf_table {
    int64_t n_lines = 0, n_records = 0, n_steps = 0;
    std::vector<int64_t> line, query_len, q_start, q_end, p_len, p_start, p_end, num_match, align_len, mapq, tag_off, tag_end;
    std::vector<int8_t> strand, step_fw;
    std::vector<double> identity;
    std::vector<int64_t> step_ptr, name_ptr, cigar_ptr, seg_ptr, extra;
    std::string names, segs, cigars;
    // distinct segment names in order of first appearance (built on first use by tvs_gaf_segments)
    bool interned = false;
    std::vector<int64_t> step_segment, dseg_ptr;
    std::string dsegs;
};

namespace {

template <typename T>
void append(std::vector<T>& dst, const std::vector<T>& src) { dst.insert(dst.end(), src.begin(), src.end()); }

int parse_text(const char* text, int64_t n_bytes, bool spa, double min_identity, int n_threads, bool keep_cigar, tvs_gaf_handle* out,
               const char* who) {
    using tvs::set_error;
    if (!out) { set_error("%s: out is null", who); return TVS_EINVAL; }
    *out = nullptr;
    if (n_bytes < 0 || (n_bytes > 0 && !text)) { set_error("%s: null text or negative size", who); return TVS_EINVAL; }
    if (n_threads <= 0) n_threads = (int)std::min(64u, std::max(1u, std::thread::hardware_concurrency()));
    // blocks of >= 1 MiB cut at line ends (a few per thread: lines differ in cost)
    const int64_t want = std::max<int64_t>(1, std::min<int64_t>((int64_t)n_threads * 4, n_bytes / (1 << 20) + 1));
    std::vector<Block> blocks;
    {
        const char* end = text + n_bytes;
        const char* p = text;
        const int64_t step = (n_bytes + want - 1) / want;
        while (p < end) {
            const char* q = p + std::min<int64_t>(step, end - p);
            if (q < end) {
                const char* nl = (const char*)memchr(q, '\n', (size_t)(end - q));
                q = nl ? nl + 1 : end;
            }
            Block b;
            b.begin = p; b.end = q;
            blocks.push_back(std::move(b));
            p = q;
        }
    }
    try {
        const int nt = (int)std::min<size_t>((size_t)n_threads, blocks.size());
        if (nt <= 1) {
            for (auto& b : blocks) parse_block(b, text, spa, min_identity, keep_cigar);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t)
                th.emplace_back([&, t]() {
                    for (size_t i = (size_t)t; i < blocks.size(); i += (size_t)nt) parse_block(blocks[i], text, spa, min_identity, keep_cigar);
                });
            for (auto& x : th) x.join();
        }
        int64_t line0 = 0;
        for (auto& b : blocks) {
            if (b.err_line >= 0) {
                set_error("%s: line %lld: %s", who, (long long)(line0 + b.err_line + 1), b.err.c_str());
                return TVS_EINVAL;
            }
            line0 += b.n_lines;
        }
        tvs_gaf_table* t = new tvs_gaf_table();
        t->step_ptr.push_back(0); t->name_ptr.push_back(0); t->cigar_ptr.push_back(0); t->seg_ptr.push_back(0);
        line0 = 0;
        for (auto& b : blocks) {
            for (int64_t l : b.line) t->line.push_back(line0 + l);
            append(t->query_len, b.query_len); append(t->q_start, b.q_start); append(t->q_end, b.q_end);
            append(t->p_len, b.p_len); append(t->p_start, b.p_start); append(t->p_end, b.p_end);
            append(t->num_match, b.num_match); append(t->align_len, b.align_len); append(t->mapq, b.mapq);
            append(t->tag_off, b.tag_off); append(t->tag_end, b.tag_end);
            append(t->strand, b.strand); append(t->identity, b.identity); append(t->step_fw, b.step_fw); append(t->extra, b.extra);
            for (int64_t n : b.n_steps) t->step_ptr.push_back(t->step_ptr.back() + n);
            for (int64_t n : b.name_len) t->name_ptr.push_back(t->name_ptr.back() + n);
            for (int64_t n : b.cigar_len) t->cigar_ptr.push_back(t->cigar_ptr.back() + n);
            for (int64_t n : b.seg_len) t->seg_ptr.push_back(t->seg_ptr.back() + n);
            t->names += b.names; t->segs += b.segs; t->cigars += b.cigars;
            line0 += b.n_lines;
            b = Block();                                       // release the block's memory as we go
        }
        t->n_lines = line0;
        t->n_records = (int64_t)t->line.size();
        t->n_steps = (int64_t)t->step_fw.size();
        *out = t;
    } catch (const std::bad_alloc&) {
        set_error("%s: out of host memory", who);
        return TVS_ENOMEM;
    }
    return TVS_OK;
}

template <typename T>
void copy_out(T* dst, const std::vector<T>& src) { if (dst && !src.empty()) memcpy(dst, src.data(), src.size() * sizeof(T)); }

}  // namespace

extern "C" {

int tvs_gaf_parse(const char* text, int64_t n_bytes, double min_identity, int n_threads, int keep_cigar, tvs_gaf_handle* out) {
    return parse_text(text, n_bytes, false, min_identity, n_threads, keep_cigar != 0, out, "tvs_gaf_parse");
}

int tvs_gaf_parse_spa(const char* text, int64_t n_bytes, int n_threads, tvs_gaf_handle* out) {
    return parse_text(text, n_bytes, true, 0.0, n_threads, false, out, "tvs_gaf_parse_spa");
}

int tvs_gaf_dims(tvs_gaf_handle t, int64_t* n_lines, int64_t* n_records, int64_t* n_steps, int64_t* name_bytes,
                 int64_t* seg_bytes, int64_t* cigar_bytes) {
    if (!t) { tvs::set_error("tvs_gaf_dims: null handle"); return TVS_EINVAL; }
    if (n_lines) *n_lines = t->n_lines;
    if (n_records) *n_records = t->n_records;
    if (n_steps) *n_steps = t->n_steps;
    if (name_bytes) *name_bytes = (int64_t)t->names.size();
    if (seg_bytes) *seg_bytes = (int64_t)t->segs.size();
    if (cigar_bytes) *cigar_bytes = (int64_t)t->cigars.size();
    return TVS_OK;
}

int tvs_gaf_columns(tvs_gaf_handle t, int64_t* line_no, int64_t* query_len, int64_t* q_start, int64_t* q_end,
                    int8_t* q_strand, int64_t* p_len, int64_t* p_start, int64_t* p_end, int64_t* num_match,
                    int64_t* align_len, int64_t* mapq, double* identity, int64_t* step_ptr, int64_t* name_ptr,
                    int64_t* cigar_ptr, int64_t* tag_off, int64_t* tag_end) {
    if (!t) { tvs::set_error("tvs_gaf_columns: null handle"); return TVS_EINVAL; }
    copy_out(line_no, t->line); copy_out(query_len, t->query_len); copy_out(q_start, t->q_start); copy_out(q_end, t->q_end);
    copy_out(q_strand, t->strand); copy_out(p_len, t->p_len); copy_out(p_start, t->p_start); copy_out(p_end, t->p_end);
    copy_out(num_match, t->num_match); copy_out(align_len, t->align_len); copy_out(mapq, t->mapq); copy_out(identity, t->identity);
    copy_out(step_ptr, t->step_ptr); copy_out(name_ptr, t->name_ptr); copy_out(cigar_ptr, t->cigar_ptr);
    copy_out(tag_off, t->tag_off); copy_out(tag_end, t->tag_end);
    return TVS_OK;
}

int tvs_gaf_steps(tvs_gaf_handle t, char* names, int8_t* step_forward, int64_t* seg_ptr, char* segs, int64_t* extra, char* cigars) {
    if (!t) { tvs::set_error("tvs_gaf_steps: null handle"); return TVS_EINVAL; }
    if (names && !t->names.empty()) memcpy(names, t->names.data(), t->names.size());
    copy_out(step_forward, t->step_fw); copy_out(seg_ptr, t->seg_ptr); copy_out(extra, t->extra);
    if (segs && !t->segs.empty()) memcpy(segs, t->segs.data(), t->segs.size());
    if (cigars && !t->cigars.empty()) memcpy(cigars, t->cigars.data(), t->cigars.size());
    return TVS_OK;
}

static void intern_segments(tvs_gaf_table* t) {
    if (t->interned) return;
    std::unordered_map<std::string_view, int64_t> ids;
    ids.reserve(1024);
    t->step_segment.resize((size_t)t->n_steps);
    t->dseg_ptr.assign(1, 0);
    const char* base = t->segs.data();
    for (int64_t i = 0; i < t->n_steps; ++i) {
        const std::string_view key(base + t->seg_ptr[(size_t)i], (size_t)(t->seg_ptr[(size_t)i + 1] - t->seg_ptr[(size_t)i]));
        auto it = ids.find(key);
        if (it == ids.end()) {
            it = ids.emplace(key, (int64_t)ids.size()).first;
            t->dsegs.append(key.data(), key.size());
            t->dseg_ptr.push_back((int64_t)t->dsegs.size());
        }
        t->step_segment[(size_t)i] = it->second;
    }
    t->interned = true;
}

int tvs_gaf_segments(tvs_gaf_handle t, int64_t* n_distinct, int64_t* distinct_bytes) {
    if (!t) { tvs::set_error("tvs_gaf_segments: null handle"); return TVS_EINVAL; }
    try {
        intern_segments(t);
    } catch (const std::bad_alloc&) {
        tvs::set_error("tvs_gaf_segments: out of host memory");
        return TVS_ENOMEM;
    }
    if (n_distinct) *n_distinct = (int64_t)t->dseg_ptr.size() - 1;
    if (distinct_bytes) *distinct_bytes = (int64_t)t->dsegs.size();
    return TVS_OK;
}

int tvs_gaf_segment_ids(tvs_gaf_handle t, int64_t* step_segment, int64_t* dseg_ptr, char* dsegs) {
    if (!t) { tvs::set_error("tvs_gaf_segment_ids: null handle"); return TVS_EINVAL; }
    if (!t->interned) { tvs::set_error("tvs_gaf_segment_ids: call tvs_gaf_segments first"); return TVS_ESTATE; }
    copy_out(step_segment, t->step_segment); copy_out(dseg_ptr, t->dseg_ptr);
    if (dsegs && !t->dsegs.empty()) memcpy(dsegs, t->dsegs.data(), t->dsegs.size());
    return TVS_OK;
}

int tvs_gaf_destroy(tvs_gaf_handle t) {
    delete t;
    return TVS_OK;
}

}  // extern "C"
