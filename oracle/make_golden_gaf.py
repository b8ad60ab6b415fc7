#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/gaf_records.json.gz.

Runs the UNMODIFIED parse workers of the reference (``_gaf_parse_worker``, ``_tsv_parse_worker`` and the record
constructors they call, traversome/GraphAlignRecords.py:27-196) on (1) the simulated plastome alignments of the cfg1
fixture (tests/golden/plastome_sim.gaf.gz, 2,009 lines), (2) hand-written GAF lines that exercise the optional tags,
stable-path coordinates, orientation prefixes, CRLF / padded lines, the identity filter and CIGAR splitting, (3) SPAligner
TSV lines, (4) malformed lines, for which the exception type the reference raises is recorded.
Run HERE, once (needs /root/reference); the output is committed.
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = "/root/reference"

EDGE_GAF = "".join([
    "readA\t1000\t10\t990\t+\t>s1<s2>s3\t5000\t100\t1080\t950\t980\t60\tNM:i:30\tid:f:0.969388\tcg:Z:500M2I478M\n",
    "readB\t200\t0\t200\t-\t<utg000001l:0-200\t200\t0\t200\t180\t200\t0\n",
    "readC 1\t300\t5\t295\t+\tchr1\t1000\t5\t295\t290\t290\t255\tid:i:1\tdv:f:0.0\ttp:A:P\n",
    "readD\t50\t0\t50\t+\t>a>b\t100\t0\t50\t20\t50\t3\tcg:Z:50M\r\n",
    "  readE\t10\t0\t10\t+\t>x\t10\t0\t10\t10\t10\t1  \n",
    "readF\t+7\t0\t1_000\t-\t>s1:5-9<s2:1-2\t2000\t 3\t1003 \t900\t1000\t9\tid:f:nan\n",
    "readG\t10\t0\t10\t+\t>s10>s1\t20\t0\t10\t9\t10\t0\tid:f:9.5e-1\tzz:Z:a:b:c\tcg:Z:3=1X6=\n",
    "readH\t10\t0\t10\t+\t\t20\t0\t10\t10\t10\t0",
])
EDGE_SPA = "".join([
    "name1\t0\t2491\t536\t1142\t2491\t44+,24+,22+,1+,38-\t909,4,115,1,1142\tACGT\n",
    "name2\t0,100\t50,200\t1,2\t3,4\t300\t1+;2-\t50;100\tAC;GT\n",
    "name3\t100\t5\t7\t9\t120\t7-\t96\n",
    " name4\t1\t2\t3\t4\t5\tedge_12+,e-\t1,1\t\r\n",
])
BAD_GAF = [
    "r\t10\t0\t10\t+\t>x\t10\t0\t10\t10\t10",                              # 11 columns
    "r\t10\t0\t10\t+\t>x\t10\t0\tten\t10\t10\t1",                         # non-integer
    "r\t10\t0\t10\t*\t>x\t10\t0\t10\t10\t10\t1",                          # strand
    "r\t10\t0\t10\t+\t>x\t10\t0\t10\t10\t10\t1\tNM5",                     # tag without colons
    "r\t10\t0\t10\t+\t>x\t10\t0\t10\t10\t10\t1\tNM:i:x",                  # integer tag, bad value
    "r\t10\t0\t10\t+\t>x\t10\t0\t10\t10\t10\t1\tid:Z:high",               # identity that cannot be compared
    "r\t10\t0\t10\t+\t>x\t10\t0\t10\t10\t0\t1",                           # align_len 0
    "",                                                                   # blank line
]
BAD_SPA = [
    "name",                                                               # one column
    "name\t0\t10\t5\t9\t10\t7\t96",                                       # path step without strand
    "name\t0\t10\t5\t9\t10\t7+\tx",                                       # non-integer length
    "name\t0\t10\t5",                                                     # too few columns
]


def gaf_dict(r):
    return {"query_name": r.query_name, "query_len": r.query_len, "q_start": r.q_start, "q_end": r.q_end, "q_strand": r.q_strand,
            "path": [list(p) for p in r.path], "p_len": r.p_len, "p_start": r.p_start, "p_end": r.p_end, "p_align_len": r.p_align_len,
            "num_match": r.num_match, "align_len": r.align_len, "align_quality": r.align_quality,
            "identity": (repr(r.identity) if r.identity != r.identity else r.identity), "optional_fields": {
                k: (repr(v) if isinstance(v, float) and v != v else v) for k, v in r.optional_fields.items()},
            "cigar": None if r.cigar is None else [list(c) for c in r.cigar]}


def spa_dict(r):
    return {"query_name": r.query_name, "q_start": r.q_start, "q_end": r.q_end, "p_start": r.p_start, "query_len": r.query_len,
            "q_align_len": r.q_align_len, "path": [list(p) for p in r.path], "path_align_lengths": r.path_align_lengths,
            "q_strand": r.q_strand, "p_align_len": r.p_align_len, "p_end": r.p_end, "align_len": r.align_len}


def main():
    sys.path.insert(0, REFERENCE)
    from traversome import GraphAlignRecords as G
    out = {}
    with gzip.open(os.path.join(GOLDEN, "plastome_sim.gaf.gz"), "rt") as h:
        lines = h.readlines()
    recs, n = G._gaf_parse_worker(lines, 0.0, False)
    out["plastome"] = {"n_file_records": n, "records": [gaf_dict(r) for r in recs]}
    for thr, cg in ((0.0, False), (0.0, True), (0.95, True), (1.0, False)):
        recs, n = G._gaf_parse_worker(EDGE_GAF.splitlines(True), thr, cg)
        out["edge_gaf_%g_%d" % (thr, int(cg))] = {"min_record_identity": thr, "parse_cigar": cg, "n_file_records": n,
                                                  "records": [gaf_dict(r) for r in recs]}
    out["edge_gaf_text"] = EDGE_GAF
    recs, n = G._tsv_parse_worker(EDGE_SPA.splitlines(True))
    out["edge_spa"] = {"text": EDGE_SPA, "n_file_records": n, "records": [spa_dict(r) for r in recs]}
    bad = []
    for line in BAD_GAF:
        try:
            G._gaf_parse_worker([line + "\n"], 0.0, False)
            bad.append([line, None])
        except Exception as e:
            bad.append([line, type(e).__name__])
    out["bad_gaf"] = bad
    bad = []
    for line in BAD_SPA:
        try:
            G._tsv_parse_worker([line + "\n"])
            bad.append([line, None])
        except Exception as e:
            bad.append([line, type(e).__name__])
    out["bad_spa"] = bad
    path = os.path.join(GOLDEN, "gaf_records.json.gz")
    with gzip.open(path, "wt") as h:
        json.dump(out, h)
    print("wrote", path, {k: (len(v["records"]) if isinstance(v, dict) and "records" in v else v if k.startswith("bad") else "...") for k, v in out.items()})


if __name__ == "__main__":
    main()
