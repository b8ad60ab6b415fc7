"""CPU (host code through the C ABI, include/tvs_gaf.h): the alignment-file parser against vectors made by the reference's
own parse workers (oracle/make_golden_gaf.py -> tests/golden/gaf_records.json.gz): the cfg1 fixture's 2,009 simulated
alignments, hand-written lines for tags / stable-path coordinates / orientation / padding / the identity filter / CIGARs,
SPAligner TSV, and malformed lines (the reference raises; the library fails with the line number).
Reference: GraphAlignRecords.py:27-196 (records, workers), utils.py:1065-1082 (path strings), :1024-1134 (file loop)."""
import gzip
import json
import math
import os

import numpy as np
import pytest

from oracle import gaf_port
from traversome_b200 import _cabi, gaf

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def golden(lib_built):
    with gzip.open(os.path.join(GOLDEN, "gaf_records.json.gz"), "rt") as h:
        return json.load(h)


def _norm(v):
    if isinstance(v, float) and math.isnan(v):
        return "nan"
    if isinstance(v, (tuple, list)):
        return [_norm(x) for x in v]
    if isinstance(v, dict):
        return {k: _norm(x) for k, x in v.items()}
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.bool_,)):
        return bool(v)
    return v


def _gaf_view(r):
    return {k: _norm(getattr(r, k)) for k in gaf_port.GAF_FIELDS}


def _spa_view(r):
    return {k: _norm(getattr(r, k)) for k in gaf_port.SPA_FIELDS}


def _same(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        w = _norm(w)
        for k, wv in w.items():
            assert g[k] == wv, (k, g[k], wv, w["query_name"])
            if k != "identity":                                 # "id:i:1" is an int in the reference, a float here: equal, not identical
                assert type(g[k]) is type(wv), (k, type(g[k]), type(wv))


def test_plastome_alignments_equal_the_reference_records(golden):
    with gzip.open(os.path.join(GOLDEN, "plastome_sim.gaf.gz"), "rb") as h:
        text = h.read()
    want = golden["plastome"]
    table = gaf.GafTable(text)
    assert table.n_file_records == want["n_file_records"] == 2009 and table.n_records == len(want["records"])
    _same([_gaf_view(r) for r in table.records()], want["records"])
    port, n = gaf_port.parse_gaf(text.decode().splitlines(True))
    assert n == want["n_file_records"]
    _same([{k: _norm(r[k]) for k in gaf_port.GAF_FIELDS} for r in port], want["records"])
    # the columns themselves
    assert table.line_no.tolist() == list(range(2009)) and sorted(table.segment_names) == ["IR", "LSC", "SSC"]
    assert table.p_align_len.tolist() == [r["p_align_len"] for r in want["records"]]
    assert np.array_equal(np.diff(table.step_ptr), [len(r["path"]) for r in want["records"]])


@pytest.mark.parametrize("key", ["edge_gaf_0_0", "edge_gaf_0_1", "edge_gaf_0.95_1", "edge_gaf_1_0"])
def test_hand_written_gaf_lines(golden, key):
    want = golden[key]
    text = golden["edge_gaf_text"]
    table = gaf.GafTable(text, min_record_identity=want["min_record_identity"], parse_cigar=want["parse_cigar"], n_threads=1)
    assert table.n_file_records == want["n_file_records"]
    _same([_gaf_view(r) for r in table.records(parse_cigar=want["parse_cigar"])], want["records"])
    port, n = gaf_port.parse_gaf(text.splitlines(True), want["min_record_identity"], want["parse_cigar"])
    _same([{k: _norm(r[k]) for k in gaf_port.GAF_FIELDS} for r in port], want["records"])
    assert table.line_no.tolist() == [r["line_no"] for r in port]


def test_spaligner_lines(golden):
    want = golden["edge_spa"]
    table = gaf.GafTable(want["text"], alignment_format="SPA-TSV")
    assert table.n_file_records == want["n_file_records"] == 4 and table.n_records == 3
    _same([_spa_view(r) for r in table.records()], want["records"])
    port, n = gaf_port.parse_spa(want["text"].splitlines(True))
    assert n == 4
    _same([{k: _norm(r[k]) for k in gaf_port.SPA_FIELDS} for r in port], want["records"])


def test_lines_the_reference_raises_on_fail_with_their_line_number(golden):
    good = "ok\t10\t0\t10\t+\t>x\t10\t0\t10\t10\t10\t1\n"
    for line, exc in golden["bad_gaf"]:
        assert exc is not None                                   # the reference raised on every one of them
        with pytest.raises(_cabi.TvsError) as e:
            gaf.GafTable(good * 3 + line + "\n" + good)
        assert e.value.code == _cabi.TVS_EINVAL and "line 4:" in str(e.value), str(e.value)
        with pytest.raises(Exception):
            gaf_port.parse_gaf([line + "\n"])
    for line, exc in golden["bad_spa"]:
        assert exc is not None
        with pytest.raises(_cabi.TvsError) as e:
            gaf.GafTable("n\t0\t9\t0\t9\t10\t1+\t10\n" + line + "\n", alignment_format="SPA-TSV")
        assert "line 2:" in str(e.value)
    with pytest.raises(Exception):
        gaf.GafTable("x", alignment_format="GFA")


def test_empty_input_and_missing_final_newline(golden):
    t = gaf.GafTable(b"")
    assert t.n_file_records == 0 and t.n_records == 0 and t.records() == [] and t.paths() == []
    one = "q\t10\t0\t10\t+\t>a<b\t30\t5\t15\t9\t10\t7"
    t = gaf.GafTable(one)                                        # no '\\n' at the end of the file
    assert t.n_file_records == 1 and t.records()[0].path == (("a", True), ("b", False)) and t.records()[0].identity == 0.9
    t = gaf.GafTable(one + "\n" + one + "\n", min_record_identity=0.95)
    assert t.n_file_records == 2 and t.n_records == 0


def test_threads_and_blocks_give_the_single_thread_table(golden):
    with gzip.open(os.path.join(GOLDEN, "plastome_sim.gaf.gz"), "rb") as h:
        base = h.read().decode().splitlines(True)
    lines = []
    for rep in range(60):                                        # ~7 MB: several blocks per thread
        for i, l in enumerate(base):
            f = l.split("\t")
            f[0] = "%s_%d" % (f[0], rep)
            f[9] = str(int(f[9]) - (i + rep) % 50)               # identities below 1 for some records
            f[12] = "NM:i:%d\n" % ((i + rep) % 50)               # no id tag: identity = num_match / align_len
            lines.append("\t".join(f))
    text = "".join(lines).encode()
    one = gaf.GafTable(text, min_record_identity=0.999, n_threads=1)
    many = gaf.GafTable(text, min_record_identity=0.999, n_threads=7)
    assert one.n_file_records == many.n_file_records == len(lines) and 0 < one.n_records < len(lines)
    for k in ("line_no", "query_len", "q_start", "q_end", "p_len", "p_start", "p_end", "num_match", "align_len", "align_quality",
              "q_strand", "identity", "step_ptr", "step_forward", "step_segment"):
        assert np.array_equal(getattr(one, k), getattr(many, k)), k
    assert one.query_names() == many.query_names() and one.segment_names == many.segment_names
    want = [i for i, l in enumerate(lines) if int(l.split("\t")[9]) / float(l.split("\t")[10]) >= 0.999]
    assert one.line_no.tolist() == want
    # a malformed line deep inside: its global line number is reported whichever block it fell into
    bad = len(lines) * 2 // 3
    broken = "".join(lines[:bad] + ["oops\t1\n"] + lines[bad:]).encode()
    with pytest.raises(_cabi.TvsError) as e:
        gaf.GafTable(broken, n_threads=5)
    assert "line %d:" % (bad + 1) in str(e.value)


class _Records(object):
    """What GraphAlignRecords.__init__ sets before it calls parse_alignment_file (GraphAlignRecords.py:408-434)."""

    def __init__(self, path, fmt="GAF", min_record_identity=0.0, parse_cigar=False):
        self.alignment_file, self.alignment_format = path, fmt
        self.min_record_identity, self.parse_cigar = min_record_identity, parse_cigar
        self.raw_records, self.n_file_records = [], 0

    def parse_alignment_file(self, num_proc=1, _num_block_lines=10000):
        raise AssertionError("not swapped")


def test_the_swapped_method_fills_raw_records(golden, tmp_path):
    path = str(tmp_path / "edge.gaf")
    with open(path, "w") as h:
        h.write(golden["edge_gaf_text"])
    gaf.install_on_records(_Records)
    try:
        obj = _Records(path, min_record_identity=0.95, parse_cigar=True)
        obj.parse_alignment_file(num_proc=2)
        want = golden["edge_gaf_0.95_1"]
        assert obj.n_file_records == want["n_file_records"]
        _same([_gaf_view(r) for r in obj.raw_records], want["records"])
        gz = str(tmp_path / "edge.gaf.gz")
        with gzip.open(gz, "wt") as h:
            h.write(golden["edge_gaf_text"])
        obj2 = _Records(gz)
        obj2.parse_alignment_file()
        assert len(obj2.raw_records) == len(golden["edge_gaf_0_0"]["records"])
        obj2.raw_records[0].query_name = "renamed_block1"        # the reference's filters rename and delete records
        del obj2.raw_records[1]
    finally:
        gaf.uninstall_from_records()
    with pytest.raises(AssertionError):
        _Records(path).parse_alignment_file()


def test_reference_class_with_the_native_parser_keeps_its_read_records(golden, tmp_path):
    """The reference's GraphAlignRecords end to end (parse -> build_read_records -> filter_read_records) with and without
    the swap: same raw records, same reads kept.  Needs /root/reference (skipped on the GPU box)."""
    if not os.path.isdir("/root/reference/traversome"):
        pytest.skip("reference not present")
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from traversome.GraphAlignRecords import GraphAlignRecords
    finally:
        sys.path.remove("/root/reference")
    path = str(tmp_path / "sim.gaf")
    with gzip.open(os.path.join(GOLDEN, "plastome_sim.gaf.gz"), "rb") as i, open(path, "wb") as o:
        o.write(i.read())
    ref = GraphAlignRecords(path, alignment_format="GAF", min_align_len=100, min_identity=0.9)
    gaf.install_on_records(GraphAlignRecords)
    try:
        ours = GraphAlignRecords(path, alignment_format="GAF", min_align_len=100, min_identity=0.9)
    finally:
        gaf.uninstall_from_records()
    assert ours.n_file_records == ref.n_file_records and len(ours.raw_records) == len(ref.raw_records)
    for a, b in zip(ours.raw_records, ref.raw_records):
        for k in ("query_name", "query_len", "q_start", "q_end", "q_strand", "path", "p_len", "p_start", "p_end", "p_align_len",
                  "num_match", "align_len", "align_quality", "identity", "cigar", "optional_fields"):
            assert getattr(a, k) == getattr(b, k), k
    assert list(ours.read_records) == list(ref.read_records)
    assert [r.raw_ids for r in ours.read_records.values()] == [r.raw_ids for r in ref.read_records.values()]


class ReadRecord(object):
    """Stand-in with the fields of the reference's ReadRecord (GraphAlignRecords.py:200-216): build_read_records looks the
    class up in the module of the object it is bound to -- here this test module."""

    def __init__(self):
        self.raw_ids, self.records, self.p_align_len, self._p_identity_sum, self.p_identity, self.q_seq = [], [], 0, 0., 0., None


def _multi_hit_gaf():
    """GAF text whose reads have one to three alignments, listed out of query order and interleaved with other reads."""
    rows = []
    for i in range(60):
        n_hits = 1 + i % 3
        for h in range(n_hits):
            q_start = (n_hits - 1 - h) * 1000 + 7 * i                       # later hits first
            length = 400 + 13 * ((i * 7 + h * 3) % 50)
            n_match = length - (i + 2 * h) % 9
            rows.append((i, h, "read%d\t5000\t%d\t%d\t+\t>s%d<s%d\t9000\t%d\t%d\t%d\t%d\t60" % (
                i, q_start, q_start + length, 1 + i % 4, 2 + h, 100 + h, 100 + h + length, n_match, length)))
    rows.sort(key=lambda r: (r[1], r[0]))                                    # all first hits, then all second hits, ...
    return "".join(r[2] + "\n" for r in rows)


def test_build_read_records_follows_the_reference_loop():
    """The swapped build_read_records against a restatement of the reference's loop (GraphAlignRecords.py:225-230 append,
    :252-256 sort_by, :936-950): grouping in order of first appearance, sums in raw order, records sorted by
    (q_start, q_end, cigar)."""
    text = _multi_hit_gaf().encode()
    obj = _Records("unused")
    obj.raw_records = gaf.GafTable(text, n_threads=1).records()
    gaf.build_read_records(obj)
    want = {}
    for go_r, rec in enumerate(obj.raw_records):
        rr = want.setdefault(rec.query_name, ReadRecord())
        rr.raw_ids.append(go_r); rr.records.append(rec)
        rr.p_align_len += rec.p_align_len
        rr._p_identity_sum += rec.p_align_len * rec.identity
        rr.p_identity = rr._p_identity_sum / rr.p_align_len
    for rr in want.values():
        order = sorted(range(len(rr.raw_ids)), key=lambda x: [rr.records[x].q_start, rr.records[x].q_end, rr.records[x].cigar])
        rr.records, rr.raw_ids = [rr.records[x] for x in order], [rr.raw_ids[x] for x in order]
    got = obj.read_records
    assert list(got) == list(want) and len(got) == 60
    assert sum(len(r.raw_ids) > 1 for r in got.values()) == 40
    for name in want:
        a, b = got[name], want[name]
        assert isinstance(a, ReadRecord) and a.raw_ids == b.raw_ids and [id(r) for r in a.records] == [id(r) for r in b.records]
        assert a.p_align_len == b.p_align_len and a._p_identity_sum == b._p_identity_sum and a.p_identity == b.p_identity
        assert a.q_seq is None
        assert [r.q_start for r in a.records] == sorted(r.q_start for r in a.records)


def test_reference_read_records_with_and_without_the_swaps(tmp_path):
    """The reference's own build_read_records / filter_read_records on a file with multi-hit reads: identical read
    records (ids, order, sums, identities) and identical reads kept.  Needs /root/reference."""
    if not os.path.isdir("/root/reference/traversome"):
        pytest.skip("reference not present")
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from traversome.GraphAlignRecords import GraphAlignRecords
    finally:
        sys.path.remove("/root/reference")
    path = str(tmp_path / "multi.gaf")
    with open(path, "w") as h:
        h.write(_multi_hit_gaf())
    for min_identity in (0.9, 0.992):
        ref = GraphAlignRecords(path, alignment_format="GAF", min_align_len=450, min_identity=min_identity)
        gaf.install_on_records(GraphAlignRecords)
        try:
            ours = GraphAlignRecords(path, alignment_format="GAF", min_align_len=450, min_identity=min_identity)
        finally:
            gaf.uninstall_from_records()
        assert GraphAlignRecords.build_read_records is not gaf.build_read_records
        assert len(ours.raw_records) == len(ref.raw_records) and list(ours.read_records) == list(ref.read_records)
        assert list(ref.read_records) != ["read%d" % i for i in range(60)]      # the filters did drop / split reads
        for a, b in zip(ours.read_records.values(), ref.read_records.values()):
            assert a.raw_ids == b.raw_ids and a.p_align_len == b.p_align_len
            assert a._p_identity_sum == b._p_identity_sum and a.p_identity == b.p_identity
            assert [(r.query_name, r.q_start, r.q_end, r.p_start, r.p_end) for r in a.records] == \
                   [(r.query_name, r.q_start, r.q_end, r.p_start, r.p_end) for r in b.records]
