"""Alignment-file parsing in array form (the part BEFORE read-path construction).

Mirror of what ``GraphAlignRecords.parse_alignment_file`` leaves behind (traversome/GraphAlignRecords.py:1024-1134):
``raw_records`` -- one record object per kept alignment line -- and ``n_file_records``.  The reference builds every record
with a Python constructor (``GAFRecord.__init__`` :33-69, ``SPATSVRecord.__init__`` :109-131, ``gaf_str_to_path``
utils.py:1065-1082), in a process pool when ``num_proc > 1`` (the objects are pickled back).  Here the text of the file goes
to the native parser of ``libtvs.so`` once (``include/tvs_gaf.h``: threads over blocks of lines, integer columns and string
pools) and the record objects are made from the columns; the arrays themselves (``GafTable``) are what
``traversome_b200.readpaths.generate_read_paths`` needs.

``install_on_records(GraphAlignRecords)`` swaps ``parse_alignment_file`` and ``build_read_records`` (:936-950) of the
reference's class; the per-read filter that follows (``filter_read_records``, :952-1022) is the reference's and sees the same
objects.
"""
import ctypes
import gc
import gzip
import re

import numpy as np

from . import _cabi

CIGAR_ALPHA_REG = "([MIDNSHPX=])"        # GraphAlignRecords.py:24


class GAFRecord(object):
    """One alignment record with the attribute names of the reference's GAFRecord (GraphAlignRecords.py:27-87).
    ``optional_fields`` is read from the line's tag text on first use; ``path_str`` is rebuilt from ``path`` (coordinates of
    stable-path steps, which the reference drops when it builds ``path``, are not kept); ``identity`` is a float also when it
    came from an "id:i:" tag."""
    __slots__ = ("query_name", "query_len", "q_start", "q_end", "q_strand", "path", "p_len", "p_start", "p_end", "p_align_len",
                 "num_match", "align_len", "align_quality", "identity", "cigar", "_tags", "_fields")

    @property
    def optional_fields(self):
        if self._fields is None:
            fields = {}
            if self._tags:
                for flag_type_val in self._tags.split("\t"):
                    op_flag, op_type, op_val = flag_type_val.split(":", 2)
                    if op_type == "i":
                        fields[op_flag] = int(op_val)
                    elif op_type == "Z":
                        fields[op_flag] = op_val
                    elif op_type == "f":
                        fields[op_flag] = float(op_val)
            self._fields = fields
        return self._fields

    @property
    def path_str(self):
        return "".join((">" if forward else "<") + name for name, forward in self.path)

    def split_cigar_str(self):
        cigar_split = re.split(CIGAR_ALPHA_REG, self.optional_fields["cg"])[:-1]
        return [(int(cigar_split[i]), cigar_split[i + 1]) for i in range(0, len(cigar_split), 2)]


class SPATSVRecord(object):
    """One SPAligner record with the attribute names of the reference's SPATSVRecord (GraphAlignRecords.py:89-148)."""
    __slots__ = ("query_name", "q_start", "q_end", "p_start", "query_len", "q_align_len", "path", "path_align_lengths", "q_strand",
                 "p_align_len", "p_end", "align_len")

    @property
    def path_str(self):
        return ",".join(name + ("+" if forward else "-") for name, forward in self.path)


class GafTable(object):
    """The kept records of one alignment file as columns (numpy) + string pools.

    Attributes: ``n_file_records`` (lines read), ``n_records``; int64 columns ``line_no, query_len, q_start, q_end, p_len,
    p_start, p_end, num_match, align_len, align_quality``; ``q_strand`` (bool), ``identity`` (float64), ``p_align_len``;
    ``step_ptr`` [n+1] into ``step_forward`` (bool) / ``step_segment`` (index into ``segment_names``)."""

    def __init__(self, text, alignment_format="GAF", min_record_identity=0.0, parse_cigar=False, n_threads=0):
        if isinstance(text, str):
            text = text.encode()
        if alignment_format not in ("GAF", "SPA-TSV"):
            raise Exception("unsupported format!")
        self.alignment_format = alignment_format
        self._text = text
        lib = _cabi.load()
        h = ctypes.c_void_p()
        if alignment_format == "GAF":
            _cabi._check(lib.tvs_gaf_parse(text, len(text), float(min_record_identity), int(n_threads), 1 if parse_cigar else 0, ctypes.byref(h)))
        else:
            _cabi._check(lib.tvs_gaf_parse_spa(text, len(text), int(n_threads), ctypes.byref(h)))
        try:
            dims = [ctypes.c_int64() for _ in range(6)]
            _cabi._check(lib.tvs_gaf_dims(h, *[ctypes.byref(d) for d in dims]))
            n_lines, n, n_steps, name_bytes, seg_bytes, cigar_bytes = [int(d.value) for d in dims]
            self.n_file_records, self.n_records = n_lines, n
            col = {k: np.empty(n, dtype=np.int64) for k in ("line_no", "query_len", "q_start", "q_end", "p_len", "p_start", "p_end",
                                                           "num_match", "align_len", "align_quality", "tag_off", "tag_end")}
            strand = np.empty(n, dtype=np.int8)
            self.identity = np.empty(n, dtype=np.float64)
            self.step_ptr = np.empty(n + 1, dtype=np.int64)
            self._name_ptr = np.empty(n + 1, dtype=np.int64)
            self._cigar_ptr = np.empty(n + 1, dtype=np.int64)
            p = _cabi._ptr
            _cabi._check(lib.tvs_gaf_columns(h, p(col["line_no"]), p(col["query_len"]), p(col["q_start"]), p(col["q_end"]), p(strand),
                                             p(col["p_len"]), p(col["p_start"]), p(col["p_end"]), p(col["num_match"]), p(col["align_len"]),
                                             p(col["align_quality"]), p(self.identity), p(self.step_ptr), p(self._name_ptr),
                                             p(self._cigar_ptr), p(col["tag_off"]), p(col["tag_end"])))
            for k, v in col.items():
                setattr(self, k if not k.startswith("tag_") else "_" + k, v)
            self.q_strand = strand.astype(bool)
            self.p_align_len = self.p_end - self.p_start
            names = np.empty(max(name_bytes, 1), dtype=np.uint8)
            step_fw = np.empty(n_steps, dtype=np.int8)
            self.step_extra = np.empty(n_steps, dtype=np.int64)
            cigars = np.empty(max(cigar_bytes, 1), dtype=np.uint8)
            _cabi._check(lib.tvs_gaf_steps(h, p(names), p(step_fw), None, None, p(self.step_extra), p(cigars)))
            # distinct segment names -> ids, numbered in order of first appearance (a graph has far fewer segments than the
            # file has steps)
            n_seg, dbytes = ctypes.c_int64(), ctypes.c_int64()
            _cabi._check(lib.tvs_gaf_segments(h, ctypes.byref(n_seg), ctypes.byref(dbytes)))
            self.step_segment = np.empty(n_steps, dtype=np.int64)
            dptr = np.empty(n_seg.value + 1, dtype=np.int64)
            dsegs = np.empty(max(dbytes.value, 1), dtype=np.uint8)
            _cabi._check(lib.tvs_gaf_segment_ids(h, p(self.step_segment), p(dptr), p(dsegs)))
        finally:
            lib.tvs_gaf_destroy(h)
        self._names = names[:name_bytes].tobytes()
        self._cigars = cigars[:cigar_bytes].tobytes()
        self.step_forward = step_fw.astype(bool)
        raw, dp = dsegs[:dbytes.value].tobytes(), dptr.tolist()
        self.segment_names = [raw[dp[i]:dp[i + 1]].decode() for i in range(n_seg.value)]

    @classmethod
    def from_file(cls, path, alignment_format="GAF", **kw):
        if path.endswith("gz"):
            with gzip.open(path, "rb") as h:
                text = h.read()
        else:
            with open(path, "rb") as h:
                text = h.read()
        return cls(text, alignment_format=alignment_format, **kw)

    def __len__(self):
        return self.n_records

    def query_names(self):
        ptr = self._name_ptr.tolist()
        raw = self._names
        if raw.isascii():                       # one decode for the whole pool, then string slices
            txt = raw.decode("ascii")
            return [txt[a:b] for a, b in zip(ptr, ptr[1:])]
        return [raw[a:b].decode() for a, b in zip(ptr, ptr[1:])]

    def paths(self):
        """Per record the path as the reference's tuple of (segment name, forward) pairs; equal steps share one tuple."""
        n_seg = len(self.segment_names)
        pair = [(name, False) for name in self.segment_names] + [(name, True) for name in self.segment_names]
        code = (self.step_segment + n_seg * self.step_forward.astype(np.int64)).tolist()
        steps = list(map(pair.__getitem__, code))
        ptr = self.step_ptr.tolist()
        return [tuple(steps[a:b]) for a, b in zip(ptr, ptr[1:])]

    def records(self, parse_cigar=False):
        """The reference's ``raw_records``: a list of record objects in file order.  The cyclic garbage collector is paused
        while the objects are made (none of them can be garbage; with it running, its generation scans are 60 % of the time
        at 300k records)."""
        was_on = gc.isenabled()
        gc.disable()
        try:
            return self._records(parse_cigar)
        finally:
            if was_on:
                gc.enable()

    def _records(self, parse_cigar):
        n = self.n_records
        names, paths = self.query_names(), self.paths()
        if self.alignment_format == "GAF":
            text = self._text
            tag_off, tag_end = self._tag_off.tolist(), self._tag_end.tolist()
            if text.isascii():
                whole = text.decode("ascii")
                tags = [whole[a:b] if a >= 0 else "" for a, b in zip(tag_off, tag_end)]
            else:
                tags = [text[a:b].decode() if a >= 0 else "" for a, b in zip(tag_off, tag_end)]
            cols = [getattr(self, k).tolist() for k in ("query_len", "q_start", "q_end", "q_strand", "p_len", "p_start", "p_end",
                                                         "p_align_len", "num_match", "align_len", "align_quality", "identity")]
            out = []
            add = out.append
            for (name, path, tg, ql, qs, qe, st, pl, ps, pe, pal, nm, al, aq, idn) in zip(names, paths, tags, *cols):
                r = GAFRecord()
                r.query_name = name; r.path = path; r.query_len = ql; r.q_start = qs; r.q_end = qe; r.q_strand = st
                r.p_len = pl; r.p_start = ps; r.p_end = pe; r.p_align_len = pal; r.num_match = nm; r.align_len = al
                r.align_quality = aq; r.identity = idn; r._tags = tg; r._fields = None; r.cigar = None
                add(r)
            if parse_cigar:
                for r in out:
                    if "cg" in r.optional_fields:
                        r.cigar = r.split_cigar_str()
            return out
        ptr = self.step_ptr.tolist()
        extra = self.step_extra.tolist()
        cols = [getattr(self, k).tolist() for k in ("q_start", "q_end", "p_start", "query_len", "q_strand", "p_align_len", "p_end", "align_len")]
        out = []
        for i in range(n):
            r = SPATSVRecord()
            r.query_name, r.path = names[i], paths[i]
            r.q_start, r.q_end, r.p_start, r.query_len, r.q_strand, r.p_align_len, r.p_end, r.align_len = [c[i] for c in cols]
            r.q_align_len = abs(r.q_start - r.q_end) + 1
            r.path_align_lengths = extra[ptr[i]:ptr[i + 1]]
            out.append(r)
        return out


def parse_alignment_file(self, num_proc=1, _num_block_lines=10000):
    """Replacement for ``GraphAlignRecords.parse_alignment_file`` (GraphAlignRecords.py:1024-1134): fills
    ``self.raw_records`` and ``self.n_file_records`` from the native parser (``num_proc`` threads, 0 = all cores)."""
    from loguru import logger
    logger.info("Parsing alignment ({})".format(self.alignment_format))
    table = GafTable.from_file(self.alignment_file, alignment_format=self.alignment_format,
                               min_record_identity=getattr(self, "min_record_identity", 0.0),
                               parse_cigar=getattr(self, "parse_cigar", False), n_threads=int(num_proc))
    self.raw_records = table.records(parse_cigar=getattr(self, "parse_cigar", False))
    self.n_file_records = table.n_file_records
    self.gaf_table = table


def build_read_records(self):
    """Replacement for ``GraphAlignRecords.build_read_records`` (GraphAlignRecords.py:936-950): ``self.read_records`` =
    OrderedDict query name -> ReadRecord (the reference's class) holding the read's alignment records sorted by
    (q_start, q_end, cigar), its raw ids, the summed ``p_align_len`` and the length-weighted identity.

    The reference makes one ``ReadRecord.append`` call per alignment (three attribute updates and a division each,
    :225-230) and one ``sort_by`` per read (:252-256).  Here the raw ids are grouped by name in one pass and every
    ReadRecord is filled once; the sums run in raw order, so ``_p_identity_sum`` and ``p_identity`` are the same floats.
    ``filter_read_records`` (:952-1022) stays the reference's and calls this twice."""
    import sys
    from collections import OrderedDict
    from loguru import logger
    logger.debug("Build reads records ..")
    read_record_cls = sys.modules[type(self).__module__].ReadRecord
    raw = self.raw_records
    groups = OrderedDict()
    for go_r, record in enumerate(raw):
        ids = groups.get(record.query_name)
        if ids is None:
            groups[record.query_name] = [go_r]
        else:
            ids.append(go_r)
    new = read_record_cls.__new__
    was_on = gc.isenabled()
    gc.disable()                                   # as in GafTable.records: nothing made here can be garbage
    try:
        for name, ids in groups.items():
            rr = new(read_record_cls)
            rr.q_seq = None
            if len(ids) == 1:
                record = raw[ids[0]]
                rr.raw_ids = ids
                rr.records = [record]
                rr.p_align_len = 0 + record.p_align_len
                rr._p_identity_sum = 0. + record.p_align_len * record.identity
            else:
                records = [raw[i] for i in ids]
                p_align_len, identity_sum = 0, 0.
                for record in records:
                    p_align_len += record.p_align_len
                    identity_sum += record.p_align_len * record.identity
                rr.p_align_len = p_align_len
                rr._p_identity_sum = identity_sum
                order = sorted(range(len(ids)), key=lambda x: [records[x].q_start, records[x].q_end, records[x].cigar])
                rr.records = [records[x] for x in order]
                rr.raw_ids = [ids[x] for x in order]
            rr.p_identity = rr._p_identity_sum / rr.p_align_len
            groups[name] = rr
    finally:
        if was_on:
            gc.enable()
    self.read_records = groups


_SAVED = {}
_SWAPPED = ("parse_alignment_file", "build_read_records")


def install_on_records(cls=None):
    """Swap ``parse_alignment_file`` and ``build_read_records`` of the reference's GraphAlignRecords class (default:
    imported here)."""
    if cls is None:
        from traversome.GraphAlignRecords import GraphAlignRecords as cls
    if cls not in _SAVED:
        _SAVED[cls] = {name: cls.__dict__.get(name) for name in _SWAPPED}
    cls.parse_alignment_file = parse_alignment_file
    cls.build_read_records = build_read_records
    return cls


def uninstall_from_records():
    for cls, saved in list(_SAVED.items()):
        for name, fn in saved.items():
            if fn is not None:
                setattr(cls, name, fn)
            elif name in cls.__dict__:
                delattr(cls, name)
        del _SAVED[cls]
