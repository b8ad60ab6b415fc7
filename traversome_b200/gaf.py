"""Alignment-file parsing in array form (the part BEFORE read-path construction).

Mirror of what ``GraphAlignRecords.parse_alignment_file`` leaves behind (traversome/GraphAlignRecords.py:1024-1134):
``raw_records`` -- one record object per kept alignment line -- and ``n_file_records``.  The reference builds every record
with a Python constructor (``GAFRecord.__init__`` :33-69, ``SPATSVRecord.__init__`` :109-131, ``gaf_str_to_path``
utils.py:1065-1082), in a process pool when ``num_proc > 1`` (the objects are pickled back).  Here the text of the file goes
to the native parser of ``libtvs.so`` once (``include/tvs_gaf.h``: threads over blocks of lines, integer columns and string
pools) and the record objects are made from the columns; the arrays themselves (``GafTable``) are what
``traversome_b200.readpaths.generate_read_paths`` needs.

``install_on_records(GraphAlignRecords)`` swaps ``parse_alignment_file`` of the reference's class; the per-read filters that
follow (``build_read_records``, ``filter_read_records``, :920-1022) are the reference's and see the same objects.
"""
import ctypes
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
            seg_ptr = np.empty(n_steps + 1, dtype=np.int64)
            segs = np.empty(max(seg_bytes, 1), dtype=np.uint8)
            self.step_extra = np.empty(n_steps, dtype=np.int64)
            cigars = np.empty(max(cigar_bytes, 1), dtype=np.uint8)
            _cabi._check(lib.tvs_gaf_steps(h, p(names), p(step_fw), p(seg_ptr), p(segs), p(self.step_extra), p(cigars)))
        finally:
            lib.tvs_gaf_destroy(h)
        self._names = names[:name_bytes].tobytes()
        self._cigars = cigars[:cigar_bytes].tobytes()
        self.step_forward = step_fw.astype(bool)
        # distinct segment names -> ids (a graph has far fewer segments than the file has steps)
        seg_bytes_all = segs[:seg_bytes].tobytes()
        ptr = seg_ptr.tolist()
        ids, names_list, step_segment = {}, [], np.empty(n_steps, dtype=np.int64)
        for i in range(n_steps):
            key = seg_bytes_all[ptr[i]:ptr[i + 1]]
            j = ids.get(key)
            if j is None:
                j = ids[key] = len(names_list)
                names_list.append(key.decode())
            step_segment[i] = j
        self.segment_names = names_list
        self.step_segment = step_segment

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
        return [raw[ptr[i]:ptr[i + 1]].decode() for i in range(self.n_records)]

    def paths(self):
        """Per record the path as the reference's tuple of (segment name, forward) pairs; equal steps share one tuple."""
        names, seg, fw = self.segment_names, self.step_segment.tolist(), self.step_forward.tolist()
        pair = {}
        steps = []
        for s, f in zip(seg, fw):
            k = (s, f)
            t = pair.get(k)
            if t is None:
                t = pair[k] = (names[s], f)
            steps.append(t)
        ptr = self.step_ptr.tolist()
        return [tuple(steps[ptr[i]:ptr[i + 1]]) for i in range(self.n_records)]

    def records(self, parse_cigar=False):
        """The reference's ``raw_records``: a list of record objects in file order."""
        n = self.n_records
        names, paths = self.query_names(), self.paths()
        if self.alignment_format == "GAF":
            text = self._text
            tag_off, tag_end = self._tag_off.tolist(), self._tag_end.tolist()
            cols = [getattr(self, k).tolist() for k in ("query_len", "q_start", "q_end", "q_strand", "p_len", "p_start", "p_end",
                                                         "p_align_len", "num_match", "align_len", "align_quality", "identity")]
            out = []
            for i in range(n):
                r = GAFRecord()
                r.query_name, r.path = names[i], paths[i]
                (r.query_len, r.q_start, r.q_end, r.q_strand, r.p_len, r.p_start, r.p_end, r.p_align_len, r.num_match, r.align_len,
                 r.align_quality, r.identity) = [c[i] for c in cols]
                r._tags = text[tag_off[i]:tag_end[i]].decode() if tag_off[i] >= 0 else ""
                r._fields = None
                r.cigar = r.split_cigar_str() if (parse_cigar and "cg" in r.optional_fields) else None
                out.append(r)
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


_SAVED = {}


def install_on_records(cls=None):
    """Swap ``parse_alignment_file`` of the reference's GraphAlignRecords class (default: imported here)."""
    if cls is None:
        from traversome.GraphAlignRecords import GraphAlignRecords as cls
    if cls not in _SAVED:
        _SAVED[cls] = cls.__dict__.get("parse_alignment_file")
    cls.parse_alignment_file = parse_alignment_file
    return cls


def uninstall_from_records():
    for cls, fn in list(_SAVED.items()):
        if fn is not None:
            cls.parse_alignment_file = fn
        elif "parse_alignment_file" in cls.__dict__:
            del cls.parse_alignment_file
        del _SAVED[cls]
