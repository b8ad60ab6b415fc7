"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's alignment-line parsing.

Follows, line by line of the file:
  GAFRecord.__init__            traversome/GraphAlignRecords.py:33-69   (12 columns, optional tags, identity)
  gaf_str_to_path               traversome/utils.py:1065-1082
  _gaf_parse_worker             traversome/GraphAlignRecords.py:151-171 (strip, split, identity filter, line count)
  SPATSVRecord.__init__         traversome/GraphAlignRecords.py:109-131, parse_spa_tsv_path :140-144
  _tsv_parse_worker             traversome/GraphAlignRecords.py:178-196
Pinned against the unmodified reference by oracle/make_golden_gaf.py -> tests/golden/gaf_records.json.gz
(tests/test_gaf_parser.py checks this port AND the native parser against those vectors).  Only tests/ may import it.
"""
import re

GAF_FIELDS = ("query_name", "query_len", "q_start", "q_end", "q_strand", "path", "p_len", "p_start", "p_end", "p_align_len",
              "num_match", "align_len", "align_quality", "identity", "optional_fields", "cigar")
SPA_FIELDS = ("query_name", "q_start", "q_end", "p_start", "query_len", "q_align_len", "path", "path_align_lengths", "q_strand",
              "p_align_len", "p_end", "align_len")
_STRAND = {"+": True, "-": False}


def gaf_str_to_path(path_str):
    path = []
    for segment in re.findall(r".[^\s><]*", path_str):
        if segment[0] == ">":
            path.append((segment[1:].split(":")[0], True))
        elif segment[0] == "<":
            path.append((segment[1:].split(":")[0], False))
        else:
            path.append((segment.split(":")[0], True))
    return tuple(path)


def gaf_record(split, parse_cigar=False):
    r = {"query_name": split[0], "query_len": int(split[1]), "q_start": int(split[2]), "q_end": int(split[3]),
         "q_strand": _STRAND[split[4]], "path": gaf_str_to_path(split[5]), "p_len": int(split[6]), "p_start": int(split[7]),
         "p_end": int(split[8])}
    r["p_align_len"] = r["p_end"] - r["p_start"]
    r["num_match"], r["align_len"], r["align_quality"] = int(split[9]), int(split[10]), int(split[11])
    fields = {}
    for tag in split[12:]:
        flag, ty, val = tag.split(":", maxsplit=2)
        if ty == "i":
            fields[flag] = int(val)
        elif ty == "Z":
            fields[flag] = val
        elif ty == "f":
            fields[flag] = float(val)
    r["optional_fields"] = fields
    r["cigar"] = None
    if parse_cigar and "cg" in fields:
        parts = re.split("([MIDNSHPX=])", fields["cg"])[:-1]
        r["cigar"] = [(int(parts[i]), parts[i + 1]) for i in range(0, len(parts), 2)]
    r["identity"] = fields.get("id", r["num_match"] / float(r["align_len"]))
    return r


def parse_gaf(lines, min_record_identity=0.0, parse_cigar=False):
    """-> (records kept, with their 0-based line number under "line_no"; number of lines read)"""
    out, n = [], 0
    for line in lines:
        split = line.strip().split("\t")
        n += 1
        rec = gaf_record(split, parse_cigar)
        if rec["identity"] >= min_record_identity:
            rec["line_no"] = n - 1
            out.append(rec)
    return out, n


def spa_record(split):
    r = {"query_name": split[0], "q_start": int(split[1]), "q_end": int(split[2]), "p_start": int(split[3]),
         "query_len": int(split[5])}
    r["q_align_len"] = abs(r["q_start"] - r["q_end"]) + 1
    r["path"] = tuple((v[:-1], _STRAND[v[-1]]) for v in split[6].split(","))
    r["path_align_lengths"] = [int(x) for x in split[7].split(",")]
    r["q_strand"] = r["q_end"] >= r["q_start"]
    r["p_align_len"] = sum(r["path_align_lengths"])
    r["p_end"] = r["p_start"] + r["p_align_len"]
    r["align_len"] = max(r["p_align_len"], r["q_align_len"])
    return r


def parse_spa(lines):
    out, n = [], 0
    for line in lines:
        split = line.strip().split("\t")
        n += 1
        if "," in split[1]:
            continue
        rec = spa_record(split)
        rec["line_no"] = n - 1
        out.append(rec)
    return out, n
