#!/usr/bin/env python
"""Developer tool (host only, no GPU): throughput of the alignment-file parser (include/tvs_gaf.h) against the reference's
per-line Python constructors (restated in oracle/gaf_port.py; the reference's own _gaf_parse_worker when /root/reference
is present).  Usage: tools/gaf_bench.py [megabytes]"""
import gzip, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gaf_port  # noqa: E402
from traversome_b200 import gaf  # noqa: E402

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
with gzip.open(os.path.join(ROOT, "tests", "golden", "plastome_sim.gaf.gz"), "rb") as h:
    base = h.read().decode().splitlines(True)
lines, size, rep = [], 0, 0
while size < mb * 1e6:
    for l in base:
        f = l.split("\t")
        f[0] = "%s/%d" % (f[0], rep)
        f[12] = "NM:i:%d\tid:f:%.6f\tcg:Z:%sM\n" % (rep % 97, 1.0 - (rep % 97) * 1e-4, f[10])
        s = "\t".join(f)
        lines.append(s)
        size += len(s)
    rep += 1
text = "".join(lines).encode()
print("text: %.1f MB, %d lines" % (len(text) / 1e6, len(lines)))
for nt in (1, 4, 0):
    t = time.perf_counter()
    tab = gaf.GafTable(text, n_threads=nt)
    dt = time.perf_counter() - t
    print("native parser, %s thread(s): %.3f s = %.0f MB/s, %.2f M records/s (%d records, %d steps)" % (
        nt or "all", dt, len(text) / 1e6 / dt, tab.n_records / 1e6 / dt, tab.n_records, tab.step_forward.size))
t = time.perf_counter()
recs = tab.records()
dt = time.perf_counter() - t
print("record objects from the columns: %.3f s = %.2f M records/s" % (dt, len(recs) / 1e6 / dt))
sample = lines[:max(1, len(lines) // 20)]
sb = sum(len(s) for s in sample)
t = time.perf_counter()
gaf_port.parse_gaf(sample)
dt = time.perf_counter() - t
print("restated reference constructors (1 process, %.1f MB sample): %.0f MB/s, %.3f M records/s" % (sb / 1e6, sb / 1e6 / dt, len(sample) / 1e6 / dt))
if os.path.isdir("/root/reference/traversome"):
    sys.path.insert(0, "/root/reference")
    from traversome import GraphAlignRecords as G
    t = time.perf_counter()
    G._gaf_parse_worker(sample, 0.0, False)
    dt = time.perf_counter() - t
    print("reference _gaf_parse_worker (1 process, same sample): %.0f MB/s, %.3f M records/s" % (sb / 1e6 / dt, len(sample) / 1e6 / dt))
