#!/usr/bin/env python
"""Developer tool: opcode mix and hot instructions from `ncu -i X.ncu-rep --page source --csv`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hi = [i for i, r in enumerate(rows) if "Source" in r][0]
hdr = rows[hi]
si = hdr.index("Source")
ie = hdr.index("Instructions Executed")
ws = hdr.index("Warp Stall Sampling (All Samples)")
body = [r for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body)
tots = sum(int(r[ws]) for r in body)
print("warp instructions", tot, "stall samples", tots, "sass lines", len(body))
mix = collections.Counter()
for r in body:
    op = [o for o in r[si].split() if not o.startswith("@")][0].split(".")[0]
    mix[op] += int(r[ie])
for k, v in mix.most_common(24):
    print("  %-8s %10d %5.1f%%" % (k, v, 100.0 * v / tot))
for i, r in enumerate(body):
    if int(r[ie]) > tot * thr or int(r[ws]) > tots * thr:
        print("%4d %-72s %9s %6s" % (i, r[si].strip()[:72], r[ie], r[ws]))
