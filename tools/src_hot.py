#!/usr/bin/env python
"""Developer tool: per CUDA-source-line instruction / stall shares from
`ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Line No"][0]
hdr = rows[hi]
ie = hdr.index("Instructions Executed")
ws = hdr.index("Warp Stall Sampling (All Samples)")


def num(x):
    return int(x) if x.isdigit() else 0


body = [r for r in rows[hi + 1:] if len(r) > ie and r[0] != ""]
tot = sum(num(r[ie]) for r in body)
tots = sum(num(r[ws]) for r in body)
print("warp instructions", tot, "stall samples", tots)
for r in body:
    if num(r[ie]) > tot * thr or num(r[ws]) > tots * thr:
        print("%5s %-100s inst %5.1f%% stall %5.1f%%" % (r[0], r[1].strip()[:100], 100.0 * num(r[ie]) / tot, 100.0 * num(r[ws]) / tots))
