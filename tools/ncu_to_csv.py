#!/usr/bin/env python
"""Developer tool: `ncu -i X.ncu-rep --page raw --csv` -> the 3-column form committed under profiles/ (metric,unit,value),
one block per profiled launch.  Usage: tools/ncu_to_csv.py X.ncu-rep OUT.csv [launch index ...]"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
pick = [int(x) for x in sys.argv[3:]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]
with open(out, "w", newline="") as h:
    w = csv.writer(h)
    w.writerow(["metric", "unit", "value"])
    for i, r in enumerate(body):
        if pick and i not in pick:
            continue
        for name, unit, val in zip(hdr, units, r):
            if name in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
                continue
            w.writerow([name, unit, val])
print("wrote", out, len(body), "launch(es)")
