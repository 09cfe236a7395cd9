#!/usr/bin/env python3
"""`ncu -i X.ncu-rep --page raw --csv` (one wide row per launch) -> metric,unit,value rows of the LAST launch in the report."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(txt.splitlines()) if r]
h, u, v = rows[0], rows[1], rows[-1]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for n, un, val in zip(h, u, v):
        w.writerow([n, un, val])
print(out, len(h), "metrics; kernel:", dict(zip(h, v)).get("Kernel Name"))
