#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total ms, share.

    python tools/summarize_launches.py gpurun_out/launches.csv --steps 2 > profiles/rNN_launch_summary.json
"""
import argparse
import collections
import csv
import json
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--steps", type=int, default=2, help="timed fine-tune steps inside the capture")
    a = ap.parse_args()
    rows = [r for r in csv.reader(open(a.csv, errors="replace")) if r and not r[0].startswith("==")]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols = rows[hdr]
    ik, iv, iu = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*$", "", r[ik]).strip()
        v = float(r[iv].replace(",", ""))
        ms = v / 1e3 if r[iu] in ("usecond", "us") else (v / 1e6 if r[iu] in ("nsecond", "ns") else v)
        agg[name][0] += 1
        agg[name][1] += ms
    total = sum(v[1] for v in agg.values())
    out = {"steps": a.steps, "total_ms": total, "ms_per_step_serialised": total / a.steps,
           "kernels": {k: {"launches": v[0], "ms": v[1], "share": v[1] / total}
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
