#!/usr/bin/env python
"""Executed warp instructions and stall samples of an ncu report (--import-source on) aggregated over source-line
ranges of viterbi_fill_batch.cu given as name:lo-hi.  usage: python profiles/byrange.py rep name:lo-hi ..."""
import csv, io, subprocess, sys
rep = sys.argv[1]
ranges = []
for a in sys.argv[2:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""
agg = {n: [0, 0] for n, _, _ in ranges}
agg["other"] = [0, 0]
tot = [0, 0]
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].isdigit():
        try:
            s, ie = float(r[6] or 0), float(r[7] or 0)
        except ValueError:
            continue
        line = int(r[0])
        key = "other"
        if cur_file == "viterbi_fill_batch.cu":
            for n, lo, hi in ranges:
                if lo <= line <= hi:
                    key = n
                    break
        agg[key][0] += s
        agg[key][1] += ie
        tot[0] += s
        tot[1] += ie
print(f"# {rep}: total warp instructions {tot[1]:.4g}, stall samples {tot[0]:.4g}")
for k, (s, ie) in agg.items():
    print(f"{k:14s} instructions {100 * ie / tot[1]:5.1f}%  samples {100 * s / tot[0]:5.1f}%")
