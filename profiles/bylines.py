#!/usr/bin/env python
"""Per-source-line view of an ncu report captured with --import-source on (-lineinfo build):
stall samples and executed warp instructions aggregated by CUDA source line.
usage: python profiles/bylines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].isdigit():
        try:
            lines.append((float(r[6] or 0), float(r[7] or 0), float(r[10] or 0), cur_file, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print(f"# {rep}: by source line (samples%, warp-instr%, avg active threads)")
for s, ie, thr, f, n, src in sorted(lines, key=lambda t: -t[0])[:top]:
    print(f"{100*s/tot:5.1f}% {100*ie/toti:5.1f}% {thr:5.1f}  {f}:{n}  {src[:110]}")
