#!/usr/bin/env python
"""Turns an `ncu --set full --import-source on` report into the small text summary committed here.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rNN_fill_summary.txt"""
import csv
import io
import re
import subprocess
import sys
import collections

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
print(f"# {rep}")
for h, u, v in zip(hdr, units, vals):
    if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
data = []
for r in rows[2:]:
    try:
        data.append((float(r[2]), float(r[5]), r[1].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
agg = collections.defaultdict(lambda: [0, 0])
for s, ie, sx in data:
    m = re.match(r"(@!?U?P\d\s+)?([A-Z0-9_.]+)", sx)
    op = m.group(2).split(".")[0] if m else sx[:10]
    agg[op][0] += s
    agg[op][1] += ie
print("\n# warp-stall samples and executed instructions by SASS opcode (share of kernel)")
for op, (s, ie) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print(f"{op:10s} samples {100 * s / tot:5.1f}%   instructions {100 * ie / toti:5.1f}%")
print("\n# top instructions by stall samples")
for s, ie, sx in sorted(data, key=lambda t: -t[0])[:12]:
    print(f"{100 * s / tot:5.1f}%  executed {int(ie):>11d}  {sx[:100]}")
print("\n# Blackwell evidence (SASS mnemonics present)")
ops = set(re.match(r"(@!?U?P\d\s+)?([A-Z0-9_.]+)", d[2]).group(2) for d in data if re.match(r"(@!?U?P\d\s+)?([A-Z0-9_.]+)", d[2]))
for key in ["UCGABAR_ARV", "UCGABAR_WAIT", "LD.E.64.STRONG.GPU", "ATOMS.OR", "ATOMS.EXCH", "ATOMS.ADD.S32", "MEMBAR.SC.CTA", "DADD", "DSETP.GEU.AND", "PRMT"]:
    print(f"{key}: {'yes' if any(o.startswith(key) for o in ops) else 'no'}")
