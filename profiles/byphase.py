#!/usr/bin/env python
"""Phase view of an ncu report of viterbiFillPushKernel captured with --import-source on: executed warp
instructions, active lanes and warp-stall samples per phase of the column loop (emission step, first closure
pass, level scan, push, level-end barrier, predecessor pass), found through the comment markers in the source.
usage: python profiles/byphase.py gpurun_out/prof.ncu-rep"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, cur, lines = None, "", []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif len(r) > 8 and r[0].isdigit():
        lines.append((cur, int(r[0]), r))
ix = {n: i for i, n in enumerate(hdr)}
src = open(os.path.join(ROOT, "dnastore_b200", "csrc", "viterbi_fill_push.cu")).read().split("\n")


def find(marker):
    for i, l in enumerate(src):
        if marker in l:
            return i + 1


marks = [("helpers", 1), ("pushState", find("uint32_t pushState(")), ("kernel set-up", find("viterbiFillPushKernel(const __grid_constant__")),
         ("(1) S0 copy into shared memory", find("// ---- (1) S0(pos) into shared memory")), ("(2a) first closure pass", find("// ---- (2a) closure, first pass")),
         ("(2b) level scan + queue barrier", find("// ---- (2b) closure, PUSH levels")), ("(2b) push loop", find("// one hop per level, breadth first")),
         ("(2b) level-end barrier / cluster meeting", find("// every push of this level has flagged or queued")),
         ("(3) predecessor pass + emission step of the next column", find("// ---- (3) predecessor records")),
         ("end of read", find("// ---- end of read"))]


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def phase(f, n):
    if f != "viterbi_fill_push.cu":
        return "inlined from " + f
    p = "helpers"
    for name, ln in marks:
        if ln and n >= ln:
            p = name
    return p


stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
agg = {}
for f, n, r in lines:
    a = agg.setdefault(phase(f, n), dict(inst=0, thr=0, samp=0, **{c: 0 for c in stall_cols}))
    a["inst"] += num(r[ix["Instructions Executed"]])
    a["thr"] += num(r[ix["Thread Instructions Executed"]])
    a["samp"] += num(r[ix["# Samples"]])
    for c in stall_cols:
        a[c] += num(r[ix[c]])
ti = sum(a["inst"] for a in agg.values())
ts = sum(a["samp"] for a in agg.values())
print(f"# {rep}: by phase (pushState is inlined into the push loop; helpers = the small device functions)")
print(f"# total warp instructions {ti:.3g}, stall samples {ts:.0f}")
for p, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"]):
    top = sorted(((c, a[c]) for c in stall_cols), key=lambda kv: -kv[1])[:4]
    print("%-42s instr %5.1f%%  lanes %4.1f  samples %5.1f%%   %s" % (
        p, 100 * a["inst"] / ti, a["thr"] / max(a["inst"], 1), 100 * a["samp"] / ts,
        ", ".join("%s %.0f%%" % (c[6:], 100 * v / max(a["samp"], 1)) for c, v in top)))
