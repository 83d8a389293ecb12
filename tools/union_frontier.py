#!/usr/bin/env python
"""Union-frontier inflation of the read-batched closure.

For every BASELINE workload: decode groups of R = 1, 8, 16, 32 synthetic reads on the CPU with the
breadth-first closure of tools/union_frontier.c and report, per state-column,
  visits/read   per-read pushes (what a one-read-per-cluster kernel does)
  union visits  pushes of the group when a state is relaxed for all R lanes as soon as any read raised it
  per lane      union visits / R  = warp-level work per read when the R reads are the SIMD lanes
  levels        breadth-first levels per column: mean per read, and of the group (= the barriers a batched kernel pays)

Usage: python tools/union_frontier.py [--workloads cfg1,cfg4,...] [--groups G] [--out profiles/r02_union_frontier.json]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def lib():
    out = os.path.join(ROOT, "tools", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libunion_frontier.so")
    src = os.path.join(ROOT, "tools", "union_frontier.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src])
    l = C.CDLL(so)
    l.union_frontier.restype = C.c_int
    l.union_frontier.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return l


def main():
    import bench
    import dnab_testutil as util
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="cfg1,cfg4,cfg5,cfg3,cfg2")
    ap.add_argument("--sizes", default="1,8,16,32")
    ap.add_argument("--groups", type=int, default=2)
    ap.add_argument("--maxlen", type=int, default=0, help="truncate reads (0 = full length)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    l = lib()
    rows = []
    for name in a.workloads.split(","):
        w = bench.WORKLOADS[name]
        compiled = util.compiled_for(w["recipe"], dict(length=w["length"]), True)
        t = compiled.t
        for R in [int(x) for x in a.sizes.split(",")]:
            groups = a.groups if t.n_states < 20000 else 1
            acc = np.zeros(8)
            for g in range(groups):
                reads = bench.make_reads(w, R, seed=1000 + g)
                if a.maxlen:
                    reads = [r[:a.maxlen] for r in reads]
                toks = np.concatenate([util.tokens(r) for r in reads])
                off = np.zeros(R + 1, dtype=np.int64)
                off[1:] = np.cumsum([len(r) for r in reads])
                out = np.zeros(8)
                l.union_frontier(C.addressof(t), toks.ctypes.data, off.ctypes.data, R, out.ctypes.data)
                acc += out
            # out[5] = n * columns of the longest read; per-read state-columns = sum over reads
            row = dict(workload=name, n_states=int(t.n_states), R=R, groups=groups,
                       visits_per_read=acc[0] / (acc[5] * R) if True else 0,
                       union_visits=acc[1] / acc[5], union_per_lane=acc[1] / acc[5] / R,
                       union_visits_after_level0=acc[6] / acc[5],
                       levels_group=acc[2] / acc[4], levels_per_read=acc[3] / (acc[4] * R))
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        json.dump(dict(tool="tools/union_frontier.py", rows=rows), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
