#!/usr/bin/env python
"""Where a column's time goes in the read-batched kernel (debug build, rank 0 / warp 0 of the team):
   python tools/gpu_phase_probe.py cfg2 64 [--opt=value ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dnab_testutil as util  # noqa: E402
import dnastore_b200 as d  # noqa: E402
import bench  # noqa: E402


def main():
    wl, n = sys.argv[1], int(sys.argv[2])
    opts = dict(kv.lstrip("-").split("=") for kv in sys.argv[3:] if kv.startswith("--"))
    w = bench.WORKLOADS[wl]
    compiled = util.machine_from_recipe(w["recipe"]).compile(d.ErrorFlags(length=w["length"], global_=True))
    reads = bench.make_reads(w, n, seed=4242)
    dec = d.Decoder(compiled, device=0)
    dec.set_option("kernel", 1)
    for k, v in opts.items():
        dec.set_option(k, int(v))
    dec.viterbi(reads[:32])
    dec.set_debug(True)
    dec.viterbi(reads)
    c = list(dec.debug_counters().values())
    bi = dec.batch_info()
    groups = (n + 31) // 32
    cols = sum(max(len(r) for r in sorted(reads, key=len)[g * 32:(g + 1) * 32]) + 1 for g in range(groups))
    cyc = dict(init=c[8], barrier=c[9], closure=c[4], of_which_passive=c[7], record=c[5])
    tot = c[8] + c[9] + c[4] + c[5]
    print(f"{wl} T={bi['team_size']} M={bi['states_per_cta']} groups={groups} columns~{cols} total {tot / max(cols, 1):.0f} cycles/column")
    for k, v in cyc.items():
        print(f"   {k:18s} {v / max(cols, 1):9.0f} cycles/column  {100.0 * v / tot:5.1f} %")
    nw = bi.get('warps_per_cta', 32)
    print(f"   per warp and column (rank 0, {nw} warps): in relax {c[10] / nw / max(cols, 1):.0f} cycles, in notification flush {c[11] / nw / max(cols, 1):.0f} cycles, of closure {c[4] / max(cols, 1):.0f}")
    print(f"   visits/state/column {c[2] / max(cols, 1) / bi['states_per_cta']:.2f} (rank 0)  wakes/column {c[6] / max(cols, 1):.2f}  edges/visit {c[3] / max(c[2], 1):.2f}")


if __name__ == "__main__":
    main()
