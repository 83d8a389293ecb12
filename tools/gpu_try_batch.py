#!/usr/bin/env python
"""Bring-up aid for the read-batched kernel: golden cases through kernel=1, mismatches printed, not asserted.
Usage: python tools/gpu_try_batch.py [case ...]   (default: every golden case)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import dnab_testutil as util  # noqa: E402
import dnastore_b200 as d  # noqa: E402


def agree(workload, n, opts):
    """batch kernel vs push kernel on n synthetic reads of a bench workload (several groups per team)."""
    import bench
    w = bench.WORKLOADS[workload]
    compiled = util.machine_from_recipe(w["recipe"]).compile(d.ErrorFlags(length=w["length"], global_=True))
    reads = bench.make_reads(w, n, seed=4242)
    outs = []
    for kern in (1, 2):
        dec = d.Decoder(compiled, device=0)
        dec.set_option("kernel", kern)
        if kern == 1:
            for k, v in opts.items():
                dec.set_option(k, int(v))
        t0 = time.time()
        outs.append(dec.viterbi(reads, want_path=True))
        print(f"   kernel {kern}: {time.time() - t0:.2f}s", flush=True)
    a, b = outs
    bad = sum(1 for i in range(n) if a["loglike"][i].tobytes() != b["loglike"][i].tobytes() or a["decoded"][i] != b["decoded"][i]
              or a["path"][i].tolist() != b["path"][i].tolist() or a["status"][i] != b["status"][i])
    print(f"AGREE {workload} n={n} bad={bad}", flush=True)
    return bad


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--agree":
        opts = dict(kv.lstrip("-").split("=") for kv in sys.argv[4:] if kv.startswith("--"))
        return 1 if agree(sys.argv[2], int(sys.argv[3]), opts) else 0
    names = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = dict(kv.lstrip("-").split("=") for kv in sys.argv[1:] if kv.startswith("--"))
    cases = [c for c in util.load_golden() if not names or c["name"] in names]
    bad = 0
    for case in cases:
        compiled = util.compiled_for_case(case)
        dec = d.Decoder(compiled, device=0)
        dec.set_option("kernel", 1)
        for k, v in opts.items():
            dec.set_option(k, int(v))
        try:
            bi = dec.batch_info()
        except d.DnabError as e:
            print(case["name"], "SKIP:", e)
            continue
        reads = [r["seq"] for r in case["reads"]]
        t0 = time.time()
        dec.set_debug(True)
        out = dec.viterbi(reads, want_path=True)
        dt = time.time() - t0
        nbad = 0
        for i, r in enumerate(case["reads"]):
            ok_ll = util.hexf(out["loglike"][i]) == util.hexf(r["loglike_hex"])
            ok_dec = out["decoded"][i] == r["decoded"]
            ok_path = out["path"][i].tolist() == r["path"]
            if not (ok_ll and ok_dec and ok_path):
                nbad += 1
                if nbad <= 3:
                    print("   MISMATCH", case["name"], r["name"], "ll", ok_ll, out["loglike"][i], r["loglike"], "dec", ok_dec,
                          "path", ok_path, "status", out["status"][i], "len", len(r["seq"]))
                    if not ok_path:
                        a, b = out["path"][i].tolist(), r["path"]
                        for j in range(min(len(a), len(b))):
                            if a[j] != b[j]:
                                print("      first path difference at", j, a[j], b[j], "of", len(a), len(b))
                                break
        bad += nbad
        print(f"{case['name']:28s} states={compiled.t.n_states:6d} k={compiled.t.k} T={bi['team_size']:3d} M={bi['states_per_cta']:4d} "
              f"cross={bi['cross_cta_transition_fraction']:.2f} reads={len(reads):3d} bad={nbad} {dt:.2f}s dbg={dec.debug_counters()}", flush=True)
    print("TOTAL BAD", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
