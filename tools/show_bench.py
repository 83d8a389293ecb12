#!/usr/bin/env python
"""Print the headline fields of bench logs: python tools/show_bench.py gpurun_out/bench_x_*.log"""
import json, sys
for f in sys.argv[1:]:
    for ln in open(f):
        if ln.startswith("{"):
            j = json.loads(ln)
            print(f"{f}: reads/s {j['reads_per_sec']:.0f} cells/s {j['value']:.3g} frac {j['roofline']['frac']:.4f} e2e {j['e2e']['reads_per_sec']:.0f}",
                  {k: j["config"].get(k) for k in ("team_size", "states_per_cta", "threads_per_cta", "reads_in_flight")})
        elif ln.startswith("rc="):
            print("   ", ln.strip())
