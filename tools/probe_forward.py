"""Scratch probe (GPU): forward kernel throughput."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench, dnab_testutil as util, dnastore_b200 as d
wl = sys.argv[1]; n = int(sys.argv[2]); cut = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mode = sys.argv[4] if len(sys.argv) > 4 else "forward"
w = bench.WORKLOADS[wl]
m = util.machine_from_recipe(w['recipe']); c = m.compile(d.ErrorFlags(length=w['length'], global_=True))
dec = d.Decoder(c)
reads = bench.make_reads(w, n, 7)
if cut: reads = [r[:cut] for r in reads]
run = dec.forward if mode == "forward" else dec.fwdback_counts
run(reads[:2])
t0 = time.perf_counter(); out = run(reads); dt = time.perf_counter() - t0
if mode != "forward":
    out["sweeps"] = np.zeros(len(reads))
    print("mean counts", np.round(out["counts"].mean(axis=0)[:5 + c.t.k], 4).tolist(), "max |ll_back - ll|", float(np.abs(out["loglike_back"] - out["loglike"]).max()))
st = dec.stats()
cols = sum(len(r) + 1 for r in reads)
print(json.dumps(dict(mode=mode, workload=wl, reads=n, kernel_ms=st['last_fill_ms'], wall_s=dt, reads_per_s=n / (st['last_fill_ms'] * 1e-3),
      sweeps_per_col=float(out['sweeps'].sum()) / cols, us_per_sweep_per_cta=st['last_fill_ms'] * 1e3 * min(n, 148) / max(float(out['sweeps'].sum()), 1),
      cells_per_s=c.t.n_states * cols * (c.t.k + 2) / (st['last_fill_ms'] * 1e-3), status_ok=bool((out['status'] == 0).all()))))
