import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, 'tests')
import numpy as np, dnab_testutil as util, dnastore_b200 as d
case = util.golden_case(sys.argv[1]); rname = sys.argv[2]
cfg = dict(cluster_size=int(sys.argv[3]), t_in_smem_mode=int(sys.argv[4]), table_mode=int(sys.argv[5]), partition_mode=int(sys.argv[6]))
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 10
r = [x for x in case["reads"] if x["name"] == rname][0]
compiled = util.compiled_for_case(case)
o = util.oracle_viterbi(compiled, r["seq"], want_cells=True)
bad_runs = 0
for rep in range(reps):
    dec = d.Decoder(compiled); dec.configure(**cfg)
    ll, cells = dec.viterbi_cells(r["seq"])
    diff = np.argwhere(cells.view(np.uint64) != o["cells"].view(np.uint64))
    out = dec.viterbi([r["seq"]], want_path=True)
    pathok = out["path"][0].tolist() == r["path"]
    if len(diff) or not pathok:
        bad_runs += 1
        print("rep", rep, "cells differing:", len(diff), "path ok:", pathok, "first:", [(tuple(x), cells[tuple(x)], o["cells"][tuple(x)]) for x in diff[:5]])
print(cfg, "bad runs", bad_runs, "of", reps)
