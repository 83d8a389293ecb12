#!/usr/bin/env python
"""Round-2 golden set: the BENCHMARK read distributions at full length, pinned to the UNMODIFIED reference.

    python tests/golden/make_golden_r2.py        (build container only: needs oracle/_ref/refdriver)

Writes tests/golden/viterbi_golden_r2.json.gz with, per read, the reference's decoded string, fp64
log-likelihood (hex float) and traceback path (the same method as make_golden.py: oracle/_ref/refdriver
linked against the reference's own objects):

  cfg2_bench   64 reads   flusher*mixradar6*dnastore-l4 (46,670 states), bench.make_reads(cfg2, seed 20262)
  cfg3_bench   32 reads   sync16*hamming74*dnastore-l4, indels
  cfg4_bench   32 reads   watermark64.1*dnastore-l4
  cfg5_bench   32 reads   dnastore-l8 (k = 4)
  cfg1_bench   64 reads   dnastore-l4
  l10_full      8 reads   dnastore-l10 (57,090 states, k = 5), 160-bit payloads encoded by the reference, 1 % substitutions
and the sha256 of EVERY DP cell (reference layout, little-endian fp64) of a short read on the 46,670-state
machine, global and local mode -- cell-exactness on a machine that needs a multi-CTA cluster / team.
The reads are the ones bench.py decodes (same pools, same mutation code), so the headline workload is
pinned by the reference itself, not by agreement between two of this repository's kernels.
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import bench  # noqa: E402
from benchdata import synth  # noqa: E402

RECIPES = {
    "cfg1": ["l4c4"], "cfg2": ["l4c4", "flusher", "mixradar6"], "cfg3": ["l4c4", "sync16", "flusher", "hamming74"],
    "cfg4": ["l4c4", "water64.1"],
}


def _machine_file(name):
    tf = tempfile.NamedTemporaryFile("wb", suffix=".json", delete=False)
    tf.write(gzip.open(os.path.join(HERE, "machines", name + ".json.gz"), "rb").read())
    tf.close()
    return tf.name


def _run_chunk(job):
    kind, recipe, flags, glob, reads, machine_file = job
    if machine_file:
        mg.machine_args = lambda _r: ["--machine", machine_file]
    return mg.ref_viterbi(recipe, flags, glob, reads)


def ref_parallel(recipe, flags, glob, reads, machine_file=None, workers=8, chunk=4):
    jobs = [("v", recipe, flags, glob, reads[i:i + chunk], machine_file) for i in range(0, len(reads), chunk)]
    with ProcessPoolExecutor(workers) as ex:
        parts = list(ex.map(_run_chunk, jobs))
    res = [r for p in parts for r in p]
    for r, (_n, seq) in zip(res, reads):
        r["seq"] = seq
    return res


def main():
    cases = []
    for wl, n in (("cfg1", 64), ("cfg2", 64), ("cfg3", 32), ("cfg4", 32), ("cfg5", 32)):
        w = bench.WORKLOADS[wl]
        reads = [(f"r{i}", s) for i, s in enumerate(bench.make_reads(w, n, seed=20262))]
        flags = dict(length=w["length"])
        mfile = _machine_file("l8c4") if wl == "cfg5" else None
        recipe = RECIPES.get(wl, ["l8c4"])
        res = ref_parallel(recipe, flags, True, reads, mfile)
        _a, f = mg.flag_args(flags, True)
        cases.append(dict(name=f"{wl}_bench", recipe=list(w["recipe"]), flags=f, global_=True,
                          note=f"bench.make_reads({wl}, {n}, seed=20262): {w['desc']}", reads=res))
        print(wl, len(res), "reads", flush=True)
    # dnastore -l 10 at full length
    m10 = _machine_file("l10c4")
    mg.machine_args = lambda _r: ["--machine", m10]
    rng = np.random.default_rng(0xD5A57012 + 210)
    enc = mg.ref_encode(["l10c4"], [synth.random_bits(rng, 160) for _ in range(8)])
    reads10 = [(f"r{i}", synth.mutate(e, rng, sub_rate=0.01)) for i, e in enumerate(enc)]
    res = ref_parallel(["l10c4"], dict(length=10), True, reads10, m10, chunk=1)
    _a, f = mg.flag_args(dict(length=10), True)
    cases.append(dict(name="l10_full", recipe=["l10c4"], flags=f, global_=True,
                      note="dnastore -l 10 (57,090 states, k = 5), 160-bit payloads, 1 % substitutions", reads=res))
    print("l10", len(res), "reads", [len(s) for _n, s in reads10], flush=True)
    # every DP cell of a short read on the 46,670-state machine, as a hash
    mg.machine_args = lambda recipe: sum([["--compose", os.path.join(mg.DATA, c + ".json")] if i else
                                          ["--machine", os.path.join(mg.DATA, c + ".json")] for i, c in enumerate(recipe)], [])
    cells = []
    seq = cases[1]["reads"][0]["seq"][:14]
    for glob in (True, False):
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as tf:
            tmp = tf.name
        r = mg.ref_viterbi(RECIPES["cfg2"], dict(length=4), glob, [("x", seq)], cells_file=tmp)[0]
        raw = np.fromfile(tmp, dtype=np.float64)
        os.unlink(tmp)
        _a, f = mg.flag_args(dict(length=4), glob)
        cells.append(dict(name="cfg2_cells_" + ("global" if glob else "local"), recipe=RECIPES["cfg2"], flags=f, global_=glob,
                          seq=seq, n_cells=int(raw.size), sha256=hashlib.sha256(raw.tobytes()).hexdigest(),
                          finite=int(np.isfinite(raw).sum()), loglike_hex=r["loglike_hex"]))
        print("cells", cells[-1]["name"], raw.size, cells[-1]["finite"], flush=True)
    out = os.path.join(HERE, "viterbi_golden_r2.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as fz:
        fz.write(json.dumps(dict(generator="tests/golden/make_golden_r2.py", cases=cases, cells=cells)).encode())
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
