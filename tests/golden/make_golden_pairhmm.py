#!/usr/bin/env python
"""Regenerates tests/golden/pairhmm_golden.json from the UNMODIFIED reference
(oracle/_ref/refdriver fb|fit = FwdBackMatrix / baumWelchParams themselves, built by oracle/Makefile).

Cases: the reference's own forward-backward goldens (Makefile:156-163: data/dup.stk, dup.sub.stk,
dup.sub.misaligned.stk with -l6 and 1e-9 error probabilities; data/tiny.stk, test.stk fits with
--strict-guides) and seeded synthetic alignments (reference-encoded DNA mutated by the errdecode.pl
simulator, alignment kept), strict and non-strict envelopes.  Per alignment: forward and backward
log-likelihood as hex floats and every expected count with 17 significant digits; per fit: the fitted
parameters with 17 significant digits.  The Stockholm text itself is stored in the JSON.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from benchdata import synth  # noqa: E402

DATA = "/root/reference/data"
DRV = os.path.join(ROOT, "oracle", "_ref", "refdriver")
OUT = os.path.dirname(os.path.abspath(__file__))
DEFAULTS = dict(length=12, sub=.01, iv=10., dup=.001, delopen=.001, delext=.01)


def flag_args(f):
    return ["-l", str(f["length"]), "--sub", repr(f["sub"]), "--iv", repr(f["iv"]), "--dup", repr(f["dup"]),
            "--delopen", repr(f["delopen"]), "--delext", repr(f["delext"])]


def run(mode, stk_text, flags, strict):
    with tempfile.NamedTemporaryFile("w", suffix=".stk", delete=False) as tf:
        tf.write(stk_text)
        path = tf.name
    cmd = [DRV, mode] + flag_args(flags) + (["--strict"] if strict else []) + ["--stk", path]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    os.unlink(path)
    return out


def main():
    cases = []

    def add(name, stk_text, flags, strict, fit=False, note=""):
        f = dict(DEFAULTS)
        f.update(flags)
        aligns = []
        for ln in run("fb", stk_text, f, strict).strip().split("\n"):
            idx, fwd, back, counts = ln.split("\t")
            aligns.append(dict(fwd_hex=fwd, back_hex=back, counts=[float(x) for x in counts.split()],
                               counts_text=counts.split()))
        case = dict(name=name, stk=stk_text, flags=f, strict=strict, alignments=aligns, note=note)
        if fit:
            case["fit"] = run("fit", stk_text, f, strict).split()
        cases.append(case)
        print(" ", name, len(aligns), "alignments", "fit" if fit else "", flush=True)

    tiny = dict(length=6, sub=1e-9, dup=1e-9, delopen=1e-9)
    for nm in ["dup", "dup.sub", "dup.sub.misaligned"]:
        add(nm, open(f"{DATA}/{nm}.stk").read(), tiny, False, note=f"reference Makefile testcount on data/{nm}.stk")
    add("tiny_fit", open(f"{DATA}/tiny.stk").read(), {}, True, fit=True, note="reference Makefile testfit")
    add("test_fit", open(f"{DATA}/test.stk").read(), {}, True, fit=True, note="reference Makefile testfit")
    add("test_nonstrict", open(f"{DATA}/test.stk").read(), {}, False, fit=True)

    rng = np.random.default_rng(0xFB)
    pool = synth.load_pool("cfg1_l4c4_200b")
    rows = []
    for i in range(24):
        src = pool[i][: int(rng.integers(20, 120))]
        a, b = synth.mutate_aligned(src, rng, sub_rate=0.05, dup_rate=0.02, max_dup=3, del_rate=0.02, max_del=4)
        rows.append(f"# STOCKHOLM 1.0\nin  {a}\nout {b}\n//\n")
    rows.append("# STOCKHOLM 1.0\nin  ACGT\nout ----\n//\n")        # everything deleted
    rows.append("# STOCKHOLM 1.0\nin  AC--GT\nout ACACGT\n//\n")    # a clean tandem duplication
    stk = "".join(rows)
    add("synthetic_strict", stk, dict(sub=.05, dup=.02, delopen=.02, delext=.3), True, fit=True)
    add("synthetic_band", stk, dict(length=8, sub=.05, dup=.02, delopen=.02, delext=.3), False, fit=True)
    json.dump(dict(generator="tests/golden/make_golden_pairhmm.py", cases=cases),
              open(os.path.join(OUT, "pairhmm_golden.json"), "w"))
    print("done")


if __name__ == "__main__":
    main()
