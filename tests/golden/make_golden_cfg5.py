#!/usr/bin/env python
"""Adds the golden case of BASELINE config 5 (the -l 8 machine, k = 4) -> viterbi_golden_cfg5.json.

Same method as make_golden.py (the UNMODIFIED reference through oracle/_ref/refdriver: decoded string,
fp64 log-likelihood, traceback path per read); kept separate so that the seeded cases of
viterbi_golden.json stay byte-identical.  The machine is tests/golden/machines/l8c4.json.gz, emitted
once by the reference builder in the build container (`oracle/_ref/dnastore -v0 -l 8 --save-machine`).
"""
import gzip
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from benchdata import synth  # noqa: E402


def main():
    tf = tempfile.NamedTemporaryFile("wb", suffix=".json", delete=False)
    tf.write(gzip.open(os.path.join(HERE, "machines", "l8c4.json.gz"), "rb").read())
    tf.close()
    mg.machine_args = lambda recipe: ["--machine", tf.name]  # one machine, no composition
    rng = np.random.default_rng(0xD5A57012 + 5)
    payloads = [synth.random_bits(rng, 150) for _ in range(3)]
    enc = mg.ref_encode(["l8c4"], payloads)
    reads = [(f"r{i}", synth.mutate(e, rng, sub_rate=0.01)) for i, e in enumerate(enc)]
    reads.append(("indels", synth.mutate(enc[0], rng, sub_rate=0.01, dup_rate=0.01, max_dup=4, del_rate=0.01, max_del=4)))
    cases = []
    for name, glob in (("cfg5_l8_global", True), ("cfg5_l8_local", False)):
        mine = reads if glob else reads[:2] + reads[3:]
        res = mg.ref_viterbi(["l8c4"], dict(length=8), glob, mine)
        for r, (_n, seq) in zip(res, mine):
            r["seq"] = seq
        _a, f = mg.flag_args(dict(length=8), glob)
        cases.append(dict(name=name, recipe=["l8c4"], flags=f, global_=bool(glob),
                          note="BASELINE config 5 machine: dnastore -l 8, 10,746 states, k = 4", reads=res))
    # the -l 10 machine (57,090 states here): the largest machine in the suite, 5 CTAs per read on the GPU
    tf10 = tempfile.NamedTemporaryFile("wb", suffix=".json", delete=False)
    tf10.write(gzip.open(os.path.join(HERE, "machines", "l10c4.json.gz"), "rb").read())
    tf10.close()
    mg.machine_args = lambda recipe: ["--machine", tf10.name]
    rng10 = np.random.default_rng(0xD5A57012 + 10)
    enc10 = mg.ref_encode(["l10c4"], [synth.random_bits(rng10, 40) for _ in range(2)])
    reads10 = [(f"r{i}", synth.mutate(e, rng10, sub_rate=0.02, del_rate=0.02, max_del=2)) for i, e in enumerate(enc10)]
    res = mg.ref_viterbi(["l10c4"], dict(length=10), True, reads10)
    for r, (_n, seq) in zip(res, reads10):
        r["seq"] = seq
    _a, f = mg.flag_args(dict(length=10), True)
    cases.append(dict(name="l10_global", recipe=["l10c4"], flags=f, global_=True,
                      note="dnastore -l 10, 57,090 states, k = 5: the largest machine in the suite", reads=res))
    os.unlink(tf10.name)
    json.dump(dict(generator="tests/golden/make_golden_cfg5.py", cases=cases),
              open(os.path.join(HERE, "viterbi_golden_cfg5.json"), "w"))
    os.unlink(tf.name)
    for c in cases:
        print(c["name"], [(r["name"], r["loglike"] if isinstance(r["loglike"], str) else round(r["loglike"], 3)) for r in c["reads"]])


if __name__ == "__main__":
    main()
