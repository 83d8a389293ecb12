#!/usr/bin/env python
"""Generates benchdata/pools/*.txt.gz: pools of reference-ENCODED payloads (DNA strings, one per
line) for the BASELINE.json configurations.  Run once in the build container after
`make -C oracle ref`; uses the reference's own encoder (Encoder<FastaWriter>::encodeSymbolString,
reference src/encoder.h:239-242) through oracle/_ref/refdriver.  Payloads are i.i.d. uniform bits from
numpy's PCG64 seeded with 0xD5A57012 + config index.
"""
import gzip
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchdata import synth  # noqa: E402

DATA = "/root/reference/data"
DRV = os.path.join(ROOT, "oracle", "_ref", "refdriver")

POOLS = {
    # name: (config index, recipe, payload bits, pool size)
    "cfg2_flusher_mixradar6_l4c4_204b": (2, ["l4c4", "flusher", "mixradar6"], 204, 4096),
    "cfg3_s16h74l4c4_92b": (3, ["l4c4", "sync16", "flusher", "hamming74"], 92, 1024),
    "cfg4_water64.1_l4c4_64b": (4, ["l4c4", "water64.1"], 64, 1024),
    "cfg1_l4c4_200b": (1, ["l4c4"], 200, 1024),
    # BASELINE configs[4]: the -l 8 machine, emitted once by the reference builder in this container
    # (`oracle/_ref/dnastore -v0 -l 8 --save-machine`, 10,746 states here; SURVEY.md 8c: the builder's output
    # depends on std::sort tie order, so the JSON under tests/golden/machines/ is the unit of reproducibility)
    "cfg5_l8c4_150b": (5, ["l8c4"], 150, 1024),
}


def machine_json(name):
    """Path of a plain-JSON machine: the reference's data/ directory, else tests/golden/machines/*.json.gz unpacked."""
    p = f"{DATA}/{name}.json"
    if os.path.exists(p):
        return p
    import tempfile
    src = os.path.join(ROOT, "tests", "golden", "machines", name + ".json.gz")
    tf = tempfile.NamedTemporaryFile("wb", suffix=".json", delete=False)
    tf.write(gzip.open(src, "rb").read())
    tf.close()
    return tf.name


def main():
    only = set(sys.argv[1:])
    os.makedirs(os.path.join(ROOT, "benchdata", "pools"), exist_ok=True)
    for name, (idx, recipe, nbits, n) in POOLS.items():
        rng = np.random.default_rng(0xD5A57012 + idx)
        payloads = [synth.random_bits(rng, nbits) for _ in range(n)]
        if only and name not in only:
            continue
        args = [DRV, "encode", "--machine", machine_json(recipe[0])]
        for c in recipe[1:]:
            args += ["--compose", machine_json(c)]
        if name.startswith("cfg5"):
            pass  # the encoder does not depend on -l (only the error model does)
        out = subprocess.run(args, input="\n".join(payloads) + "\n", capture_output=True, text=True, check=True).stdout
        enc = [ln for ln in out.split("\n") if ln]
        assert len(enc) == n, (name, len(enc))
        with gzip.GzipFile(os.path.join(ROOT, "benchdata", "pools", name + ".txt.gz"), "wb", mtime=0) as f:
            f.write(("\n".join(enc) + "\n").encode())
        with gzip.GzipFile(os.path.join(ROOT, "benchdata", "pools", name + ".payloads.txt.gz"), "wb", mtime=0) as f:
            f.write(("\n".join(payloads) + "\n").encode())
        lens = [len(e) for e in enc]
        print(name, n, "reads, length", min(lens), "-", max(lens), "mean", sum(lens) / n, flush=True)


if __name__ == "__main__":
    main()
