#!/usr/bin/env python
"""Regenerates the committed golden fixtures from the UNMODIFIED reference.

Run in the build container (where /root/reference exists) after
`make -C oracle ref`:

    python tests/golden/make_golden.py

Everything it writes is derived by RUNNING the reference (oracle/_ref/dnastore and
oracle/_ref/refdriver, both compiled from /root/reference by oracle/Makefile) on the
reference's own data files or on seeded synthetic reads:

  machines/*.json.gz       base transducers, re-emitted through the reference's own
                           `--load-machine X --save-machine -` and gzip-compressed
                           (inputs for compose/compile tests; no reference source code)
  composed_sha256.json     sha256 of the machines the reference composes (incl. its own
                           goldens data/mr2l4c4.json, h74l4c4.json, s16mr2l4c4.json,
                           s16h74l4c4.json) -- pins Machine::compose
  viterbi_golden.json      per case: machine recipe, error flags, reads, and for every
                           read the reference's decoded string, log-likelihood (hex
                           float) and traceback path.  Cases = the reference's 11
                           Viterbi known-answer tests (Makefile:147,148,154,169,170,171,
                           177,178,184,185,186), BASELINE config 1, and seeded synthetic
                           reads on every BASELINE machine family
  cells_*.npz              every DP cell of one short read, reference layout
  tables_golden.json       score-table anchors
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from benchdata import synth  # noqa: E402

REF = "/root/reference"
DATA = os.path.join(REF, "data")
BIN = os.path.join(ROOT, "oracle", "_ref", "dnastore")
DRV = os.path.join(ROOT, "oracle", "_ref", "refdriver")
OUT = os.path.dirname(os.path.abspath(__file__))
MACH = os.path.join(OUT, "machines")

BASE_MACHINES = ["l4c4", "mixradar2", "mixradar6", "hamming74", "sync16", "flusher", "water64.1", "echobits", "l1c0t0"]

DEFAULT_FLAGS = dict(length=12, sub=.01, iv=10., dup=.001, delopen=.001, delext=.01)


def run(cmd, stdin=None):
    return subprocess.run(cmd, input=stdin, capture_output=True, text=True, check=True).stdout


def machine_args(recipe):
    base, comps = recipe[0], recipe[1:]
    a = ["--machine", os.path.join(DATA, base + ".json")]
    for c in comps:
        a += ["--compose", os.path.join(DATA, c + ".json")]
    return a


def flag_args(flags, global_):
    f = dict(DEFAULT_FLAGS)
    f.update(flags)
    a = ["-l", str(f["length"]), "--sub", repr(f["sub"]), "--iv", repr(f["iv"]), "--dup", repr(f["dup"]),
         "--delopen", repr(f["delopen"]), "--delext", repr(f["delext"])]
    if global_:
        a.append("--global")
    return a, f


def ref_viterbi(recipe, flags, global_, reads, cells_file=None):
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as fa:
        for name, seq in reads:
            fa.write(f">{name}\n{seq}\n")
        path = fa.name
    fargs, _ = flag_args(flags, global_)
    cmd = [DRV, "viterbi"] + machine_args(recipe) + fargs + ["--fasta", path, "--path"]
    if cells_file:
        cmd += ["--cells", cells_file]
    out = run(cmd)
    os.unlink(path)
    res = []
    for ln in out.split("\n"):
        if not ln:
            continue
        name, ll, llhex, decoded, pathstr = ln.split("\t")
        triples = [tuple(int(x) for x in t.split(":")) for t in pathstr.split()] if pathstr else []
        # the reference logs the start cell once before its loop and again on the first iteration
        triples = triples[1:]
        res.append(dict(name=name, loglike=ll, loglike_hex=llhex, decoded=decoded,
                        path=[list(t) for t in triples]))
    return res


def ref_encode(recipe, payloads):
    out = run([DRV, "encode"] + machine_args(recipe), stdin="\n".join(payloads) + "\n")
    return [ln for ln in out.split("\n") if ln]


def read_fasta_text(path):
    """(name, concatenated sequence) pairs -- plain FASTA, multi-line records joined."""
    recs, name, seq = [], None, []
    for ln in open(path):
        ln = ln.rstrip("\n")
        if ln.startswith(">"):
            if name is not None:
                recs.append((name, "".join(seq)))
            name, seq = ln[1:].split()[0], []
        else:
            seq.append(ln.strip())
    if name is not None:
        recs.append((name, "".join(seq)))
    return recs


def main():
    os.makedirs(MACH, exist_ok=True)
    # ---- base machines, re-emitted by the reference itself ------------------------------
    for m in BASE_MACHINES:
        text = run([BIN, "-v0", "--load-machine", os.path.join(DATA, m + ".json"), "--save-machine", "-"])
        with gzip.GzipFile(os.path.join(MACH, m + ".json.gz"), "wb", mtime=0) as f:
            f.write(text.encode())

    # ---- composed machines: sha256 of what the reference composes ----------------------------
    recipes = {
        "mr2l4c4": ["l4c4", "mixradar2"],
        "h74l4c4": ["l4c4", "hamming74"],
        "s16mr2l4c4": ["l4c4", "sync16", "flusher", "mixradar2"],
        "s16h74l4c4": ["l4c4", "sync16", "flusher", "hamming74"],
        "cfg2_flusher_mixradar6_l4c4": ["l4c4", "flusher", "mixradar6"],
        "cfg4_water64.1_l4c4": ["l4c4", "water64.1"],
    }
    sha = {}
    for name, recipe in recipes.items():
        with tempfile.NamedTemporaryFile(suffix=".json", delete=False) as tf:
            tmp = tf.name
        run([DRV, "compose"] + machine_args(recipe) + ["--save", tmp])
        text = open(tmp).read()
        os.unlink(tmp)
        golden = os.path.join(DATA, name + ".json")
        if os.path.exists(golden):
            assert open(golden).read() == text, f"reference compose != its own golden {name}"
        sha[name] = dict(recipe=recipe, sha256=hashlib.sha256(text.encode()).hexdigest(), n_states=text.count('{"n":'))
    json.dump(sha, open(os.path.join(OUT, "composed_sha256.json"), "w"), indent=1)

    cases = []

    def add_case(name, recipe, flags, global_, reads, expect=None, note=""):
        res = ref_viterbi(recipe, flags, global_, reads)
        _, f = flag_args(flags, global_)
        for r, (rn, seq) in zip(res, reads):
            r["seq"] = seq
            if expect is not None:
                assert r["decoded"] == expect, (name, r["decoded"], expect)
        cases.append(dict(name=name, recipe=recipe, flags=f, global_=bool(global_), note=note, reads=res))
        print(f"  {name}: {len(res)} reads", flush=True)

    # ---- the reference's 11 Viterbi known-answer tests + BASELINE config 1 --------------------
    padded = open(os.path.join(DATA, "hello.padded.bits")).read().strip()
    exact = open(os.path.join(DATA, "hello.exact.bits")).read().strip()
    NOERRS = dict(sub=0., dup=0., delopen=0.)
    ONLYDUPS = dict(sub=0., delopen=0.)
    kats = [
        ("kat147", ["l4c4"], NOERRS, True, "hello.fa", padded),
        ("kat148", ["l4c4"], ONLYDUPS, True, "hello.dup.fa", padded),
        ("kat154", ["mr2l4c4"], NOERRS, True, "hello.mr2.fa", exact),
        ("kat169", ["h74l4c4"], NOERRS, True, "hello.h74.fa", exact),
        ("kat170", ["h74l4c4"], {}, False, "hello.h74.fa", exact),
        ("kat171", ["h74l4c4"], {}, False, "hello.h74.sub.fa", exact),
        ("kat177", ["s16mr2l4c4"], NOERRS, True, "hello.s16mr2.fa", exact),
        ("kat178", ["s16mr2l4c4"], {}, False, "hello.s16mr2.fa", exact),
        ("kat184", ["s16h74l4c4"], NOERRS, True, "hello.s16h74.fa", exact),
        ("kat185", ["s16h74l4c4"], {}, False, "hello.s16h74.fa", exact),
        ("kat186", ["s16h74l4c4"], {}, False, "hello.s16h74.del.fa", exact),
        ("cfg1", ["l4c4"], dict(length=4), True, "hello.fa", padded),
    ]
    composed_recipe = {k: v["recipe"] for k, v in sha.items()}
    for name, recipe, flags, glob, fa, expect in kats:
        rec = composed_recipe.get(recipe[0], recipe)  # composites are rebuilt from base machines
        add_case(name, rec, flags, glob, read_fasta_text(os.path.join(DATA, fa)), expect,
                 note=f"reference Makefile known-answer test on data/{fa}")

    # ---- seeded synthetic reads on each machine family ---------------------------------------
    rng = np.random.default_rng(0xD5A57012)

    def synth_reads(recipe, nbits, n, **mut):
        payloads = [synth.random_bits(rng, nbits) for _ in range(n)]
        enc = ref_encode(recipe, payloads)
        return [(f"r{i}", synth.mutate(e, rng, **mut)) for i, e in enumerate(enc)]

    allerr = dict(sub_rate=0.02, dup_rate=0.01, max_dup=2, del_rate=0.01, max_del=4)
    l4 = dict(length=4)
    add_case("l4c4_global_mixed", ["l4c4"], l4, True, synth_reads(["l4c4"], 96, 24, **allerr))
    add_case("l4c4_local_mixed", ["l4c4"], l4, False, synth_reads(["l4c4"], 64, 8, **allerr))
    add_case("l4c4_len12_local", ["l4c4"], {}, False, synth_reads(["l4c4"], 64, 4, **allerr),
             note="default -l 12: k=4 on a 4-context machine")
    add_case("l4c4_edge", ["l4c4"], l4, True,
             [("one", "T"), ("garbage", "ACGTACGTTTTTGGGGCCCCAAAA"), ("short", "TGTC")],
             note="very short reads, a read that is not a codeword (an EMPTY record segfaults the reference "
                  "CLI, so empty reads are checked GPU-vs-oracle only)")
    add_case("l4c4_noerr_garbage", ["l4c4"], dict(length=4, **NOERRS), True,
             [("garbage", "ACGTACGTTTTTGGGGCCCCAAAA"), ("hello", read_fasta_text(os.path.join(DATA, "hello.fa"))[0][1])],
             note="-inf scores: the first read has no valid decoding (loglike -inf, empty string)")
    cfg2 = composed_recipe["cfg2_flusher_mixradar6_l4c4"]
    add_case("cfg2_global_subs", cfg2, l4, True, synth_reads(cfg2, 204, 3, sub_rate=0.01),
             note="BASELINE config 2: 46,670 states")
    cfg3 = composed_recipe["s16h74l4c4"]
    add_case("cfg3_global_indels", cfg3, l4, True,
             synth_reads(cfg3, 92, 4, sub_rate=0.01, dup_rate=0.01, max_dup=2, del_rate=0.01, max_del=4),
             note="BASELINE config 3: hub states with 98 null-in transitions")
    cfg4 = composed_recipe["cfg4_water64.1_l4c4"]
    add_case("cfg4_global_dels", cfg4, l4, True, synth_reads(cfg4, 64, 4, sub_rate=0.01, del_rate=0.01, max_del=4),
             note="BASELINE config 4: watermark composite, 64-bit payloads")
    add_case("mr2l4c4_local", composed_recipe["mr2l4c4"], l4, False,
             synth_reads(composed_recipe["mr2l4c4"], 60, 6, **allerr))
    json.dump(dict(generator="tests/golden/make_golden.py", cases=cases),
              open(os.path.join(OUT, "viterbi_golden.json"), "w"))

    # ---- full DP matrices of short reads --------------------------------------------------------
    for tag, recipe, flags, glob, seq in [
        ("l4c4_global", ["l4c4"], l4, True, cases[12]["reads"][0]["seq"][:40]),
        ("l4c4_local", ["l4c4"], {}, False, cases[13]["reads"][0]["seq"][:24]),
    ]:
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as tf:
            tmp = tf.name
        res = ref_viterbi(recipe, flags, glob, [("x", seq)], cells_file=tmp)
        raw = np.fromfile(tmp, dtype=np.float64)
        os.unlink(tmp)
        _, f = flag_args(flags, glob)
        np.savez_compressed(os.path.join(OUT, f"cells_{tag}.npz"), cells=raw, seq=np.array(seq),
                            recipe=np.array(recipe), flags=np.array(json.dumps(f)), global_=np.array(glob),
                            loglike_hex=np.array(res[0]["loglike_hex"]))
    print("done")


if __name__ == "__main__":
    main()
