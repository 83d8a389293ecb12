"""CPU: the oracle (oracle/viterbi_oracle.c) pinned against the UNMODIFIED reference.

Golden vectors (tests/golden/viterbi_golden.json, made by tests/golden/make_golden.py by
running oracle/_ref/refdriver, i.e. the reference's own ViterbiMatrix) hold, per read,
the decoded string, the fp64 log-likelihood as a hex float and the traceback path.
The bar is bit-exact on all three, plus every DP cell of two short reads.
"""
import numpy as np
import pytest

import dnab_testutil as util

CASES = [c["name"] for c in util.load_golden()]
# the 46,670-state reads take a few seconds each in the oracle: keep one in the CPU suite
SLOW_LIMIT = {"cfg2_global_subs": 1, "cfg3_global_indels": 2, "cfg4_global_dels": 2, "l10_global": 1}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    case = util.golden_case(name)
    compiled = util.compiled_for_case(case)
    reads = case["reads"][:SLOW_LIMIT.get(name, len(case["reads"]))]
    for r in reads:
        o = util.oracle_viterbi(compiled, r["seq"])
        assert o["rc"] in (0, 1), (name, r["name"], o["rc"])
        assert o["decoded"] == r["decoded"], (name, r["name"])
        assert util.hexf(o["loglike"]) == util.hexf(r["loglike_hex"]), (name, r["name"], o["loglike"], r["loglike"])
        assert o["path"] == r["path"], (name, r["name"])
        if r["loglike"] == "-inf":
            assert o["rc"] == 1 and o["decoded"] == ""


def test_kats_decode_to_hello():
    """The reference's 11 Viterbi known-answer tests all decode 'HELLO' (Makefile:146-186)."""
    padded = "^00010010101000100011001000110010111100100$"
    exact = "^0001001010100010001100100011001011110010$"
    for name in CASES:
        if name.startswith("kat") or name == "cfg1":
            want = padded if name in ("kat147", "kat148", "cfg1") else exact
            assert util.golden_case(name)["reads"][0]["decoded"] == want


@pytest.mark.parametrize("tag", ["l4c4_global", "l4c4_local"])
def test_oracle_cells_bit_exact(tag):
    import json
    z = np.load(f"{util.GOLDEN}/cells_{tag}.npz")
    flags = json.loads(str(z["flags"]))
    compiled = util.compiled_for([str(x) for x in z["recipe"]], flags, bool(z["global_"]))
    seq = str(z["seq"])
    o = util.oracle_viterbi(compiled, seq, want_cells=True)
    ref = z["cells"].reshape(o["cells"].shape)
    assert o["cells"].tobytes() == ref.tobytes()
    assert util.hexf(o["loglike"]) == util.hexf(str(z["loglike_hex"]))


def test_oracle_empty_read():
    """L = 0 (the reference CLI segfaults on an empty record; the lattice is still well defined)."""
    compiled = util.compiled_for(["l4c4"], dict(length=4), True)
    o = util.oracle_viterbi(compiled, "")
    assert o["rc"] == 0 and o["decoded"] == "^$" and o["loglike"] < 0


@pytest.mark.parametrize("name,n", [("cfg1_bench", 64), ("cfg4_bench", 6), ("cfg3_bench", 3), ("cfg5_bench", 3), ("cfg2_bench", 1)])
def test_oracle_matches_round2_goldens(name, n):
    """The benchmark read distributions at full length (make_golden_r2.py, unmodified reference): the oracle
    reproduces log-likelihood, decoded string and traceback path bit for bit (a bounded sample per machine keeps
    the CPU suite short; the GPU tests check every read)."""
    case = util.golden_r2_case(name)
    compiled = util.compiled_for_case(case)
    for r in case["reads"][:n]:
        o = util.oracle_viterbi(compiled, r["seq"])
        assert o["decoded"] == r["decoded"], (name, r["name"])
        assert util.hexf(o["loglike"]) == util.hexf(r["loglike_hex"]), (name, r["name"])
        assert o["path"] == r["path"], (name, r["name"])


def test_oracle_cfg2_cells_hash():
    """Every DP cell of a short read on the 46,670-state machine: the oracle's matrix hashes to the reference's."""
    import hashlib
    for c in util.load_golden_r2()["cells"]:
        compiled = util.compiled_for(c["recipe"], c["flags"], c["global_"])
        o = util.oracle_viterbi(compiled, c["seq"], want_cells=True)
        assert o["cells"].size == c["n_cells"]
        assert hashlib.sha256(o["cells"].tobytes()).hexdigest() == c["sha256"], c["name"]
        assert util.hexf(o["loglike"]) == util.hexf(c["loglike_hex"])
