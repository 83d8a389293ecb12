"""Pair-HMM forward / backward / expected counts / Baum-Welch.

CPU: the oracle (oracle/pairhmm_oracle.c) and the host code (Stockholm reader, envelope, counts,
log-sum-exp table) pinned against the UNMODIFIED reference (tests/golden/pairhmm_golden.json, made
by tests/golden/make_golden_pairhmm.py with oracle/_ref/refdriver) and the reference's own goldens.
GPU: the CUDA kernel through the C ABI against the oracle on the same inputs (log-likelihoods bit
for bit: same lookup table, same operand order; counts to 1e-12 -- they go through exp(), where
device and host libm may differ in the last ulp) and against the reference golden.
"""
import ctypes as C
import functools
import json
import os

import numpy as np
import pytest

import dnab_testutil as util
import dnastore_b200 as d

GOLD = json.load(open(os.path.join(util.GOLDEN, "pairhmm_golden.json")))["cases"]
NAMES = [c["name"] for c in GOLD]


def case(name):
    return [c for c in GOLD if c["name"] == name][0]


@functools.lru_cache(maxsize=None)
def oracle():
    lib = C.CDLL(os.path.join(util.ROOT, "oracle", "_build", "libpairhmm_oracle.so"))
    lib.dnab_oracle_pairhmm_fb.restype = C.c_int
    lib.dnab_oracle_pairhmm_fb.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dnab_oracle_lse_table.restype = C.POINTER(C.c_double)
    lib.dnab_oracle_lse_table.argtypes = [C.POINTER(C.c_int)]
    lib.dnab_oracle_lse.restype = C.c_double
    lib.dnab_oracle_lse.argtypes = [C.c_double, C.c_double]
    return lib


def params_for(c):
    f = c["flags"]
    return d.MutatorParams.from_flags(d.ErrorFlags(length=f["length"], sub_prob=f["sub"], iv_ratio=f["iv"],
                                                   dup_prob=f["dup"], del_open=f["delopen"], del_ext=f["delext"]))


def load_db(c, tmp_path):
    p = tmp_path / (c["name"] + ".stk")
    p.write_text(c["stk"])
    return d.PairDb(p)


def oracle_fb(params, al, strict):
    tin, tout, a, b = [np.ascontiguousarray(x) for x in al]
    k = params.max_dup_len
    probs, plen = np.ascontiguousarray(params.probs()), np.ascontiguousarray(params.plen())
    counts = np.zeros(5 + k + 16)
    f, bk = C.c_double(0), C.c_double(0)
    oracle().dnab_oracle_pairhmm_fb(probs.ctypes.data, plen.ctypes.data, k, tin.ctypes.data, len(tin), tout.ctypes.data,
                                    len(tout), a.ctypes.data, b.ctypes.data, 0 if strict else k, C.byref(f), C.byref(bk),
                                    counts.ctypes.data, None, None)
    return f.value, bk.value, counts


# ------------------------------------------------------------------------------------------ CPU
def test_lse_table_matches_oracle_and_reference_semantics():
    n = C.c_int32(0)
    mine = np.ctypeslib.as_array(d.lib.dnab_lse_table(C.byref(n)), shape=(100001,)).copy()
    m = C.c_int(0)
    ref = np.ctypeslib.as_array(oracle().dnab_oracle_lse_table(C.byref(m)), shape=(100001,)).copy()
    assert n.value == m.value == 100001 and mine.tobytes() == ref.tobytes()
    assert mine[0] == np.log(2.0)
    lse = oracle().dnab_oracle_lse
    assert lse(-np.inf, -np.inf) == -np.inf and lse(-np.inf, 1.5) == 1.5 and lse(0.0, 20.0) == 20.0
    # interpolated table: close to, but not equal to, the exact value
    assert abs(lse(0.3, 1.0) - np.logaddexp(0.3, 1.0)) < 1e-9


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name, tmp_path):
    c = case(name)
    db, params = load_db(c, tmp_path), params_for(c)
    assert len(db) == len(c["alignments"])
    for i, g in enumerate(c["alignments"]):
        f, b, counts = oracle_fb(params, db.alignment(i), c["strict"])
        assert util.hexf(f) == util.hexf(g["fwd_hex"]) and util.hexf(b) == util.hexf(g["back_hex"]), (name, i)
        assert counts.tolist() == g["counts"], (name, i)


def test_reference_count_goldens_to_six_digits(tmp_path):
    """reference Makefile:157 (data/dup.counts.json values, 6 significant digits)."""
    c = case("dup")
    g = c["alignments"][0]["counts"]
    k = c["flags"]["length"] // 2
    assert f"{g[2]:.6g}" == "29" and f"{g[1]:.6g}" == "1"  # nNoGap, nTanDup of the clean tandem duplication
    assert f"{sum(g[5:5 + k]):.6g}" == "1"


def test_counts_json_format():
    cnt = d.MutatorCounts()
    cnt.max_dup_len = 3
    cnt.n_no_gap = 29
    cnt.n_sub[0] = 7
    txt = cnt.to_json()
    assert txt.startswith('{\n "nDelOpen": 0,\n "nTanDup": 0,\n "nNoGap": 29,') and '"nMatch": 7,' in txt


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_forward_backward_counts(name, tmp_path):
    c = case(name)
    db, params = load_db(c, tmp_path), params_for(c)
    aligns = [db.alignment(i) for i in range(len(db))]
    fwd, back, counts, ms = d.pairhmm_fb_batch(params, aligns, strict=c["strict"])
    for i, g in enumerate(c["alignments"]):
        of, ob, oc = oracle_fb(params, aligns[i], c["strict"])
        assert util.hexf(fwd[i]) == util.hexf(of) and util.hexf(back[i]) == util.hexf(ob), (name, i)
        np.testing.assert_allclose(counts[i].flat(), oc, rtol=1e-12, atol=1e-300)
        # and the unmodified reference: log-likelihoods within the north-star's 1e-9 (in fact bit-equal here)
        rf, rb = float.fromhex(g["fwd_hex"]), float.fromhex(g["back_hex"])
        assert abs(fwd[i] - rf) <= 1e-12 * abs(rf) and abs(back[i] - rb) <= 1e-12 * abs(rb)
        np.testing.assert_allclose(counts[i].flat(), np.array(g["counts"]), rtol=1e-11, atol=1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in NAMES if "fit" in case(n)])
def test_gpu_baum_welch_matches_reference(name, tmp_path):
    c = case(name)
    db, params = load_db(c, tmp_path), params_for(c)
    fitted, iters = d.baum_welch(params, db, strict=c["strict"])
    want = [float(x) for x in c["fit"]]
    got = list(fitted.probs()) + list(fitted.plen())
    np.testing.assert_allclose(got, want, rtol=1e-10)
    assert 1 <= iters <= 100


@pytest.mark.gpu
def test_gpu_expected_counts_reference_goldens(tmp_path):
    """reference Makefile:157-159: --error-counts output, 6 significant digits, through the C ABI."""
    c = case("dup.sub")
    total, ll = d.expected_counts(params_for(c), load_db(c, tmp_path), strict=False)
    txt = total.to_json()
    assert '"nTanDup": 1.25882,' in txt and '"nNoGap": 29,' in txt and '"nMatch": 31.2588,' in txt


@pytest.mark.gpu
def test_gpu_cli_error_counts_and_fit_error(tmp_path):
    """bin/dnastore-b200 --error-counts / --fit-error (reference t/dnastore.cpp:135-149, Makefile:157-163): the CLI prints
    what the C ABI returns, in the reference's JSON layout."""
    import subprocess
    exe = os.path.join(util.ROOT, "bin", "dnastore-b200")
    c = case("dup.sub")
    stk = tmp_path / "dup.sub.stk"
    stk.write_text(c["stk"])
    out = subprocess.run([exe, "-v0", "-l6", "--error-sub-prob", "1e-9", "--error-dup-prob", "1e-9", "--error-del-open", "1e-9",
                          "--error-counts", str(stk)], capture_output=True, text=True, check=True).stdout
    total, _ll = d.expected_counts(params_for(c), load_db(c, tmp_path), strict=False)
    assert out == total.to_json()
    assert '"nTanDup": 1.25882,' in out and '"nMatch": 31.2588,' in out
    c = case("tiny_fit")
    stk = tmp_path / "tiny.stk"
    stk.write_text(c["stk"])
    out = subprocess.run([exe, "-v0", "--fit-error", str(stk), "--strict-guides"], capture_output=True, text=True, check=True).stdout
    fitted, _it = d.baum_welch(params_for(c), load_db(c, tmp_path), strict=True)
    assert out == fitted.to_json()


def _mixed_alignments(tmp_path, n=330, seed=5):
    """Alignments of very different lengths (3-150 nt) with substitutions, duplications and deletions: a warp of the kernel
    holds 32 of them and walks the union of their envelopes."""
    from benchdata import synth
    rng = np.random.default_rng(seed)
    pool = synth.load_pool("cfg1_l4c4_200b")
    rows = []
    for i in range(n):
        strand = pool[i % len(pool)][:int(rng.integers(3, 150))]
        a, b = synth.mutate_aligned(strand, rng, sub_rate=0.03, dup_rate=0.03, max_dup=3, del_rate=0.03, max_del=4)
        if not a.replace("-", "") or not b.replace("-", ""):
            continue
        rows.append(f"# STOCKHOLM 1.0\nin  {a}\nout {b}\n//\n")
    p = tmp_path / "mixed.stk"
    p.write_text("".join(rows))
    db = d.PairDb(p)
    return [db.alignment(i) for i in range(len(db))]


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [False, True])
def test_gpu_mixed_batch_matches_oracle_whatever_the_chunking(tmp_path, strict):
    aligns = _mixed_alignments(tmp_path)
    assert len(aligns) > 300 and len({len(a[0]) for a in aligns}) > 50
    params = d.MutatorParams.from_flags(d.ErrorFlags(length=8, sub_prob=.03, dup_prob=.02, del_open=.02, del_ext=.3))
    try:
        runs = []
        for cells in (0, 700, 1):  # automatic; a few warps per launch; one block (two warps) per launch
            d.pairhmm_set_chunk_cells(cells)
            runs.append(d.pairhmm_fb_batch(params, aligns, strict=strict))
    finally:
        d.pairhmm_set_chunk_cells(0)
    fwd, back, counts, _ms = runs[0]
    for i, al in enumerate(aligns):
        of, ob, oc = oracle_fb(params, al, strict)
        assert util.hexf(fwd[i]) == util.hexf(of) and util.hexf(back[i]) == util.hexf(ob), i
        np.testing.assert_allclose(counts[i].flat(), oc, rtol=1e-12, atol=1e-300)
    for f2, b2, c2, _ in runs[1:]:
        assert np.asarray(f2).tobytes() == np.asarray(fwd).tobytes() and np.asarray(b2).tobytes() == np.asarray(back).tobytes()
        assert all(np.array_equal(x.flat(), y.flat()) for x, y in zip(c2, counts))
