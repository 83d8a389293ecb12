"""Forward (sum-product) log-likelihood over the machine-state x DNA-position lattice, SURVEY.md 8a-12.

PARITY UNPINNED: the reference has no such computation, so there is no golden vector and no reference
run to compare with.  What holds the specification (oracle/forward_oracle.c) instead:
  * an INDEPENDENT computation in probability space with a direct sparse linear solve of every
    column's closure (exact exp/log, no table): agreement within the accuracy of the reference's
    table-based log_sum_exp, which drops terms more than 10 nats below the running sum -- so the
    specification may only fall short of the exact value, by a small amount;
  * forward >= Viterbi on every read (a sum over paths contains the best path);
and the CUDA kernel must reproduce the specification BIT FOR BIT (same sweeps, same operand order).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import dnab_testutil as util


def _arr(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, (max(n, 1),))[:n].astype(dtype)


def exact_forward(compiled, seq):
    """Probability-space forward with a direct solve per column; returns the log-likelihood."""
    t = compiled.t
    N, k = t.n_states, t.k
    eo, no = _arr(t.emit_off, N + 1, np.int64), _arr(t.null_off, N + 1, np.int64)
    es, ns = _arr(t.emit_src, t.n_emit, np.int64), _arr(t.null_src, t.n_null, np.int64)
    esc, nsc = _arr(t.emit_score, t.n_emit, np.float64), _arr(t.null_score, t.n_null, np.float64)
    eb = _arr(t.emit_base, t.n_emit, np.int64)
    ed, nd = np.repeat(np.arange(N), np.diff(eo)), np.repeat(np.arange(N), np.diff(no))
    ctx = _arr(t.ctx, N * k, np.int64).reshape(N, k) if k else np.zeros((N, 0), np.int64)
    mdl = _arr(t.mdl, N, np.int64)
    sub = np.exp(np.array(list(t.sub))).reshape(4, 4)
    ln = np.exp(_arr(t.len, k, np.float64)) if k else np.zeros(0)
    pNoGap, pOpen, pExt, pEnd, pDup = (np.exp(v) for v in (t.noGap, t.delOpen, t.delExtend, t.delEnd, t.tanDup))
    E = sp.csr_matrix((np.exp(esc), (ed, es)), shape=(N, N))
    Nn = sp.csr_matrix((np.exp(nsc), (nd, ns)), shape=(N, N))
    eye = sp.identity(N, format="csr")
    A = sp.bmat([[eye - Nn, -pEnd * eye], [-pOpen * E, eye - pExt * E - Nn]], format="csc")
    lu = spla.splu(A)
    tok = util.tokens(seq)
    s = np.zeros(N)
    if t.local:
        s[:] = 1.0
    else:
        s[0] = 1.0
    T = np.zeros((k, N))
    logscale = 0.0
    for pos in range(len(seq) + 1):
        if pos > 0:
            x = tok[pos - 1]
            s0 = np.zeros(N)
            np.add.at(s0, ed, s[es] * np.exp(esc) * pNoGap * sub[eb, x])
            Tn = np.zeros((k, N))
            if k:
                has = mdl > 0
                s0[has] += T[0, has] * sub[ctx[has, 0], x]
                for i in range(k - 1):
                    m = mdl - 1 > i
                    Tn[i, m] = T[i + 1, m] * sub[ctx[m, i + 1], x]
        else:
            s0, Tn = s, np.zeros((k, N))
        sol = lu.solve(np.concatenate([s0, np.zeros(N)]))
        s = sol[:N]
        if pos > 0:
            for i in range(k):
                m = mdl > i
                Tn[i, m] += s[m] * pDup * ln[i]
        T = Tn
        scale = s.max()
        if scale > 0:
            s, T, logscale = s / scale, T / scale, logscale + np.log(scale)
    total = s.sum() if t.local else s[N - 1]
    return (np.log(total) + logscale) if total > 0 else -np.inf


def _reads(name, n):
    return [r["seq"] for r in util.golden_case(name)["reads"]][:n]


@pytest.mark.parametrize("name,n", [("l4c4_global_mixed", 6), ("l4c4_local_mixed", 4), ("mr2l4c4_local", 2), ("kat148", 1)])
def test_forward_specification_against_exact_linear_solve(name, n):
    compiled = util.compiled_for_case(util.golden_case(name))
    for seq in _reads(name, n):
        o = util.oracle_forward(compiled, seq)
        assert o["rc"] == 0
        exact = exact_forward(compiled, seq)
        if np.isinf(exact):
            assert np.isinf(o["loglike"])
            continue
        # the table log_sum_exp ignores terms below e^-10 of the running sum: never above exact, barely below
        assert o["loglike"] <= exact + 1e-9 * max(1.0, abs(exact)), (name, seq)
        assert exact - o["loglike"] < 2e-3, (name, seq, exact, o["loglike"])


@pytest.mark.parametrize("name", ["l4c4_global_mixed", "l4c4_local_mixed", "cfg4_global_dels"])
def test_forward_at_least_viterbi(name):
    case = util.golden_case(name)
    compiled = util.compiled_for_case(case)
    for r in case["reads"][:3]:
        f = util.oracle_forward(compiled, r["seq"])["loglike"]
        v = util.oracle_viterbi(compiled, r["seq"], want_path=False)["loglike"]
        assert f >= v or (np.isinf(f) and np.isinf(v)), (name, f, v)


def _perturbed(compiled, field, index, eps):
    """A copy of the tables with one score shifted by eps (the arrays stay shared with `compiled`)."""
    import ctypes as C
    t = type(compiled.t)()
    C.memmove(C.addressof(t), C.addressof(compiled.t), C.sizeof(t))
    if field == "sub":
        arr = (C.c_double * 16)(*list(t.sub))
        arr[index] += eps
        t.sub = arr
    elif field == "len":
        arr = (C.c_double * t.k)(*[t.len[i] for i in range(t.k)])
        arr[index] += eps
        t.len = C.cast(arr, C.POINTER(C.c_double))
        t._keep = arr
    else:
        setattr(t, field, getattr(t, field) + eps)
    return t


@pytest.mark.parametrize("name,idx", [("l4c4_global_mixed", 0), ("l4c4_global_mixed", 5), ("l4c4_local_mixed", 1)])
def test_backward_and_counts_specification(name, idx):
    """Backward pass + posterior expected counts (the machine-lattice analogue of FwdBackMatrix::counts):
    backward and forward log-likelihoods agree; every read base is emitted by exactly one move (sum of nSub
    = L); every deletion run that opens also ends; and each count is the derivative of the forward
    log-likelihood with respect to its score (central differences; loose, the table log_sum_exp is only
    piecewise smooth)."""
    case = util.golden_case(name)
    compiled = util.compiled_for_case(case)
    seq = case["reads"][idx]["seq"]
    fb = util.oracle_fwdback(compiled, seq)
    assert fb["rc"] == 0
    k = compiled.t.k
    c = fb["counts"]
    assert abs(fb["loglike_back"] - fb["loglike"]) < 2e-3
    assert abs(c[5 + k:].sum() - len(seq)) < 2e-3 * len(seq)
    assert abs(c[0] - c[4]) < 1e-3 * max(1.0, c[0])          # nDelOpen == nDelEnd
    assert abs(c[1] - c[5:5 + k].sum()) < 1e-12 * max(1.0, c[1])  # nTanDup == sum nLen
    eps = 0.02
    checks = [("delOpen", 0, 0), ("tanDup", 0, 1), ("noGap", 0, 2), ("delExtend", 0, 3), ("delEnd", 0, 4), ("len", 0, 5)]
    big = int(np.argmax(c[5 + k:]))
    checks.append(("sub", big, 5 + k + big))
    for field, index, ci in checks:
        up = util.oracle_forward(compiled, seq, tables=_perturbed(compiled, field, index, +eps))["loglike"]
        dn = util.oracle_forward(compiled, seq, tables=_perturbed(compiled, field, index, -eps))["loglike"]
        deriv = (up - dn) / (2 * eps)
        assert abs(deriv - c[ci]) < 0.02 + 0.03 * abs(c[ci]), (field, index, deriv, c[ci])


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,cut", [("l4c4_global_mixed", 24, None), ("l4c4_local_mixed", 8, None), ("l4c4_edge", 3, None),
                                        ("mr2l4c4_local", 3, None), ("cfg3_global_indels", 2, 48), ("cfg4_global_dels", 2, 60),
                                        ("cfg2_global_subs", 1, 40), ("cfg5_l8_global", 2, 40), ("cfg5_l8_local", 1, 30)])
def test_gpu_forward_matches_specification_bit_for_bit(name, n, cut):
    import dnastore_b200 as d
    case = util.golden_case(name)
    compiled = util.compiled_for_case(case)
    reads = [s[:cut] if cut else s for s in _reads(name, n)]
    dec = d.Decoder(compiled, device=0)
    out = dec.forward(reads)
    for i, s in enumerate(reads):
        o = util.oracle_forward(compiled, s)
        assert util.hexf(out["loglike"][i]) == util.hexf(o["loglike"]), (name, i)
        assert out["sweeps"][i] == o["sweeps"] and out["status"][i] == o["rc"], (name, i)


@pytest.mark.gpu
def test_gpu_forward_cells_bit_exact_and_empty_read():
    import dnastore_b200 as d
    case = util.golden_case("l4c4_global_mixed")
    compiled = util.compiled_for_case(case)
    dec = d.Decoder(compiled, device=0)
    seq = case["reads"][1]["seq"]
    out = dec.forward([seq, "", "ACGT"], want_cells=True)
    o = util.oracle_forward(compiled, seq, want_cells=True)
    assert out["cells"].tobytes() == o["cells"].tobytes()
    for i, s in enumerate([seq, "", "ACGT"]):
        assert util.hexf(out["loglike"][i]) == util.hexf(util.oracle_forward(compiled, s)["loglike"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,cut", [("l4c4_global_mixed", 6, None), ("l4c4_local_mixed", 4, None), ("cfg4_global_dels", 1, 40),
                                        ("cfg5_l8_global", 2, 40), ("cfg2_global_subs", 1, 24)])
def test_gpu_fwdback_counts_match_specification(name, n, cut):
    """Backward log-likelihood bit for bit, posterior counts within 1e-9 (the sums run in a different order)."""
    import dnastore_b200 as d
    case = util.golden_case(name)
    compiled = util.compiled_for_case(case)
    reads = [s[:cut] if cut else s for s in _reads(name, n)]
    dec = d.Decoder(compiled, device=0)
    out = dec.fwdback_counts(reads)
    for i, s in enumerate(reads):
        o = util.oracle_fwdback(compiled, s)
        assert out["status"][i] == o["rc"]
        assert util.hexf(out["loglike"][i]) == util.hexf(o["loglike"]), (name, i)
        assert util.hexf(out["loglike_back"][i]) == util.hexf(o["loglike_back"]), (name, i)
        np.testing.assert_allclose(out["counts"][i], o["counts"], rtol=1e-9, atol=1e-12, err_msg=f"{name} read {i}")
