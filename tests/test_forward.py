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


# ----------------------------------------------------------------------------------------------
# posterior (soft) decoding: per read base, the posterior of the class of the move that emitted it
# ----------------------------------------------------------------------------------------------
def brute_force_posterior(compiled, seq, classes, floor=1e-13):
    """Every path of the lattice, enumerated explicitly in probability space (exact exp, no table, no closure
    solve): returns (likelihood, [L, n_classes] posterior of the class of the move that emits each base).  Only
    for tiny machines and reads: deletion cycles make the path set infinite, paths lighter than `floor` are cut."""
    import sys
    t = compiled.t
    N, k = t.n_states, t.k
    eo, no = _arr(t.emit_off, N + 1, np.int64), _arr(t.null_off, N + 1, np.int64)
    es, ns = _arr(t.emit_src, t.n_emit, np.int64), _arr(t.null_src, t.n_null, np.int64)
    esc, nsc = _arr(t.emit_score, t.n_emit, np.float64), _arr(t.null_score, t.n_null, np.float64)
    eb, ein = _arr(t.emit_base, t.n_emit, np.int64), _arr(t.emit_in, t.n_emit, np.int64)
    ctx = _arr(t.ctx, N * k, np.int64).reshape(N, k) if k else np.zeros((N, 0), np.int64)
    mdl = _arr(t.mdl, N, np.int64)
    sub = np.exp(np.array(list(t.sub))).reshape(4, 4)
    ln = np.exp(_arr(t.len, k, np.float64)) if k else np.zeros(0)
    pNoGap, pOpen, pExt, pEnd, pDup = (np.exp(v) for v in (t.noGap, t.delOpen, t.delExtend, t.delEnd, t.tanDup))
    out_e = [[] for _ in range(N)]
    out_n = [[] for _ in range(N)]
    for dst in range(N):
        for e in range(eo[dst], eo[dst + 1]):
            out_e[es[e]].append((dst, np.exp(esc[e]), int(eb[e]), classes.index(chr(ein[e])) if ein[e] else 0))
        for e in range(no[dst], no[dst + 1]):
            out_n[ns[e]].append((dst, np.exp(nsc[e])))
    tok = util.tokens(seq)
    L = len(seq)
    dup = len(classes) - 1
    acc = np.zeros((L, len(classes)))
    total = [0.0]
    sys.setrecursionlimit(100000)

    def walk(s, pos, mut, p, labels):
        if p < floor:
            return
        if mut == 0 and pos == L and (t.local or s == N - 1):
            total[0] += p
            for q, c in enumerate(labels):
                acc[q, c] += p
            if not t.local:
                pass  # the end state may still have outgoing moves: keep walking
        if mut == 0:
            if pos < L:
                for dst, w, b, c in out_e[s]:
                    walk(dst, pos + 1, 0, p * w * pNoGap * sub[b, tok[pos]], labels + [c])
            for dst, w in out_n[s]:
                walk(dst, pos, 0, p * w, labels)
            for dst, w, _b, _c in out_e[s]:
                walk(dst, pos, 1, p * w * pOpen, labels)
            if pos > 0:
                for i in range(mdl[s]):
                    walk(s, pos, 2 + i, p * pDup * ln[i], labels)
        elif mut == 1:
            for dst, w, _b, _c in out_e[s]:
                walk(dst, pos, 1, p * w * pExt, labels)
            for dst, w in out_n[s]:
                walk(dst, pos, 1, p * w, labels)
            walk(s, pos, 0, p * pEnd, labels)
        elif pos < L:
            i = mut - 2
            w = sub[ctx[s, i], tok[pos]]
            if i == 0:
                walk(s, pos + 1, 0, p * w, labels + [dup])
            else:
                walk(s, pos + 1, 2 + i - 1, p * w, labels + [dup])

    starts = range(N) if t.local else [0]
    for s0 in starts:
        walk(s0, 0, 0, 1.0, [])
    return total[0], acc / total[0] if total[0] > 0 else acc


@pytest.mark.parametrize("machine,global_,reads", [("l1c0t0", True, ["ACG", "TT", "GATC"]), ("l1c0t0", False, ["CA"]),
                                                    ("echobits", True, None)])
def test_posterior_specification_against_path_enumeration(machine, global_, reads):
    """The posterior of the class of the move that emits each base (oracle/forward_oracle.c) against an explicit
    enumeration of every lattice path in exact arithmetic: equal within the table log_sum_exp's accuracy, rows sum to 1."""
    import dnastore_b200 as d
    try:
        m = util.machine_from_recipe((machine,))
        compiled = m.compile(d.ErrorFlags(length=2, global_=global_, sub_prob=.05, dup_prob=.02, del_open=.02, del_ext=.1))
    except d.DnabError as e:
        pytest.skip(f"{machine} is not a DNA-output machine for the decoder: {e}")
    if compiled.t.n_states > 64:
        pytest.skip("too large for path enumeration")
    classes = util.posterior_classes(compiled)
    for seq in reads or ["AC", "GGT"]:
        o = util.oracle_posterior(compiled, seq, classes)
        like, exact = brute_force_posterior(compiled, seq, classes)
        if like <= 0:
            continue
        assert abs(np.log(like) - o["loglike"]) < 3e-3, (machine, seq)
        np.testing.assert_allclose(o["post"], exact, atol=4e-3)
        np.testing.assert_allclose(o["post"].sum(axis=1), 1.0, atol=4e-3)


@pytest.mark.parametrize("name", ["l4c4_global_mixed", "l4c4_local_mixed"])
def test_posterior_rows_sum_to_one_and_match_counts(name):
    case = util.golden_case(name)
    compiled = util.compiled_for_case(case)
    classes = util.posterior_classes(compiled)
    k = compiled.t.k
    for r in case["reads"][:3]:
        o = util.oracle_posterior(compiled, r["seq"], classes)
        np.testing.assert_allclose(o["post"].sum(axis=1), 1.0, atol=3e-3)
        # the same posterior weights, binned differently: sum over positions of the non-duplication classes = nNoGap,
        # everything = sum of nSub
        assert abs(o["post"][:, :-1].sum() - o["counts"][2]) < 1e-9 * max(1.0, o["counts"][2])
        assert abs(o["post"].sum() - o["counts"][5 + k:].sum()) < 1e-9 * len(r["seq"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["l4c4_global_mixed", "l4c4_local_mixed", "mr2l4c4_local", "cfg5_l8_global"])
def test_gpu_posterior_matches_specification(name):
    """dnab_posterior_batch on B200 against the specification: same log-likelihood bits, posteriors within 1e-9
    (the kernel sums the states in a different, fixed order), per-base decision = argmax."""
    import dnastore_b200 as d
    case = util.golden_case(name)
    compiled = util.compiled_for_case(case)
    dec = d.Decoder(compiled, device=0)
    classes = dec.posterior_classes()
    assert classes == util.posterior_classes(compiled)
    reads = [r["seq"] for r in case["reads"][:4]]
    out = dec.posterior(reads)
    for i, seq in enumerate(reads):
        o = util.oracle_posterior(compiled, seq, classes)
        assert util.hexf(out["loglike"][i]) == util.hexf(o["loglike"])
        np.testing.assert_allclose(out["post"][i], o["post"], rtol=1e-9, atol=1e-12)
        assert out["decoded"][i] == "".join(classes[j] for j in o["post"].argmax(axis=1))
        assert set(out["decoded"][i]) <= set(classes)
