"""Shared helpers for the tests: golden fixtures, machine recipes, the oracle binding.

The ORACLE (oracle/viterbi_oracle.c) is loaded here and only here (plus
__graft_entry__.smoke and bench.py's CPU baseline): it is the checker, never the
product.
"""
import ctypes as C
import functools
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MACHINES = os.path.join(GOLDEN, "machines")


def machine_path(name):
    return os.path.join(MACHINES, name + ".json.gz")


@functools.lru_cache(maxsize=None)
def load_golden():
    cases = json.load(open(os.path.join(GOLDEN, "viterbi_golden.json")))["cases"]
    extra = os.path.join(GOLDEN, "viterbi_golden_cfg5.json")  # BASELINE config 5's machine (make_golden_cfg5.py)
    if os.path.exists(extra):
        cases = cases + json.load(open(extra))["cases"]
    return cases


@functools.lru_cache(maxsize=None)
def load_golden_r2():
    """Round-2 golden set (make_golden_r2.py): the benchmark read distributions at full length, run through the
    unmodified reference -- dict(cases=[...], cells=[...])."""
    import gzip
    return json.load(gzip.open(os.path.join(GOLDEN, "viterbi_golden_r2.json.gz"), "rt"))


def golden_r2_case(name):
    for c in load_golden_r2()["cases"]:
        if c["name"] == name:
            return c
    raise KeyError(name)


def golden_case(name):
    for c in load_golden():
        if c["name"] == name:
            return c
    raise KeyError(name)


@functools.lru_cache(maxsize=None)
def machine_from_recipe(recipe):
    """recipe = (base, compose1, compose2, ...) as on the dnastore command line."""
    import dnastore_b200 as d
    return d.Machine.load_composed(machine_path(recipe[0]), [machine_path(c) for c in recipe[1:]])


def error_flags(flags, global_):
    import dnastore_b200 as d
    return d.ErrorFlags(length=flags["length"], global_=global_, sub_prob=flags["sub"], iv_ratio=flags["iv"],
                        dup_prob=flags["dup"], del_open=flags["delopen"], del_ext=flags["delext"])


@functools.lru_cache(maxsize=None)
def _compiled_cached(recipe, flag_items, global_):
    m = machine_from_recipe(recipe)
    return m.compile(error_flags(dict(flag_items), global_))


def compiled_for_case(case):
    return _compiled_cached(tuple(case["recipe"]), tuple(sorted(case["flags"].items())), bool(case["global_"]))


def compiled_for(recipe, flags=None, global_=True):
    f = dict(length=12, sub=.01, iv=10., dup=.001, delopen=.001, delext=.01)
    f.update(flags or {})
    return _compiled_cached(tuple(recipe), tuple(sorted(f.items())), bool(global_))


# ----------------------------------------------------------------------------- oracle
@functools.lru_cache(maxsize=None)
def oracle_lib():
    path = os.path.join(ROOT, "oracle", "_build", "libviterbi_oracle.so")
    lib = C.CDLL(path)
    lib.dnab_oracle_viterbi.restype = C.c_int
    lib.dnab_oracle_viterbi.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_char_p, C.c_int,
                                        C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
    lib.dnab_oracle_viterbi_batch.restype = C.c_int
    lib.dnab_oracle_viterbi_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p]
    return lib


_TOK = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate("ACGT"):
    _TOK[ord(_ch)] = _i
    _TOK[ord(_ch.lower())] = _i


def tokens(seq):
    t = _TOK[np.frombuffer(seq.encode(), dtype=np.uint8)] if seq else np.zeros(0, dtype=np.uint8)
    assert (t < 4).all()
    return np.ascontiguousarray(t)


def oracle_viterbi(compiled, seq, want_cells=False, want_path=True):
    """Returns dict(rc, loglike, decoded, path[list of [state,pos,mut]], cells)."""
    lib = oracle_lib()
    t = compiled.t
    tok = tokens(seq)
    L = len(seq)
    ll = C.c_double(0)
    cap = 16 * L + 4096
    dec = C.create_string_buffer(cap)
    dec_len = C.c_int(0)
    pcap = 32 * L + 8192
    path = np.zeros((pcap, 3), dtype=np.int32)
    plen = C.c_int(0)
    cells = np.zeros((L + 1, t.n_states, t.k + 2), dtype=np.float64) if want_cells else None
    rc = lib.dnab_oracle_viterbi(C.addressof(t), tok.ctypes.data, L, C.byref(ll), dec, cap, C.byref(dec_len),
                                 path.ctypes.data if want_path else None, pcap, C.byref(plen),
                                 cells.ctypes.data if want_cells else None)
    return dict(rc=rc, loglike=ll.value, decoded=dec.raw[:dec_len.value].decode("latin1"),
                path=path[:plen.value].tolist(), cells=cells)


@functools.lru_cache(maxsize=None)
def forward_oracle_lib():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libforward_oracle.so"))
    lib.dnab_oracle_forward.restype = C.c_int
    lib.dnab_oracle_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double),
                                        C.POINTER(C.c_long), C.c_void_p]
    lib.dnab_oracle_backward_posterior.restype = C.c_int
    lib.dnab_oracle_backward_posterior.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double,
                                                   C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_long), C.c_char_p, C.c_int,
                                                   C.c_void_p]
    lib.dnab_oracle_backward_counts.restype = C.c_int
    lib.dnab_oracle_backward_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double,
                                                C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_long)]
    return lib


def oracle_fwdback(compiled, seq, max_sweeps=4096, tables=None):
    """Forward cells, backward log-likelihood and posterior expected counts of the error-model events
    [nDelOpen, nTanDup, nNoGap, nDelExtend, nDelEnd, nLen[k], nSub[16]] (oracle/forward_oracle.c)."""
    t = tables if tables is not None else compiled.t
    f = oracle_forward(compiled, seq, want_cells=True, max_sweeps=max_sweeps, tables=t)
    tok = tokens(seq)
    counts = np.zeros(5 + t.k + 16, dtype=np.float64)
    llb = C.c_double(0)
    sw = C.c_long(0)
    rc = forward_oracle_lib().dnab_oracle_backward_counts(C.addressof(t), tok.ctypes.data, len(seq), max_sweeps,
                                                          f["cells"].ctypes.data, f["loglike"], C.byref(llb),
                                                          counts.ctypes.data, C.byref(sw))
    return dict(rc=rc | f["rc"], loglike=f["loglike"], loglike_back=llb.value, counts=counts, sweeps_back=sw.value,
                sweeps=f["sweeps"])


def posterior_classes(compiled):
    """'-' + the input symbols in the order the decoder numbers them (first appearance in the emit list, then the
    null list) + '+': the classes of dnab_posterior_batch."""
    t = compiled.t
    seen = []
    for arr, n in ((t.emit_in, t.n_emit), (t.null_in, t.n_null)):
        for e in range(n):
            ch = arr[e]
            if ch and ch not in seen:
                seen.append(ch)
    return "-" + "".join(chr(c) for c in seen) + "+"


def oracle_posterior(compiled, seq, classes, max_sweeps=4096):
    """The posterior specification (oracle/forward_oracle.c: dnab_oracle_backward_posterior): [L, n_classes]."""
    t = compiled.t
    f = oracle_forward(compiled, seq, want_cells=True, max_sweeps=max_sweeps)
    tok = tokens(seq)
    counts = np.zeros(5 + t.k + 16, dtype=np.float64)
    post = np.zeros((len(seq), len(classes)), dtype=np.float64)
    llb = C.c_double(0)
    sw = C.c_long(0)
    rc = forward_oracle_lib().dnab_oracle_backward_posterior(C.addressof(t), tok.ctypes.data, len(seq), max_sweeps,
                                                             f["cells"].ctypes.data, f["loglike"], C.byref(llb), counts.ctypes.data,
                                                             C.byref(sw), classes.encode("latin1"), len(classes), post.ctypes.data)
    return dict(rc=rc | f["rc"], loglike=f["loglike"], post=post, counts=counts)


def oracle_forward(compiled, seq, want_cells=False, max_sweeps=4096, tables=None):
    """The forward specification (oracle/forward_oracle.c). Returns dict(rc, loglike, sweeps, cells)."""
    t = tables if tables is not None else compiled.t
    tok = tokens(seq)
    L = len(seq)
    ll = C.c_double(0)
    sw = C.c_long(0)
    cells = np.zeros((L + 1, t.n_states, t.k + 2), dtype=np.float64) if want_cells else None
    rc = forward_oracle_lib().dnab_oracle_forward(C.addressof(t), tok.ctypes.data, L, max_sweeps, C.byref(ll), C.byref(sw),
                                                  cells.ctypes.data if want_cells else None)
    return dict(rc=rc, loglike=ll.value, sweeps=sw.value, cells=cells)


def hexf(x):
    """Canonical hex-float string of an fp64 (bit-exact comparisons; accepts C's %a output too)."""
    if isinstance(x, str):
        x = float.fromhex(x) if "x" in x else float(x)
    return float(x).hex()
