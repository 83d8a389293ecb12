"""CPU: the host side above the C ABI -- machine JSON, compose, table compiler, FASTA, packing."""
import ctypes as C
import gzip
import hashlib
import json
import os
import re

import numpy as np
import pytest

import dnastore_b200 as d
import dnab_testutil as util


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(util.ROOT, "include", "dnastore_b200.h")).read()
    names = set(re.findall(r"\b(dnab_[a-z_0-9]+)\s*\(", header))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(d.lib, n), f"{n} declared in include/dnastore_b200.h but not exported"


def test_machine_json_round_trip():
    """reference Makefile:135: --load-machine X --save-machine - reproduces X."""
    for name in ["l4c4", "mixradar6", "sync16", "flusher", "water64.1"]:
        text = gzip.open(util.machine_path(name), "rt").read()
        assert d.Machine.from_json(text).to_json() == text


def test_lenient_json_accepts_missing_and_trailing_commas():
    text = '{"state":[\n {"n":0,"id":"B","trans":[{"in":"^","out":"^","to":1}]}\n {"n":1,"id":"E","trans":[]},\n]}'
    m = d.Machine.from_json(text)
    assert m.n_states == 2
    with pytest.raises(d.DnabError):
        d.Machine.from_json('{"state":[{"n":1,"trans":[]}]}')  # n out of sequence


def test_compose_matches_reference_sha256():
    """reference Makefile:151,166,174,181 compose goldens + the BASELINE config 2/4 composites."""
    want = json.load(open(os.path.join(util.GOLDEN, "composed_sha256.json")))
    for name, w in want.items():
        m = util.machine_from_recipe(tuple(w["recipe"]))
        assert m.n_states == w["n_states"], name
        assert hashlib.sha256(m.to_json().encode()).hexdigest() == w["sha256"], name


def test_input_alphabet_and_context():
    m = util.machine_from_recipe(("l4c4",))
    assert m.n_states == 384 and m.max_left_context == 4
    assert m.input_alphabet() == "$01AB^"
    assert util.machine_from_recipe(("l4c4", "water64.1")).input_alphabet() == "$01^"


def test_score_tables_anchor_values():
    """SURVEY.md 9.2 anchors (computed by the reference with glibc log)."""
    a = util.compiled_for(["l4c4"], dict(length=4), True).arrays()
    assert a["k"] == 2 and len(a["emit_src"]) == 686 and len(a["null_src"]) == 27
    assert a["noGap"] == -0.0020020026706730793
    assert a["delOpen"] == a["tanDup"] == -6.9077552789821368
    assert a["delExtend"] == -4.6051701859880909
    assert a["delEnd"] == -0.010050335853501451
    sub = a["sub"].reshape(4, 4)
    assert sub[0, 0] == 1.3762440252663892 and sub[0, 2] == -3.3141860046725249 and sub[0, 1] == -6.3099182782265162
    assert a["len"][0] == -0.69314718055994529
    scores = {chr(c): s for c, s in zip(a["emit_in"], a["emit_score"])}
    assert scores["0"] == scores["1"] == -1.3863019904853182
    assert scores["A"] == -12.476656879444443
    a8 = util.compiled_for(["l4c4"], {}, False).arrays()  # default -l 12: k = min(4, 6)
    assert a8["k"] == 4 and a8["local"] == 1 and a8["len"][0] == -1.791759469228055
    w = util.compiled_for(["l4c4", "water64.1"], dict(length=4), True).arrays()
    assert set(np.unique(w["emit_score"])) <= {0.0, -1.3862943611198906}


def test_tables_edge_order_is_reference_list_order():
    """incoming lists are ordered by (source index, transition index) (src/viterbi.cpp:30-58)."""
    a = util.compiled_for(["l4c4", "sync16", "flusher", "hamming74"], dict(length=4), True).arrays()
    for off, src in ((a["emit_off"], a["emit_src"]), (a["null_off"], a["null_src"])):
        for s in range(a["n_states"]):
            seg = src[off[s]:off[s + 1]]
            assert (np.diff(seg.astype(np.int64)) >= 0).all()
    deg = np.diff(a["null_off"])
    assert deg.max() == 98  # the hub states of s16h74l4c4 (SURVEY.md 7 hard part 5)


def test_null_cycle_is_rejected():
    text = ('{"state":[{"n":0,"trans":[{"to":1}]},{"n":1,"trans":[{"to":0},{"in":"$","out":"A","to":2}]},'
            '{"n":2,"trans":[]}]}')
    with pytest.raises(d.DnabError, match="cyclic"):
        d.Machine.from_json(text).compile(d.ErrorFlags(length=4))


def test_pack_reads_layout_and_errors():
    packed, off, ln = d.pack_reads(["ACGT", "ttgca", ""])
    assert ln.tolist() == [4, 5, 0] and off.tolist() == [0, 16, 32]
    assert packed[0] == 0b11100100  # A=0 C=1 G=2 T=3, base i in bits 2*(i%4)
    assert packed[16] == (3 | 3 << 2 | 2 << 4 | 1 << 6) and packed[17] == 0
    with pytest.raises(d.DnabError, match="Unknown symbol N"):
        d.pack_reads(["ACGN"])


def test_decoder_needs_a_gpu_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(d.DnabError, match="no CPU fallback"):
        d.Decoder(util.compiled_for(["l4c4"], dict(length=4), True))


def test_divide_by_table_step_is_ieee_division(tmp_path):
    """csrc/lse_table.cuh replaces the two `/ .0001` of the reference's log_sum_exp_unary
    (src/logsumexp.h:61-72) by q = x*10000; r = fma(-1e-4, q, x); q + r*10000 (one more fma).  The same
    three operations on the CPU must equal hardware division for every operand: random in [0,10), random
    exponents of both signs, and +-200 ulps around every multiple of the step (where the truncation to a
    table index is decided)."""
    import subprocess
    src = tmp_path / "divstep.c"
    src.write_text(r'''
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static double div3(double x) { const double c = .0001, y = 10000.0; double q = x * y; double r = fma(-c, q, x); return fma(r, y, q); }
static uint64_t s = 88172645463325252ULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void) {
  const double c = .0001; long bad = 0, tot = 0;
  for (long i = 0; i < 20000000L; i++) { double x = (rnd() >> 11) * (10.0 / 9007199254740992.0); bad += (x / c != div3(x)); tot++; }
  for (long i = 0; i < 20000000L; i++) {
    uint64_t bits = ((uint64_t)(1023 - 80 + (int)(rnd() % 85)) << 52) | (rnd() & 0xFFFFFFFFFFFFFULL) | ((rnd() & 1) << 63);
    double x; memcpy(&x, &bits, 8); bad += (x / c != div3(x)); tot++; }
  for (long n = 0; n <= 100000; n++) { double x0 = n * c; uint64_t b0; memcpy(&b0, &x0, 8);
    for (int d = -200; d <= 200; d++) { if (n == 0 && d < 0) continue; uint64_t bb = b0 + d; double x; memcpy(&x, &bb, 8); bad += (x / c != div3(x)); tot++; } }
  printf("%ld %ld\n", tot, bad); return 0; }
''')
    exe = tmp_path / "divstep"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(src), "-lm"], check=True)
    tot, bad = map(int, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
    assert tot > 8e7 and bad == 0


def test_fasta_ingest_forms_through_the_exact_decoder(tmp_path):
    """readFastSeqs (reference src/fastseq.cpp:123-148, kseq): plain or gzip, FASTA or FASTQ, multi-line records
    concatenated, name = header up to the first blank, lower case accepted downstream.  Exercised on the CPU through
    dnab_exact_decode_fasta (-d), which reads with the same ingest code as the Viterbi path."""
    case = [c for c in json.load(open(os.path.join(util.GOLDEN, "exact_golden.json")))["cases"] if c["kind"] == "file"][0]
    m = util.machine_from_recipe(tuple(case["recipe"]))
    dna = case["records"][0]
    forms = {
        "plain.fa": f">hello some comment\n{dna}\n",
        "multiline.fa": ">hello\n" + "\n".join(dna[i:i + 7] for i in range(0, len(dna), 7)) + "\n",
        "lower.fa": f">hello\n{dna.lower()}\n",
        "reads.fq": f"@hello\n{dna}\n+\n{'I' * len(dna)}\n",
        "two_records.fa": f">a\n{dna[:20]}\n>b\n{dna[20:]}\n",  # one decoder across records (t/dnastore.cpp:187-190)
    }
    for name, text in forms.items():
        p = tmp_path / name
        p.write_text(text)
        data, warnings = d.exact_decode_fasta(m, p)
        assert data == b"HELLO", name
        assert warnings == case["warnings"], name
    gz = tmp_path / "plain.fa.gz"
    with gzip.open(gz, "wt") as f:
        f.write(forms["multiline.fa"])
    assert d.exact_decode_fasta(m, gz)[0] == b"HELLO"
    with pytest.raises(d.DnabError):
        d.exact_decode_fasta(m, tmp_path / "missing.fa")


def test_every_option_key_is_documented_in_the_header():
    """dnab_decoder_set_option / dnab_multi_decoder_set_option: a key the library accepts is a key the ABI header explains."""
    import re
    root = util.ROOT
    header = open(os.path.join(root, "include", "dnastore_b200.h")).read()
    dec = open(os.path.join(root, "dnastore_b200", "csrc", "decoder.cu")).read()
    body = dec[dec.index("int dnab_decoder_set_option("):]
    body = body[:body.index("\n}\n")]
    keys = re.findall(r'k == "([a-z_0-9]+)"', body)
    pipe = open(os.path.join(root, "dnastore_b200", "csrc", "pipeline.cpp")).read()
    keys += re.findall(r'== "([a-z_0-9]+)"', pipe[pipe.index("dnab_multi_decoder_set_option("):])
    assert len(keys) >= 20
    missing = [k for k in keys if f'"{k}"' not in header]
    assert not missing, missing


def test_profile_files_named_in_the_docs_exist():
    """profiles/README.md and DESIGN.md cite ncu summaries by file name: every cited file is committed, and every
    round-2 summary is described in the README."""
    root = util.ROOT
    have = set(os.listdir(os.path.join(root, "profiles")))
    readme = open(os.path.join(root, "profiles", "README.md")).read()
    design = open(os.path.join(root, "DESIGN.md")).read()
    cited = set(re.findall(r"(r0[12]_[a-z0-9_]+\.(?:txt|json|csv))", readme + design))
    assert cited and not (cited - have), sorted(cited - have)
    undocumented = {f for f in have if f.startswith("r02_") and f not in readme}
    assert not undocumented, sorted(undocumented)
