"""CPU: the exact (error-free) decoder behind -d / --decode-string / --decode-bits (reference src/decoder.h:7-240,
t/dnastore.cpp:185-211) against goldens produced by RUNNING the reference (tests/golden/make_golden_exact.py):
its own testdecode known answers, reference-encoded random payloads on every machine family, every prefix of an
encoding (the "Decoder unresolved" warnings) and undecodable strings."""
import json
import os
import subprocess

import pytest

import dnastore_b200 as d
import dnab_testutil as util

CASES = json.load(open(os.path.join(util.GOLDEN, "exact_golden.json")))["cases"]
CLI = os.path.join(util.ROOT, "bin", "dnastore-b200")


def by_kind(kind):
    return [c for c in CASES if c["kind"] == kind]


def test_golden_file_has_every_kind():
    assert len(by_kind("file")) == 5 and len(by_kind("string")) >= 20 and len(by_kind("prefix")) >= 30 and len(by_kind("error")) == 3


@pytest.mark.parametrize("case", by_kind("file"), ids=lambda c: "+".join(c["recipe"]))
def test_decode_file_known_answers(case, tmp_path):
    """reference Makefile:142,153,168,176,183 -- every record through ONE decoder, BinaryWriter bytes == data/hello.txt."""
    m = util.machine_from_recipe(tuple(case["recipe"]))
    fa = tmp_path / "in.fa"
    fa.write_text("".join(f">r{i}\n{s}\n" for i, s in enumerate(case["records"])))
    data, warnings = d.exact_decode_fasta(m, fa)
    assert data == bytes.fromhex(case["bytes_hex"]) == b"HELLO"
    assert warnings == case["warnings"]
    # --decode-bits on the concatenated records (Makefile:144)
    dec = d.ExactDecoder(m).feed("".join(case["records"])).close()
    assert dec.take_symbols() == case["symbols"]
    assert dec.warnings == case["symbol_warnings"]


@pytest.mark.parametrize("case", by_kind("string"), ids=lambda c: "+".join(c["recipe"]) + ":" + c["payload"][:12])
def test_reference_encoded_payloads_decode_to_reference_output(case):
    m = util.machine_from_recipe(tuple(case["recipe"]))
    dec = d.ExactDecoder(m).feed(case["dna"]).close()
    sym = dec.take_symbols()
    assert sym == case["symbols"]
    assert dec.warnings == case["symbol_warnings"]
    data, left, warn = d.pack_decoded_symbols(sym)
    assert data == bytes.fromhex(case["bytes_hex"])
    assert dec.warnings + warn == case["warnings"]
    # lower-case input is accepted (decodeString upper-cases, decoder.h:186-189); feeding base by base is the same
    dec2 = d.ExactDecoder(m)
    for c in case["dna"].lower():
        dec2.feed(c)
    assert dec2.close().take_symbols() == sym


@pytest.mark.parametrize("case", by_kind("prefix"), ids=lambda c: str(len(c["dna"])))
def test_truncated_encodings_release_what_the_reference_releases(case):
    m = util.machine_from_recipe(tuple(case["recipe"]))
    dec = d.ExactDecoder(m).feed(case["dna"]).close()
    assert dec.take_symbols() == case["symbols"]
    assert dec.warnings == case["symbol_warnings"]
    assert dec.hypotheses == 0  # close() clears the hypothesis set (decoder.h:45)


@pytest.mark.parametrize("case", by_kind("error"), ids=lambda c: c["dna"][:12])
def test_undecodable_input_reports_the_reference_assertion(case):
    m = util.machine_from_recipe(tuple(case["recipe"]))
    with pytest.raises(d.DnabError) as e:
        d.ExactDecoder(m).feed(case["dna"])
    assert case["message"].replace("Abort: ", "") in str(e.value)
    assert case["ref_returncode"] != 0


def test_symbols_are_released_incrementally():
    """take_symbols() drains what is certain so far; the concatenation equals the one-shot result."""
    case = by_kind("string")[0]
    m = util.machine_from_recipe(tuple(case["recipe"]))
    dec, parts = d.ExactDecoder(m), []
    for i in range(0, len(case["dna"]), 7):
        dec.feed(case["dna"][i:i + 7])
        parts.append(dec.take_symbols())
    dec.close()
    parts.append(dec.take_symbols())
    assert "".join(parts) == case["symbols"] and any(parts[:-1])


def test_binary_writer_bit_order_and_ignored_symbols():
    """decoder.h:213-239: first bit = least significant; ^ and $ skipped silently, control symbols with a warning."""
    data, left, warn = d.pack_decoded_symbols("^10000000" + "01000000" + "A" + "111$")
    assert data == bytes([1, 2]) and left == "111"
    assert warn == ["Ignoring control character #0 ('A') in decoder", "3 bits (111) remaining on output"]
    data, left, warn = d.pack_decoded_symbols("1")
    assert data == b"" and left == "1" and warn == ["1 bit (1) remaining on output"]
    data, left, warn = d.pack_decoded_symbols("0001" + "0010")  # 'H' = 0x48
    assert data == b"H" and left == "" and warn == []


def test_cli_exact_decode_flags(tmp_path):
    """-d / -D / -B of bin/dnastore-b200 print what bin/dnastore prints (t/dnastore.cpp:185-211)."""
    case = by_kind("file")[0]
    mj = tmp_path / "m.json"
    mj.write_text(util.machine_from_recipe(tuple(case["recipe"])).to_json())
    fa = tmp_path / "in.fa"
    fa.write_text(">hello\n" + case["records"][0] + "\n")
    p = subprocess.run([CLI, "-v0", "-L", str(mj), "-d", str(fa)], capture_output=True)
    assert p.returncode == 0 and p.stdout == b"HELLO" and p.stderr.decode().split("\n")[0] == "Warning: " + case["warnings"][0]
    p = subprocess.run([CLI, "-v0", "--load-machine", str(mj), "--decode-string", case["records"][0]], capture_output=True)
    assert p.returncode == 0 and p.stdout == b"HELLO"
    p = subprocess.run([CLI, "-v0", "--load-machine", str(mj), "--decode-bits", case["records"][0]], capture_output=True)
    assert p.returncode == 0 and p.stdout.decode() == case["symbols"] + "\n"
    p = subprocess.run([CLI, "-v0", "--load-machine", str(mj), "-B", "AAAAAAAA"], capture_output=True)
    assert p.returncode != 0 and b"Can't decode 'A'" in p.stderr
