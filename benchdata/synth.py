"""Synthetic read generation for tests and bench.py (inputs only -- no decoding here).

"Mutated with the reference error model" means the reference's only mutation
simulator, doc/errdecode.pl:223-225,261-317, restated here:

  order: duplications, then substitutions, then deletions (errdecode.pl:223-225);
  each kind makes round(rate * current_length) events (evolve, :261-285);
  an event picks size uniform in [1, min(len, maxsize)] and a uniform start
  (randcoords, :287-293); a duplication inserts a tandem copy of the segment and is
  skipped when it overlaps an earlier mutation (the simulator marks mutated bases in
  upper case; allowoverlaps is false for dups, true for subs and dels);
  a substitution is a transversion with probability 1/(1+iv) (uniform over the two),
  else the transition A<->G, C<->T (subst, :307-317); a deletion removes the segment.

Encoded payloads come from the reference encoder (a pool generated once by
benchdata/make_pools.py with oracle/_ref/refdriver and committed), so nothing here
needs /root/reference at run time.
"""
import gzip
import os

import numpy as np

_TRANSITION = {"a": "G", "g": "A", "c": "T", "t": "C"}
_TRANSVERSION = {"a": "CT", "g": "CT", "c": "AG", "t": "AG"}

HERE = os.path.dirname(os.path.abspath(__file__))


def random_bits(rng, n):
    return "".join("01"[b] for b in rng.integers(0, 2, size=n))


def _round_half_up(x):
    return int(np.floor(x + 0.5))


def mutate(seq, rng, sub_rate=0.0, iv_ratio=10.0, dup_rate=0.0, max_dup=1, del_rate=0.0, max_del=1):
    """One read through the errdecode.pl simulator. `rng` is a numpy Generator."""
    s = list(seq.lower())

    def coords(maxsize):
        n = len(s)
        size = int(rng.random() * (min(n, maxsize) + 1 - 1)) + 1
        pos = int(rng.random() * (n + 1 - size))
        return pos, size

    for _ in range(_round_half_up(dup_rate * len(s))):
        pos, size = coords(max_dup)
        if any(c.isupper() for c in s[pos:pos + size]):
            continue
        s[pos:pos + size] = s[pos:pos + size] + [c.upper() for c in s[pos:pos + size]]
    for _ in range(_round_half_up(sub_rate * len(s))):
        pos, _size = coords(1)
        base = s[pos].lower()
        if rng.random() < 1.0 / (1.0 + iv_ratio):
            s[pos] = _TRANSVERSION[base][int(rng.random() * 2)]
        else:
            s[pos] = _TRANSITION[base]
    for _ in range(_round_half_up(del_rate * len(s))):
        pos, size = coords(max_del)
        del s[pos:pos + size]
    return "".join(s).upper()


def mutate_aligned(seq, rng, sub_rate=0.0, iv_ratio=10.0, dup_rate=0.0, max_dup=1, del_rate=0.0, max_del=1):
    """Same simulator as mutate(), but keeps the pairwise alignment (errdecode.pl tracks it in
    @origpos, :271-273, and prints it as Stockholm, :320-340): returns (original row, observed row),
    gapped with '-' and of equal length."""
    cols = [[c.upper(), c.lower()] for c in seq]  # [original base or '-', observed base (lower = untouched)]

    def observed():
        return [i for i, c in enumerate(cols) if c[1] != "-"]

    def coords(maxsize):
        obs = observed()
        n = len(obs)
        size = int(rng.random() * (min(n, maxsize) + 1 - 1)) + 1
        pos = int(rng.random() * (n + 1 - size))
        return obs[pos:pos + size]

    n0 = len(observed())
    for _ in range(_round_half_up(dup_rate * n0)):
        seg = coords(max_dup)
        if not seg or any(cols[i][1].isupper() for i in seg):
            continue
        ins = [["-", cols[i][1].upper()] for i in seg]
        at = seg[-1] + 1
        cols[at:at] = ins
    for _ in range(_round_half_up(sub_rate * len(observed()))):
        seg = coords(1)
        if not seg:
            continue
        i = seg[0]
        base = cols[i][1].lower()
        cols[i][1] = (_TRANSVERSION[base][int(rng.random() * 2)] if rng.random() < 1.0 / (1.0 + iv_ratio)
                      else _TRANSITION[base])
    for _ in range(_round_half_up(del_rate * len(observed()))):
        for i in coords(max_del):
            cols[i][1] = "-"
    cols = [c for c in cols if not (c[0] == "-" and c[1] == "-")]
    return "".join(c[0] for c in cols), "".join(c[1].upper() for c in cols)


def mutate_subs_batch(seqs, rng, sub_rate=0.01, iv_ratio=10.0):
    """Vectorised substitution-only mutation of many reads (BASELINE config 2 / 5):
    round(sub_rate*len) substitutions per read at uniform positions (with replacement,
    as the simulator allows overlapping substitutions)."""
    lens = np.array([len(s) for s in seqs])
    L = int(lens.max())
    arr = np.full((len(seqs), L), 255, dtype=np.uint8)
    lut = np.zeros(256, dtype=np.uint8)
    for i, ch in enumerate("ACGT"):
        lut[ord(ch)] = i
        lut[ord(ch.lower())] = i
    for r, s in enumerate(seqs):
        arr[r, :len(s)] = lut[np.frombuffer(s.encode(), dtype=np.uint8)]
    nsub = np.floor(sub_rate * lens + 0.5).astype(int)
    for j in range(int(nsub.max())):
        rows = np.nonzero(nsub > j)[0]
        pos = (rng.random(len(rows)) * lens[rows]).astype(int)
        base = arr[rows, pos]
        tv = rng.random(len(rows)) < 1.0 / (1.0 + iv_ratio)
        pick = (rng.random(len(rows)) * 2).astype(np.uint8)
        # codes A,C,G,T = 0..3: transition flips bit 1 (A<->G, C<->T); transversion toggles bit 0 and picks bit 1
        transition = base ^ 2
        transversion = ((base & 1) ^ 1) | (pick << 1)
        arr[rows, pos] = np.where(tv, transversion, transition)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    return [letters[arr[r, :lens[r]]].tobytes().decode() for r in range(len(seqs))]


def load_pool(name):
    """Reference-encoded DNA strings, one per line (benchdata/pools/<name>.txt.gz)."""
    with gzip.open(os.path.join(HERE, "pools", name + ".txt.gz"), "rt") as f:
        return [ln.strip() for ln in f if ln.strip()]
