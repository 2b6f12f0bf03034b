"""Seeded inputs shared by the parity tests: synthetic configs + adversarial hand-made reads."""
import random

import numpy as np

from afterqc_b200 import _abi, synth
from afterqc_b200.batch import PackedBatch

COMP = {"A": "T", "T": "A", "C": "G", "G": "C", "a": "t", "t": "a", "c": "g", "g": "c", "N": "N"}


def revcomp(s):
    return "".join(COMP.get(c, "N") for c in reversed(s))


def _rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def _rand_qual(rng, n, lo=2, hi=40):
    return "".join(chr(33 + rng.randint(lo, hi)) for _ in range(n))


def _mutate(rng, s, k, alphabet="ACGT"):
    s = list(s)
    for p in rng.sample(range(len(s)), min(k, len(s))):
        s[p] = rng.choice([c for c in alphabet if c != s[p]])
    return "".join(s)


def adversarial_pairs(seed=7, n_random=400):
    """Pairs built to sit on the decision boundaries of overlap_hm / hasPolyX / the filters."""
    rng = random.Random(seed)
    pairs = []

    def add(r1, r2, q1=None, q2=None):
        q1 = q1 if q1 is not None else _rand_qual(rng, len(r1))
        q2 = q2 if q2 is not None else _rand_qual(rng, len(r2))
        pairs.append(((r1, q1), (r2, q2)))

    # the reference's own self-check pairs (util.py:242-246)
    add("CAGCGCCTACGGGCCCCTTTTTCTGCGCGACCGCGTGGCTGTGGGCGCGGATGCCTTTGAGCGCGGTGACTTCTCACTGCGTATCGAGCCGCTGGAGGTCTCCC",
        "ACCTCCAGCGGCTCGATACGCAGTGAGAAGTCACCGCGCTCAAAGGCATCCGCGCCCACAGCCACGCGGTCGCGCAGAAAAAGGGGCCCGTAGGCGCGGCTCCC")
    add("CAGCGCCTACGGGCCCCTTTTTCTGCGCGACCGCGTGGCTGTGGGCGCGGATGCCTTTGAGCGCGGTGACTTCTCACTGCGTATCGAGC",
        "ACCTCCAGCGGCTCGATACGCAGTGAGAAGTCACCGCGCTCAAAGGCATCCGCGCCCACAGCCACGCGGTCGCGCAGAAAAAGGGGTCC")
    # overlaps of every length around the 30/31/50/51/52 boundaries, 0..5 mismatches placed before/after position 50
    for frag in list(range(28, 70)) + [80, 100, 149, 150, 151, 200, 260, 299, 300]:
        for L1, L2 in ((150, 150), (151, 120), (100, 150), (75, 75), (36, 150), (150, 31)):
            f = _rand_seq(rng, max(frag, 1))
            ad1, ad2 = _rand_seq(rng, 200), _rand_seq(rng, 200)
            r1 = (f + ad1)[:L1]
            r2 = (revcomp(f) + ad2)[:L2]
            k = rng.choice([0, 0, 1, 2, 3, 4, 5])
            where = rng.choice(["any", "head", "tail"])
            r2l = list(r2)
            for _ in range(k):
                lim = min(len(r2l), frag)
                if lim <= 0:
                    break
                if where == "head":
                    p = rng.randrange(0, min(lim, 50))
                elif where == "tail":
                    p = rng.randrange(min(lim - 1, 50), lim)
                else:
                    p = rng.randrange(0, lim)
                r2l[p] = rng.choice([c for c in "ACGT" if c != r2l[p]])
            r2 = "".join(r2l)
            # quality patterns that trigger both correction directions / skips
            q1 = "".join(rng.choice("#$%5?ACEFGHI") for _ in range(len(r1)))
            q2 = "".join(rng.choice("#$%5?ACEFGHI") for _ in range(len(r2)))
            add(r1, r2, q1, q2)
    # N runs, lowercase and foreign characters inside overlaps
    for _ in range(120):
        frag = rng.randint(40, 220)
        f = _rand_seq(rng, frag, "ACGTACGTACGTNacgtRY-")
        r1 = (f + _rand_seq(rng, 160))[:rng.randint(60, 150)]
        r2 = (revcomp(_rand_seq(rng, 0) + f) + _rand_seq(rng, 160))[:rng.randint(60, 150)]
        r2 = _mutate(rng, r2, rng.randint(0, 3), "ACGTN")
        add(r1, r2)
    # polyX: runs of 30..40 with 0..4 interruptions, at the head/middle/tail, incl. N, lowercase, foreign aborts
    for _ in range(200):
        L = rng.randint(34, 150)
        base = rng.choice("ACGTNacgt")
        run = rng.randint(28, 45)
        s = list(_rand_seq(rng, L))
        start = rng.randint(0, max(0, L - run))
        for i in range(start, min(L, start + run)):
            s[i] = base
        for _k in range(rng.randint(0, 4)):
            p = rng.randint(start, min(L - 1, start + run - 1))
            s[p] = rng.choice("ACGT")
        if rng.random() < 0.3:
            s[rng.randrange(L)] = rng.choice("RY.-n")
        r1 = "".join(s)
        r2 = _rand_seq(rng, rng.randint(34, 150))
        if rng.random() < 0.5:
            r1, r2 = r2, r1
        add(r1, r2)
    # low-quality and N-count boundaries (60/61 low quals, 5/6 N)
    for nlow in (59, 60, 61, 62):
        for nn in (4, 5, 6):
            L = 150
            s = list(_rand_seq(rng, L)); q = ["I"] * L
            for p in rng.sample(range(L), nlow):
                q[p] = rng.choice("!#$%&/")     # < Q15
            for p in rng.sample(range(L), nn):
                s[p] = "N"
            r2 = list(_rand_seq(rng, L))
            for p in rng.sample(range(L), rng.choice([0, 5, 6])):
                r2[p] = "N"
            add("".join(s), "".join(r2), "".join(q), _rand_qual(rng, L))
    # short / ragged reads
    for L1, L2 in ((5, 5), (5, 150), (150, 5), (30, 30), (31, 31), (32, 33), (34, 35), (35, 34), (64, 64), (65, 63), (96, 97), (128, 129)):
        f = _rand_seq(rng, 300)
        add(f[:L1], revcomp(f[:max(L1, L2)])[:L2])
    # plain random pairs
    for _ in range(n_random):
        add(_rand_seq(rng, rng.randint(5, 160)), _rand_seq(rng, rng.randint(5, 160)))
    return pairs


def adversarial_batch(seed=7, first_index=0):
    pairs = adversarial_pairs(seed)
    return PackedBatch.from_reads([p[0] for p in pairs], [p[1] for p in pairs], first_index=first_index)


def homopolymer_batch(n, L=150, seed=3):
    """reads that pile onto very few k-mers: the first tenth all A (mate 2 all T), then all C / all G, then an ACAC.. repeat,
    each with a sprinkle of other bases -- a CTA's 16-bit shared-memory counter of one k-mer passes 0x4000 many times, for
    k-mers in the low and the high half of a counter word, met inside and (C, G, the repeat) beyond the stamped head"""
    rng = random.Random(seed)
    r1s, r2s = [], []
    for i in range(n):
        base1, base2 = ("A", "T") if i < n // 10 else (("C", "G") if i < (6 * n) // 10 else ("AC", "GT"))
        s1 = list((base1 * L)[:L]); s2 = list((base2 * L)[:L])
        for s in (s1, s2):
            if rng.random() < 0.3:
                s[rng.randrange(L)] = rng.choice("ACGTN")
        r1s.append(("".join(s1), _rand_qual(rng, L))); r2s.append(("".join(s2), _rand_qual(rng, L)))
    return PackedBatch.from_reads(r1s, r2s)


def long_read_batch(seed=11, n=300, lo=200, hi=1000):
    rng = random.Random(seed)
    r1s, r2s = [], []
    for _ in range(n):
        frag = rng.randint(lo, hi + 200)
        f = _rand_seq(rng, frag, "ACGTACGTACGTACGTN")
        L1, L2 = rng.randint(lo, hi), rng.randint(lo, hi)
        r1 = (f + _rand_seq(rng, hi))[:L1]
        r2 = _mutate(rng, (revcomp(f) + _rand_seq(rng, hi))[:L2], rng.randint(0, 4))
        r1s.append((r1, _rand_qual(rng, L1))); r2s.append((r2, _rand_qual(rng, L2)))
    return PackedBatch.from_reads(r1s, r2s)


PARAM_SETS = {
    "default_f0": dict(),
    "trim": dict(trim_front=3, trim_tail=5, trim_front2=2, trim_tail2=7),
    "trim_front_only": dict(trim_front=10, trim_tail=0, trim_front2=0, trim_tail2=4),
    "mask": dict(mask_mismatch=1),
    "nocorr": dict(no_correction=1),
    "nocorr_mask": dict(no_correction=1, mask_mismatch=1),
    "no_overlap": dict(no_overlap=1),
    "loose": dict(seq_len_req=0, poly_size_limit=0, unqualified_base_limit=0, n_base_limit=0, qc_sample=0),
    "strict": dict(seq_len_req=60, poly_size_limit=20, allow_mismatch_in_poly=1, qualified_quality_phred=20,
                   unqualified_base_limit=30, n_base_limit=1, qc_kmer=5, qc_sample=0),
    "poly_wide": dict(poly_size_limit=50, allow_mismatch_in_poly=5, qc_kmer=3, qc_sample=300),
}


def make_params(name, paired=True):
    kw = dict(PARAM_SETS[name])
    return _abi.Params.defaults(paired=1 if paired else 0, **kw)


def synthetic(config, n, **kw):
    return synth.generate(config, n, **kw)
