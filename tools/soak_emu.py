#!/usr/bin/env python
"""Randomised differential soak of the DEVICE code under the SIMT emulator (tests/emu) against the oracle.

Random pairs (lengths 0..256 or up to 600, overlaps, adapters, N runs, foreign bytes, homopolymers, low-quality stretches,
power-of-two lengths, qualities outside the 6-bit transport range) x random parameter sets x every filter kernel x every statistics
kernel path, with the in-place qual2 column switched on at random.  Test
infrastructure; needs no GPU.

    python tools/soak_emu.py --cases 200 --seed 1
"""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import cases  # noqa: E402
import compare  # noqa: E402
import emu  # noqa: E402
from afterqc_b200 import _abi  # noqa: E402
from afterqc_b200.batch import PackedBatch  # noqa: E402
from oracle import oracle  # noqa: E402


def rand_qual(rng, n, style):
    if style == 0:
        return "".join(chr(33 + rng.randint(2, 40)) for _ in range(n))
    if style == 1:
        return "".join(rng.choice("#$%5?ACEFGHI") for _ in range(n))
    if style == 2:
        return "".join(rng.choice("#I") for _ in range(n))
    if style == 4:        # qualities above Phred 63 and below '!': odd bytes in the quality columns
        return "".join(rng.choice("I5~}a{ \"") if rng.random() < 0.3 else "F" for _ in range(n))
    return rng.choice("#/05I") * n


def make_pair(rng, maxlen, alphabet):
    L1 = rng.choice([rng.randint(0, maxlen), rng.randint(30, min(maxlen, 160)), rng.choice([31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 150, 151, 159, 160, 161, 255, 256])])
    L2 = rng.choice([L1, L1, rng.randint(0, maxlen), max(0, L1 - rng.randint(0, 9))])
    L1, L2 = min(L1, maxlen), min(L2, maxlen)
    kind = rng.random()
    if kind < 0.15:      # unrelated
        r1 = cases._rand_seq(rng, L1, alphabet); r2 = cases._rand_seq(rng, L2, alphabet)
    else:
        frag = rng.randint(1, L1 + L2 + 20)
        f = cases._rand_seq(rng, frag, alphabet)
        if rng.random() < 0.1:      # low complexity fragment
            unit = cases._rand_seq(rng, rng.randint(1, 3), "ACGT")
            f = (unit * (frag // len(unit) + 1))[:frag]
        r1 = (f + cases._rand_seq(rng, L1, alphabet))[:L1]
        r2 = (cases.revcomp(f) + cases._rand_seq(rng, L2, alphabet))[:L2]
        r2 = cases._mutate(rng, r2, rng.choice([0, 0, 0, 1, 1, 2, 3, 4, 6]), "ACGTN") if r2 else r2
        if rng.random() < 0.2 and r1:
            r1 = cases._mutate(rng, r1, rng.randint(1, 3), "ACGTN")
    if rng.random() < 0.08 and L1 > 20:      # homopolymer stretch
        b = rng.choice("ACGTN"); st = rng.randint(0, L1 - 10); run = rng.randint(8, 50)
        r1 = (r1[:st] + b * run + r1[st + run:])[:L1]
        if rng.random() < 0.5 and len(r1) > st + 5:
            p = st + rng.randint(0, min(run, len(r1) - st) - 1)
            r1 = r1[:p] + rng.choice("ACGT") + r1[p + 1:]
    if rng.random() < 0.05 and L2 > 20:
        b = rng.choice("ACGTN"); st = rng.randint(0, L2 - 10); run = rng.randint(8, 50)
        r2 = (r2[:st] + b * run + r2[st + run:])[:L2]
    s1, s2 = rng.randint(0, 4), rng.randint(0, 4)
    return (r1, rand_qual(rng, len(r1), s1)), (r2, rand_qual(rng, len(r2), s2))


def rand_params(rng, paired=True):
    kw = {}
    if rng.random() < 0.4:
        kw.update(trim_front=rng.randint(0, 12), trim_tail=rng.randint(0, 12), trim_front2=rng.randint(0, 12), trim_tail2=rng.randint(0, 12))
    if rng.random() < 0.5:
        kw.update(seq_len_req=rng.choice([0, 5, 20, 35, 60, 100]))
    if rng.random() < 0.4:
        kw.update(poly_size_limit=rng.choice([0, 8, 12, 20, 35, 50]), allow_mismatch_in_poly=rng.choice([0, 1, 2, 5]))
    if rng.random() < 0.4:
        kw.update(qualified_quality_phred=rng.choice([0, 2, 15, 20, 30, 41]), unqualified_base_limit=rng.choice([0, 1, 10, 30, 60, 200]))
    if rng.random() < 0.4:
        kw.update(n_base_limit=rng.choice([0, 1, 3, 5, 50]))
    if rng.random() < 0.15:
        kw.update(no_overlap=1)
    if rng.random() < 0.25:
        kw.update(no_correction=1)
    if rng.random() < 0.25:
        kw.update(mask_mismatch=1)
    kw.update(qc_sample=rng.choice([0, 50, 200000]), qc_kmer=rng.choice([3, 5, 8]))
    return _abi.Params.defaults(paired=1 if paired else 0, **kw)


def one_case(rng, k):
    maxlen = rng.choice([160, 160, 256, 256, 128, 600])
    alphabet = rng.choice(["ACGT", "ACGT", "ACGTN", "ACGTACGTACGTN", "ACGTNacgtRY-"])
    n = rng.randint(1, 260)
    pairs = [make_pair(rng, maxlen, alphabet) for _ in range(n)]
    pairs = [(a, b) for a, b in pairs]
    paired = rng.random() < 0.85
    # statRead raises below 5 bases (both sides flag the same sticky error); keep such reads out of the sampled window
    batch = PackedBatch.from_reads([p[0] for p in pairs], [p[1] for p in pairs] if paired else None, first_index=rng.choice([0, 0, 17, 199990]))
    p = rand_params(rng, paired)
    results = {}
    # (filter kernel, statistics kernel): the round-1 path, the shipped default, and the two mixed forms
    for kern, sk in ((_abi.KERNEL_WARP, _abi.STAT_WARP), (_abi.KERNEL_DEFAULT, _abi.STAT_DEFAULT), (_abi.KERNEL_WARP, _abi.STAT_DEFAULT), (_abi.KERNEL_LANE, _abi.STAT_WARP)):
        p.filter_kernel = kern
        p.stat_kernel = sk
        orc, eng = oracle.Oracle(p), emu.EmuEngine(p)
        try:
            err_o = err_e = None
            try:
                a = orc.filter_pairs(batch); ca = orc.counters()
            except Exception as e:      # noqa: BLE001
                err_o = getattr(e, "code", repr(e))
            try:
                b = eng.filter_pairs(batch, qual2_in_place=rng.random() < 0.5)
                cb = eng.counters()
            except Exception as e:      # noqa: BLE001
                err_e = getattr(e, "code", repr(e))
            if err_o is not None or err_e is not None:
                assert err_o == err_e, "case %d kernel %d/%d: oracle error %r engine error %r" % (k, kern, sk, err_o, err_e)
                results[(kern, sk)] = "both raise %r" % (err_o,)
                continue
            what = "case %d kernel %d stat_kernel %d" % (k, kern, sk)
            compare.assert_records_equal(batch, a, b, what)
            slots = (_abi.QC_R1_POST, _abi.QC_R2_POST) if paired else (_abi.QC_R1_POST,)
            compare.compare_backends(orc, eng, slots, what)
            results[(kern, sk)] = "ok"
            if kern == _abi.KERNEL_WARP:        # the operator and prefilter-statistics entries (pair_kernel; stat_kernel with stat_kernel = 0)
                compare.assert_records_equal(batch, orc.ops_pairs(batch), eng.ops_pairs(batch), what + " ops")
                lo = batch.first_index + rng.randint(0, max(0, batch.n - 1)); hi = lo + rng.randint(0, batch.n)
                errs = []
                ob = rng.choice([0, 1 << 30])
                for be in (orc, eng):
                    try:
                        be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE if paired else -1, stat_lo=lo, stat_hi=hi, order_base=ob)
                        be.counters()
                        errs.append(None)
                    except Exception as e:      # noqa: BLE001
                        errs.append(getattr(e, "code", repr(e)))
                if errs[0] is None and errs[1] is None:
                    compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE) if paired else (_abi.QC_R1_PRE,), what + " stat")
                else:
                    assert errs[0] == errs[1], "%s stat: oracle error %r engine error %r" % (what, errs[0], errs[1])
        finally:
            orc.close(); eng.close()
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    oracle.build()
    rng = random.Random(args.seed)
    tally = {}
    for k in range(args.cases):
        r = one_case(rng, k)
        for v in r.values():
            tally[v] = tally.get(v, 0) + 1
        if (k + 1) % 20 == 0:
            print("%d cases: %s" % (k + 1, tally), flush=True)
    print("soak done: %d cases, %s" % (args.cases, tally))


if __name__ == "__main__":
    main()
