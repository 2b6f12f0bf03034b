"""Levenshtein distance on the GPU (aqc_edit_distance_batch, csrc/aqc_edit.cuh: Myers' bit-vector algorithm, one lane per pair)
and the libed.so-compatible entry points, against the oracle's dynamic programme -- which is itself pinned against the
reference's own editdistance/_editdistance.cpp (compiled as oracle/_ref/libed_ref.so in the build container)."""
import ctypes as C
import random

import pytest

from oracle import oracle


def _pairs(seed, n):
    rng = random.Random(seed)
    out = [("", ""), ("A", ""), ("", "ACGT"), ("ACGT", "ACGT"), ("A" * 64, "A" * 63 + "C"), ("ACGT" * 16, "ACGT" * 16),
           ("ACGT" * 17, "TGCA" * 17), ("N" * 100, "ACGTN" * 20), ("ACGT" * 250, "ACGT" * 249 + "AC"), ("A" * 1000, "C" * 1000)]
    for _ in range(n):
        alpha = rng.choice(["ACGT", "ACGT", "ACGTN", "ACGTNacgtRYK-", "AB"])
        la = rng.choice([rng.randint(0, 25), rng.randint(0, 70), rng.randint(60, 70), rng.randint(0, 300), rng.randint(120, 135)])
        a = "".join(rng.choice(alpha) for _ in range(la))
        if rng.random() < 0.6:            # a mutated copy: small distances
            b = list(a)
            for _ in range(rng.randint(0, 6)):
                op = rng.random()
                pos = rng.randint(0, len(b)) if b else 0
                if op < 0.34 and b:
                    b[min(pos, len(b) - 1)] = rng.choice(alpha)
                elif op < 0.67:
                    b.insert(pos, rng.choice(alpha))
                elif b:
                    del b[min(pos, len(b) - 1)]
            b = "".join(b)
        else:
            b = "".join(rng.choice(alpha) for _ in range(rng.choice([la, rng.randint(0, 300)])))
        out.append((a, b))
    return out


def _check(engine, pairs):
    got = engine.edit_distances(pairs)
    for (a, b), g in zip(pairs, got):
        assert int(g) == oracle.edit_distance(a, b), (a, b, int(g))


def test_emulated_kernel_vs_oracle(oracle_lib):
    import emu
    from afterqc_b200 import _abi
    eng = emu.EmuEngine(_abi.Params.defaults())
    _check(eng, _pairs(1, 400))
    L = eng._L                      # the libed.so-compatible symbols of the same library
    for a, b in _pairs(2, 40):
        if len(a) <= 1000 and len(b) <= 1000:
            assert L.edit_distance(a.encode("latin-1"), len(a), b.encode("latin-1"), len(b)) == oracle.edit_distance(a, b)
    eng.close()


def test_emulated_seek_overlap_has_overlap_hm_semantics(oracle_lib):
    """the shim's seek_overlap = util.overlap_hm (not the reference's C function of that name): same triples as the operator entry"""
    import cases
    import emu
    from afterqc_b200 import _abi
    eng = emu.EmuEngine(_abi.Params.defaults())
    orc = oracle_lib.Oracle(_abi.Params.defaults())
    batch = cases.adversarial_batch().slice(0, 300)
    want = orc.ops_pairs(batch)
    L = eng._L
    comp = {"A": "T", "T": "A", "C": "G", "G": "C", "a": "t", "t": "a", "c": "g", "g": "c", "N": "N"}
    for i in range(batch.n):
        r1, _ = batch.read(1, i); r2, _ = batch.read(2, i)
        rc = "".join(comp.get(c, "N") for c in reversed(r2))
        for args in ((3, 30, 50), (3, 50, 30)):      # the reference's call site swaps the two constants (util.py:219-223)
            ret = L.seek_overlap(r1.encode("latin-1"), len(r1), rc.encode("latin-1"), len(r2), *args)
            off, ol, diff = int(want["ov_offset"][i]), int(want["ov_len"][i]), int(want["ov_diff"][i])
            if ol == 0:
                assert ret == 0x7FFFFFFF, (i, ret)
            else:
                assert ret >> 8 == off and ret & 0xFF == min(diff, 255), (i, ret, off, diff)
        assert L.seek_overlap(b"ACGT", 4, b"ACGT", 4, 5, 30, 50) == 0x7FFFFFFF       # other constants are refused
    eng.close(); orc.close()


@pytest.mark.reference
def test_oracle_dp_vs_reference_cpp(oracle_lib):
    """pins the restatement: the reference's own C++ (Myers bit-vector + DP fall-back), compiled where it lies"""
    path = oracle.build_ref()
    assert path is not None
    ref = C.CDLL(path)
    ref.edit_distance.argtypes = [C.c_char_p, C.c_uint, C.c_char_p, C.c_uint]
    ref.edit_distance.restype = C.c_uint
    for a, b in _pairs(3, 600):
        ea, eb = a.encode("latin-1"), b.encode("latin-1")
        assert ref.edit_distance(ea, len(ea), eb, len(eb)) == oracle.edit_distance(a, b), (a, b)


@pytest.mark.gpu
def test_gpu_kernel_vs_oracle(oracle_lib):
    from afterqc_b200 import _abi
    from afterqc_b200.engine import Engine
    eng = Engine(_abi.Params.defaults())
    _check(eng, _pairs(4, 3000))
    L = eng._L
    for a, b in _pairs(5, 50):
        assert L.edit_distance(a.encode("latin-1"), len(a), b.encode("latin-1"), len(b)) == oracle.edit_distance(a, b)
    r1 = "CAGCGCCTACGGGCCCCTTTTTCTGCGCGACCGCGTGGCTGTGGGCGCGGATGCCTTTGAGCGCGGTGACTTCTCACTGCGTATCGAGCCGCTGGAGGTCTCCC"
    r2 = "ACCTCCAGCGGCTCGATACGCAGTGAGAAGTCACCGCGCTCAAAGGCATCCGCGCCCACAGCCACGCGGTCGCGCAGAAAAAGGGGCCCGTAGGCGCGGCTCCC"
    comp = {"A": "T", "T": "A", "C": "G", "G": "C", "N": "N"}
    rc = "".join(comp[c] for c in reversed(r2))
    ret = L.seek_overlap(r1.encode(), len(r1), rc.encode(), len(r2), 3, 30, 50)
    assert (ret >> 8, ret & 0xFF) == (-5, 1)             # the reference's self-check pair: util.overlap_hm -> (-5, 99, 1)
    eng.close()
