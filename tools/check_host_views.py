"""GPU self-check of the host-buffer entry on column VIEWS: absolute (non-zero) first offsets, batches longer than one
staging chunk, against rebased copies of the same pairs.  numpy + the C-ABI only (no torch import: starts in seconds).
Exit code 0 = identical results and counters."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afterqc_b200 import _abi                      # noqa: E402
from afterqc_b200.batch import PackedBatch         # noqa: E402
from afterqc_b200.engine import Engine             # noqa: E402


def make(n, L=100, seed=3):
    rng = np.random.default_rng(seed)
    frag_len = rng.integers(60, 260, n)
    frag = rng.integers(0, 4, (n, 260), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    len1 = np.minimum(L, frag_len) - rng.integers(0, 3, n)
    len2 = np.minimum(L, frag_len) - rng.integers(0, 3, n)
    cols = np.arange(L)[None, :]
    r1 = lut[frag[:, :L]]
    idx2 = np.clip(frag_len[:, None] - 1 - cols, 0, 259)
    r2 = lut[3 - np.take_along_axis(frag, idx2, axis=1)]          # reverse complement of the fragment's tail (A<->T, C<->G)
    err = rng.random((n, L)) < 0.01
    r2 = np.where(err, lut[rng.integers(0, 4, (n, L))], r2)
    q = rng.choice(np.frombuffer(b"#+5AFIII", dtype=np.uint8), (n, L))

    def pack(r, lens):
        off = np.zeros(n + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
        mask = cols < lens[:, None]
        seq = np.zeros(int(off[-1]) + 64, dtype=np.uint8); qual = np.zeros(int(off[-1]) + 64, dtype=np.uint8)
        seq[:off[-1]] = r[mask]; qual[:off[-1]] = q[mask]
        return seq, qual, off.astype(np.uint32)
    s1, q1, o1 = pack(r1, len1)
    s2, q2, o2 = pack(r2, len2)
    return PackedBatch(s1, q1, o1, s2, q2, o2)


def main():
    n = 600000
    t = time.time()
    b = make(n)
    p = _abi.Params.defaults(); p.qc_sample = 400000
    cuts = [0, 999, 150000, 412345, n]
    e_full, e_view, e_copy = Engine(p), Engine(p), Engine(p)
    r_full = e_full.filter_pairs(b)                                                  # 3 staging chunks
    r_view, r_copy = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        v = PackedBatch(b.seq1, b.qual1, b.off1[lo:hi + 1], b.seq2, b.qual2, b.off2[lo:hi + 1], first_index=lo)   # absolute offsets
        r_view.append(e_view.filter_pairs(v))
        r_copy.append(e_copy.filter_pairs(b.slice(lo, hi)))                          # rebased copy, offsets from 0
    r_view, r_copy = np.concatenate(r_view), np.concatenate(r_copy)
    ok = r_full.tobytes() == r_view.tobytes() == r_copy.tobytes()
    c = [e.counters() for e in (e_full, e_view, e_copy)]
    ok = ok and np.array_equal(c[0], c[1]) and np.array_equal(c[0], c[2])
    for slot in (_abi.QC_R1_POST, _abi.QC_R2_POST):
        q = [e.qc(slot) for e in (e_full, e_view, e_copy)]
        ok = ok and q[0].tobytes() == q[1].tobytes() == q[2].tobytes()
    # stat entry on a window view (the prefilter pass: records [999, 999+limit))
    e_a, e_b = Engine(p), Engine(p)
    lo, hi = 999, 300999
    v = PackedBatch(b.seq1, b.qual1, b.off1[lo:hi + 1], first_index=lo)
    e_a.stat_reads(v, _abi.QC_R1_PRE, -1, stat_lo=lo, stat_hi=hi, order_base=0)
    s = b.slice(lo, hi); s1 = PackedBatch(s.seq1, s.qual1, s.off1, first_index=lo)
    e_b.stat_reads(s1, _abi.QC_R1_PRE, -1, stat_lo=lo, stat_hi=hi, order_base=0)
    ok = ok and e_a.qc(_abi.QC_R1_PRE).tobytes() == e_b.qc(_abi.QC_R1_PRE).tobytes()
    ka, kb = e_a.kmers(_abi.QC_R1_PRE), e_b.kmers(_abi.QC_R1_PRE)
    ok = ok and all(np.array_equal(x, y) for x, y in zip(ka, kb))
    good = int(c[0][_abi.CIDX["GOOD_READS"]])
    print("check_host_views: %s  (n=%d good=%d classes=%s, %.1fs, torch imported: %s)"
          % ("OK" if ok else "MISMATCH", n, good, np.bincount(r_full["cls"], minlength=9).tolist(), time.time() - t, "torch" in sys.modules))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
