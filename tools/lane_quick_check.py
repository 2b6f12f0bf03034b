#!/usr/bin/env python
"""Torch-free smoke of the lane-per-pair kernel on a GPU: numpy-generated PE150 pairs, lane_kernel vs pair_kernel
(records, counters, postfilter QC) and both kernel times.  Starts in a few seconds (no torch import), so it fits the
tail of a GPU budget:   python tools/lane_quick_check.py [pairs]"""
import json
import os
import sys
import time

t_start = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from afterqc_b200 import _abi  # noqa: E402
from afterqc_b200.batch import PackedBatch, SLACK  # noqa: E402
from afterqc_b200.engine import Engine  # noqa: E402


def make(n, L=150, seed=1):
    rng = np.random.default_rng(seed)
    flen = np.clip(np.rint(rng.normal(260, 70, n)).astype(np.int64), 40, 600)
    frag = rng.integers(0, 4, (n, 600), dtype=np.uint8)
    ar = np.arange(L)
    inside = ar[None, :] < flen[:, None]
    c1 = np.where(inside, frag[:, :L], rng.integers(0, 4, (n, L), dtype=np.uint8))
    idx = np.clip(flen[:, None] - 1 - ar[None, :], 0, 599)
    c2 = np.where(inside, 3 - np.take_along_axis(frag, idx, 1), rng.integers(0, 4, (n, L), dtype=np.uint8))
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)

    def finish(c):
        e = rng.random((n, L)) < 0.01
        c = np.where(e, (c + rng.integers(1, 4, (n, L), dtype=np.uint8)) % 4, c)
        q = rng.integers(30, 41, (n, L), dtype=np.uint8)
        q = np.where(e & (rng.random((n, L)) < 0.8), rng.integers(2, 15, (n, L), dtype=np.uint8), q)
        b = bases[c]
        isn = rng.random((n, L)) < 0.001
        b = np.where(isn, ord("N"), b).astype(np.uint8)
        q = np.where(isn, 2, q).astype(np.uint8) + 33
        return b, q
    b1, q1 = finish(c1)
    b2, q2 = finish(c2)
    off = (np.arange(n + 1, dtype=np.int64) * L).astype(np.uint32)

    def col(x):
        return np.concatenate([x.reshape(-1), np.zeros(SLACK, dtype=np.uint8)])
    return PackedBatch(col(b1), col(q1), off, col(b2), col(q2), off.copy())


def tiled(host, k):
    """the batch repeated k times (fast way to a multi-million-pair workload)"""
    n, L = host.n, 150
    def rep(c):
        return np.concatenate([np.tile(c[:n * L], k), np.zeros(SLACK, dtype=np.uint8)])
    off = (np.arange(n * k + 1, dtype=np.int64) * L).astype(np.uint32)
    return PackedBatch(rep(host.seq1), rep(host.qual1), off, rep(host.seq2), rep(host.qual2), off.copy())


def timing(n, k):
    """kernel times only: pair_kernel and lane_kernel with the bench's 2 % statistics mix, lane_kernel without statistics"""
    host = tiled(make(n), k)
    out = {"pairs": host.n, "t_generate_s": round(time.time() - t_start, 2)}
    sums = {}
    for name, kern, qs in (("warp", _abi.KERNEL_WARP, host.n // 50), ("lane", _abi.KERNEL_LANE, host.n // 50), ("lane_nostat", _abi.KERNEL_LANE, 1)):
        eng = Engine(_abi.Params.defaults(filter_kernel=kern, qc_sample=max(1, qs)))
        d = eng.upload(host)
        eng.filter_pairs(d); eng.sync()
        ms = []
        for _ in range(3):
            eng.filter_pairs(d); eng.sync()
            ms.append(round(eng.last_kernel_ms(), 4))
        out[name + "_ms"] = ms
        out[name + "_Mpairs_s"] = round(host.n / (max(min(ms), 1e-6) * 1e-3) / 1e6, 1)
        r = eng.fetch_results(d)
        sums[name] = int(r.view(np.uint32).sum(dtype=np.uint64))
        d.free(); eng.close()
    out["records_checksum_equal"] = sums["warp"] == sums["lane"] == sums["lane_nostat"]
    out["t_total_s"] = round(time.time() - t_start, 2)
    print(json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "timing":
        return timing(int(sys.argv[2]), int(sys.argv[3]))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    host = make(n)
    out = {"pairs": n, "t_generate_s": round(time.time() - t_start, 2)}
    ref = None
    for name, k in (("warp", _abi.KERNEL_WARP), ("lane", _abi.KERNEL_LANE)):
        eng = Engine(_abi.Params.defaults(filter_kernel=k, qc_sample=max(1000, n // 50)))
        d = eng.upload(host)
        eng.filter_pairs(d); eng.sync()
        eng.reset()
        eng.filter_pairs(d); eng.sync()
        out[name + "_ms"] = round(eng.last_kernel_ms(), 4)
        got = (eng.fetch_results(d), eng.counters(), [eng.qc(s) for s in (2, 3)])
        if ref is None:
            ref = got
        else:
            out["records_identical"] = bool(got[0].tobytes() == ref[0].tobytes())
            out["counters_identical"] = bool(np.array_equal(got[1], ref[1]))
            out["qc_identical"] = all(bool(np.array_equal(a[f], b[f])) for a, b in zip(got[2], ref[2]) for f in a.dtype.names)
            if not out["records_identical"]:
                bad = np.flatnonzero((got[0].view(np.uint8).reshape(-1, 32) != ref[0].view(np.uint8).reshape(-1, 32)).any(1))
                out["first_bad"] = [int(x) for x in bad[:5]]
                out["bad_count"] = int(bad.size)
                i = int(bad[0])
                out["lane_rec"] = str(got[0][i]); out["warp_rec"] = str(ref[0][i])
        d.free(); eng.close()
    out["good"] = int((ref[0]["cls"] == 0).sum())
    out["M_pairs_per_s"] = {k: round(n / (max(out[k + "_ms"], 1e-6) * 1e-3) / 1e6, 1) for k in ("warp", "lane")}
    out["t_total_s"] = round(time.time() - t_start, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
