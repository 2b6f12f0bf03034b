"""Struct-for-struct comparison of two backends (oracle vs engine) with readable diagnostics."""
import numpy as np

from afterqc_b200 import _abi


def describe_pair(batch, i):
    s1, q1 = batch.read(1, i)
    out = ["pair %d" % i, " r1 %s" % s1, " q1 %s" % q1]
    if batch.paired:
        s2, q2 = batch.read(2, i)
        out += [" r2 %s" % s2, " q2 %s" % q2]
    return "\n".join(out)


def assert_records_equal(batch, a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape
    if a.tobytes() == b.tobytes():
        return
    names = [n for n in a.dtype.names if n != "pad"]
    for i in range(len(a)):
        for n in names:
            if not np.array_equal(a[i][n], b[i][n]):
                raise AssertionError("%s: record %d field %s: oracle=%s engine=%s\noracle rec=%s\nengine rec=%s\n%s" % (
                    what, i, n, a[i][n], b[i][n], a[i], b[i], describe_pair(batch, i)))


def assert_counters_equal(a, b, what):
    if np.array_equal(a, b):
        return
    inv = {v: k for k, v in _abi.CIDX.items()}
    bad = np.flatnonzero(a != b)
    msgs = []
    for j in bad[:20]:
        j = int(j)
        if j in inv:
            name = inv[j]
        elif _abi.C_ERR_MATRIX <= j < _abi.C_ERR_MATRIX + 16:
            name = "ERR[%s->%s]" % (_abi.ALL_BASES[(j - 32) // 4], _abi.ALL_BASES[(j - 32) % 4])
        elif j >= _abi.C_DISTANCE_HIST:
            name = "distance_hist[%d]" % (j - _abi.C_DISTANCE_HIST)
        elif j >= _abi.C_OVERLAP_HIST:
            name = "overlap_hist[%d]" % (j - _abi.C_OVERLAP_HIST)
        else:
            name = "idx%d" % j
        msgs.append("%s oracle=%d engine=%d" % (name, a[j], b[j]))
    raise AssertionError("%s: counters differ: %s" % (what, "; ".join(msgs)))


def assert_qc_equal(a, b, what):
    for n in a.dtype.names:
        if not np.array_equal(a[n], b[n]):
            x, y = np.asarray(a[n]), np.asarray(b[n])
            idx = np.argwhere(x != y)[:8]
            raise AssertionError("%s: QC field %s differs at %s: oracle=%s engine=%s" % (
                what, n, idx.tolist(), [x[tuple(k)] for k in idx], [y[tuple(k)] for k in idx]))


def assert_kmers_equal(a, b, what):
    names = ["dense_counts", "dense_first", "side_keys", "side_counts", "side_first"]
    for n, x, y in zip(names, a, b):
        if x.shape != y.shape or not np.array_equal(x, y):
            if x.shape != y.shape:
                raise AssertionError("%s: k-mer %s shape oracle=%s engine=%s" % (what, n, x.shape, y.shape))
            idx = np.flatnonzero(x != y)[:8]
            raise AssertionError("%s: k-mer %s differs at %s: oracle=%s engine=%s" % (what, n, idx.tolist(), x[idx].tolist(), y[idx].tolist()))


def compare_backends(orc, eng, slots, what):
    assert_counters_equal(orc.counters(), eng.counters(), what)
    for s in slots:
        assert_qc_equal(orc.qc(s), eng.qc(s), "%s slot %d" % (what, s))
        assert_kmers_equal(orc.kmers(s), eng.kmers(s), "%s slot %d" % (what, s))
