"""The background native reader (csrc/aqc_stream.cpp, NativeStream) against the block parser (read_all) on the same
files: parsing quirks, batch boundaries, gzip members, errors, slot recycling, zero-copy batches."""
import gzip
import os

import numpy as np
import pytest

from afterqc_b200 import fastq_io
from afterqc_b200.batch import SLACK


def _write(path, data):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wb") as f:
        f.write(data)


def _drain(path, batch, take=None, slots=4, hold=0):
    """all records through NativeStream as python tuples; `hold` views are kept alive before done()"""
    s = fastq_io.NativeStream(path, batch, slots)
    out, held = [], []
    while True:
        k = s.available(take or batch)
        if k == 0:
            break
        v = s.take(k)
        for i in range(v.n):
            out.append((v.names.get(i), v.seqs.get(i), v.plus.get(i), v.quals.get(i)))
        held.append(v)
        while len(held) > hold:
            held.pop(0).done()
    for v in held:
        v.done()
    s.close()
    return out


def _all(path):
    r = fastq_io.read_all(path)
    return [(r.names.get(i), r.seqs.get(i), r.plus.get(i), r.quals.get(i)) for i in range(r.n)]


def _random_fastq(n, seed, crlf=False):
    rng = np.random.default_rng(seed)
    nl = b"\r\n" if crlf else b"\n"
    recs = []
    for i in range(n):
        L = int(rng.integers(1, 120))
        s = bytes(rng.choice(list(b"ACGTN"), L).tolist())
        q = bytes(rng.integers(33, 74, L, dtype=np.uint8).tolist())
        recs.append(b"@n%d some text" % i + nl + s + nl + b"+" + (b"n%d" % i if i % 3 == 0 else b"") + nl + q + nl)
    return b"".join(recs)


@pytest.mark.parametrize("name", ["a.fq", "a.fq.gz"])
@pytest.mark.parametrize("batch", [1, 7, 64, 1000])
def test_reader_matches_block_parser(tmp_path, name, batch):
    p = str(tmp_path / name)
    _write(p, _random_fastq(333, 5, crlf=(batch == 7)))
    want = _all(p)
    assert len(want) == 333
    assert _drain(p, batch) == want
    assert _drain(p, batch, take=5, slots=2) == want            # partial takes inside a batch
    assert _drain(p, batch, hold=2) == want                     # consumer holds two batches (writers in flight)


def test_reader_quirks(tmp_path):
    p = str(tmp_path / "q.fq")
    for data, n in ((b"@a\nAC\n+\nII\n@b\n \t\n+\nII\n@c\nAC\n+\nII\n", 1),       # blank (after rstrip) line = EOF, Q13
                    (b"@a\nAC\n+\nII\n\n@c\nAC\n+\nII\n", 1),
                    (b"@a\r\nACG\r\n+\r\nIII\r\n@b\nTT\n+\nII", 2),                 # CRLF, no final newline
                    (b"@a\nACG\n+\nIII\n@b\nTT\n+\n", 1),                           # partial last record dropped
                    (b"", 0), (b"\n", 0), (b"@a\nAC  \n+\nII\t\n", 1)):
        _write(p, data)
        assert _drain(p, 4) == _all(p) and len(_all(p)) == n, data
    _write(p, b"@a\nAC  \n+\nII\t\n")
    assert _drain(p, 4) == [(b"@a", b"AC", b"+", b"II")]


def test_reader_concatenated_gzip_members(tmp_path):
    p = str(tmp_path / "m.fq.gz")
    a, b = _random_fastq(50, 1), _random_fastq(70, 2)
    with open(p, "wb") as f:
        f.write(gzip.compress(a) + gzip.compress(b[:1000]) + gzip.compress(b[1000:]))
    assert len(_drain(p, 16)) == 120 and _drain(p, 16) == _all(p)


def test_reader_errors(tmp_path):
    with pytest.raises(IOError):
        fastq_io.NativeStream(str(tmp_path / "missing.fq"), 8)
    p = str(tmp_path / "e.fq")
    good = _random_fastq(20, 3)
    _write(p, good + b"@bad\nACGT\n+\nIII\n" + good)
    s = fastq_io.NativeStream(p, 8)
    n = 0
    with pytest.raises(ValueError, match="record 20"):
        while True:
            k = s.available(8)
            if k == 0:
                break
            v = s.take(k); n += v.n; v.done()
    assert n == 20                                               # the records before the bad one were delivered
    s.close()
    p = str(tmp_path / "t.fq.gz")
    z = gzip.compress(_random_fastq(2000, 4))
    with open(p, "wb") as f:
        f.write(z[:len(z) // 2])                                 # truncated gzip stream
    with pytest.raises(ValueError):
        _drain(p, 64)


def test_reader_batches_are_zero_copy_columns(tmp_path):
    p = str(tmp_path / "z.fq")
    _write(p, _random_fastq(100, 9))
    want = _all(p)
    s = fastq_io.NativeStream(p, 32)
    g = 0
    while True:
        k = s.available(10)
        if k == 0:
            break
        v = s.take(k)
        b = fastq_io.to_batch(v, None, 0, k, first_index=g)
        assert b.n == k and b.first_index == g and b.off1.dtype == np.uint32
        assert np.shares_memory(b.seq1, v.seqs.data) and len(b.seq1) >= int(b.off1[-1]) + SLACK
        for i in range(k):
            assert b.read(1, i) == (want[g + i][1].decode(), want[g + i][3].decode())
        sub = fastq_io.to_batch(v, None, 2, k, first_index=g + 2) if k > 2 else None
        if sub is not None:
            assert sub.read(1, 0) == (want[g + 2][1].decode(), want[g + 2][3].decode())
        cp = v.slice(0, k)                                       # a copy that survives the slot
        v.done()
        assert cp.seqs.get(0) == want[g][1]
        g += k
    assert g == 100
    s.close()


def test_reader_close_while_reading_ahead(tmp_path):
    p = str(tmp_path / "c.fq")
    _write(p, _random_fastq(3000, 11))
    for slots in (2, 4):
        s = fastq_io.NativeStream(p, 100, slots)
        v = s.take(s.available(100))
        assert v.n == 100
        s.close()                                                # reader thread is mid-file; must not hang
