"""FASTQ ingest/egress for the packed-column engine (host side).

Mirrors the reference's fastq.Reader/Writer semantics (fastq.py:17-104):
  * gz / bz2 / plain by file extension (:23-28);
  * every line is rstrip()'d; the first EMPTY line (after rstrip) ends the file, and a record
    needs four non-empty lines (quirk Q13, :41-48);
  * writers append ".gz" when forced, gzip level from --compression (:62-68), bz2 output refused.
Records are parsed in vectorised chunks straight into packed columns (names, bases, '+' lines,
qualities + offsets) instead of python str lists.
"""
import bz2
import gzip
import io
import sys

import numpy as np

from .batch import PackedBatch, SLACK

_WS = np.zeros(256, dtype=bool)
for _c in b" \t\n\r\x0b\x0c":
    _WS[_c] = True


def is_fastq(f):
    return f.endswith((".fq", ".fastq", ".fq.gz", ".fastq.gz", ".fq.bz2", ".fastq.bz2"))


def _open_read(path):
    if path.endswith(".gz"):
        return gzip.open(path, "rb")
    if path.endswith(".bz2"):
        return bz2.open(path, "rb")
    return open(path, "rb")


class Column:
    """A packed column of variable-length byte strings."""

    __slots__ = ("data", "off")

    def __init__(self, data, off):
        self.data = data
        self.off = off

    def __len__(self):
        return len(self.off) - 1

    def get(self, i):
        return self.data[int(self.off[i]):int(self.off[i + 1])].tobytes()

    @staticmethod
    def concat(cols):
        if len(cols) == 1:
            return cols[0]
        data = np.concatenate([c.data for c in cols])
        offs = [cols[0].off.astype(np.int64)]
        base = int(cols[0].off[-1])
        for c in cols[1:]:
            offs.append(c.off[1:].astype(np.int64) + base)
            base += int(c.off[-1])
        return Column(data, np.concatenate(offs))


class FastqRecords:
    """n records as four packed columns."""

    def __init__(self, names, seqs, plus, quals):
        self.names, self.seqs, self.plus, self.quals = names, seqs, plus, quals
        self.n = len(names)

    def lengths(self):
        return np.diff(self.seqs.off.astype(np.int64))


def _parse_block(buf, final):
    """Parse complete records out of `buf` (bytes).  Returns (records|None, consumed_bytes, hit_eof)."""
    data = np.frombuffer(buf, dtype=np.uint8)
    nl = np.flatnonzero(data == 10)
    nlines = len(nl)
    ends = nl
    if final and (len(data) > 0) and (nlines == 0 or nl[-1] != len(data) - 1):
        ends = np.append(nl, len(data))          # last line without trailing newline
        nlines += 1
    if nlines == 0:
        return None, 0, final
    starts = np.empty(nlines, dtype=np.int64)
    starts[0] = 0
    starts[1:] = ends[:-1] + 1
    e = ends.astype(np.int64).copy()
    # rstrip: drop trailing whitespace bytes
    while True:
        m = (e > starts) & _WS[data[np.maximum(e - 1, 0)]]
        if not m.any():
            break
        e[m] -= 1
    lens = e - starts
    nrec = nlines // 4           # a partial trailing group is carried over (or dropped at EOF)
    hit_eof = False
    empty = np.flatnonzero(lens[:nrec * 4] == 0)
    if len(empty):
        # the first empty line ends the file (fastq.py:44-47); its record is incomplete
        nrec = int(empty[0]) // 4
        hit_eof = True
    if nrec == 0:
        return None, 0, hit_eof or final
    consumed = int(ends[nrec * 4 - 1]) + 1

    def column(k):
        s = starts[k:nrec * 4:4]
        l = lens[k:nrec * 4:4]
        off = np.zeros(nrec + 1, dtype=np.int64)
        np.cumsum(l, out=off[1:])
        total = int(off[-1])
        idx = np.repeat(s - off[:-1], l) + np.arange(total, dtype=np.int64)
        return Column(data[idx], off)

    rec = FastqRecords(column(0), column(1), column(2), column(3))
    if not np.array_equal(rec.seqs.off, rec.quals.off):
        bad = int(np.flatnonzero(np.diff(rec.seqs.off) != np.diff(rec.quals.off))[0])
        raise ValueError("FASTQ record %d: quality line length differs from sequence length" % bad)
    return rec, consumed, hit_eof or final


def iter_records(path, block_bytes=32 << 20):
    """Yield FastqRecords chunks in file order."""
    f = _open_read(path)
    carry = b""
    try:
        while True:
            blk = f.read(block_bytes)
            final = len(blk) == 0
            buf = carry + blk
            if not buf:
                return
            rec, consumed, eof = _parse_block(buf, final)
            if rec is not None:
                yield rec
            if eof or final:
                return
            carry = buf[consumed:]
    finally:
        f.close()


def read_all(path):
    chunks = list(iter_records(path))
    if not chunks:
        z = np.zeros(1, dtype=np.int64)
        e = np.zeros(0, dtype=np.uint8)
        return FastqRecords(Column(e, z), Column(e, z), Column(e, z), Column(e, z))
    return FastqRecords(*(Column.concat([getattr(c, k) for c in chunks]) for k in ("names", "seqs", "plus", "quals")))


def to_batch(rec1, rec2, lo, hi, first_index=None):
    """PackedBatch of records [lo, hi) of rec1 (and rec2 if not None)."""
    def cut(rec):
        a, b = int(rec.seqs.off[lo]), int(rec.seqs.off[hi])
        s = np.zeros(b - a + SLACK, dtype=np.uint8)
        q = np.zeros(b - a + SLACK, dtype=np.uint8)
        s[:b - a] = rec.seqs.data[a:b]
        q[:b - a] = rec.quals.data[a:b]
        return s, q, (rec.seqs.off[lo:hi + 1] - a).astype(np.uint32)
    s1, q1, o1 = cut(rec1)
    fi = lo if first_index is None else first_index
    if rec2 is None:
        return PackedBatch(s1, q1, o1, first_index=fi)
    s2, q2, o2 = cut(rec2)
    return PackedBatch(s1, q1, o1, s2, q2, o2, first_index=fi)


class Writer:
    """fastq.Writer (fastq.py:57-104) on bytes."""

    def __init__(self, fname, force_gzip=False, gzip_compression=2):
        self.filename = fname
        if not self.filename.endswith(".gz") and force_gzip:
            self.filename = self.filename + ".gz"
        if self.filename.endswith(".gz"):
            self._f = gzip.open(self.filename, "wb", compresslevel=gzip_compression)
        elif self.filename.endswith(".bz2"):
            print("ERROR: Write bzip2 stream is not supported")
            sys.exit(1)
        else:
            self._f = open(self.filename, "wb")

    def write(self, data):
        self._f.write(data)

    def flush(self):
        self._f.flush()

    def close(self):
        if self._f is not None:
            self._f.flush()
            self._f.close()
            self._f = None
