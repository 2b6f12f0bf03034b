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
import os
import sys

import numpy as np

from .batch import PackedBatch, SLACK

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
        data = np.concatenate([c.data[int(c.off[0]):int(c.off[-1])] for c in cols])
        offs = [cols[0].off.astype(np.int64) - int(cols[0].off[0])]
        base = int(offs[0][-1])
        for c in cols[1:]:
            offs.append(c.off[1:].astype(np.int64) - int(c.off[0]) + base)
            base += int(c.off[-1]) - int(c.off[0])
        return Column(data, np.concatenate(offs))


class FastqRecords:
    """n records as four packed columns."""

    def __init__(self, names, seqs, plus, quals):
        self.names, self.seqs, self.plus, self.quals = names, seqs, plus, quals
        self.n = len(names)

    def lengths(self):
        return np.diff(self.seqs.off.astype(np.int64))

    def slice(self, a, b):
        """records [a, b) as an independent FastqRecords (offsets rebased)"""
        def cut(c):
            lo, hi = int(c.off[a]), int(c.off[b])
            return Column(c.data[lo:hi], (c.off[a:b + 1] - lo).astype(np.int64))
        return FastqRecords(cut(self.names), cut(self.seqs), cut(self.plus), cut(self.quals))

    def done(self):
        """nothing to give back (RecordsView overrides this)"""

    @staticmethod
    def concat(chunks):
        if len(chunks) == 1:
            return chunks[0]
        return FastqRecords(*(Column.concat([getattr(c, k) for c in chunks]) for k in ("names", "seqs", "plus", "quals")))

    @staticmethod
    def empty():
        z = np.zeros(1, dtype=np.int64)
        e = np.zeros(0, dtype=np.uint8)
        return FastqRecords(Column(e, z), Column(e, z), Column(e, z), Column(e, z))


class RecordStream:
    """Pull-based view of one FASTQ file: take(k) returns the next min(k, remaining) records."""

    def __init__(self, path, block_bytes=32 << 20):
        self._it = iter_records(path, block_bytes)
        self._chunks = []       # pending FastqRecords
        self._have = 0
        self._ended = False

    def _fill(self, k):
        while self._have < k and not self._ended:
            try:
                c = next(self._it)
            except StopIteration:
                self._ended = True
                break
            self._chunks.append(c)
            self._have += c.n

    def available(self, k):
        """number of records that can be taken now, up to k"""
        self._fill(k)
        return min(k, self._have)

    def take(self, k):
        self._fill(k)
        k = min(k, self._have)
        out = []
        need = k
        while need > 0:
            c = self._chunks[0]
            if c.n <= need:
                out.append(c); self._chunks.pop(0); need -= c.n
            else:
                out.append(c.slice(0, need)); self._chunks[0] = c.slice(need, c.n); need = 0
        self._have -= k
        return FastqRecords.concat(out) if out else FastqRecords.empty()

    def close(self):
        self._it = iter(())
        self._chunks, self._have, self._ended = [], 0, True


def _view(ptr, nbytes, dtype=np.uint8):
    """numpy view of native memory (no copy; the owner must outlive it)"""
    import ctypes as C
    if nbytes == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_uint8 * nbytes).from_address(ptr), dtype=dtype)


class RecordsView(FastqRecords):
    """Records [lo, hi) of one batch of the native reader: columns are views into the reader's slot buffers, valid
    until done().  `off32`/`seq_full`/`qual_full` are the aqc_batch columns of this mate (absolute offsets into the
    slot's base/quality columns), so to_batch() builds a PackedBatch without copying."""

    def __init__(self, owner, slot, cols, off32, max_len, lo, hi):
        names, seqs, plus, quals = [Column(d, o[lo:hi + 1]) for d, o in cols]
        FastqRecords.__init__(self, names, seqs, plus, quals)
        self.n = hi - lo
        self._owner, self._slot = owner, slot
        self.off32 = off32[lo:hi + 1]
        self.max_len = max_len
        self._cols, self._off32_full, self._lo = cols, off32, lo

    def slice(self, a, b):
        """copying slice: the result does not depend on the slot"""
        def cut(c):
            lo, hi = int(c.off[a]), int(c.off[b])
            return Column(c.data[lo:hi].copy(), (c.off[a:b + 1] - lo).astype(np.int64))
        return FastqRecords(cut(self.names), cut(self.seqs), cut(self.plus), cut(self.quals))

    def done(self):
        """this view is no longer used (the slot goes back to the reader when all its views are done)"""
        if self._owner is not None:
            self._owner._view_done(self._slot)
            self._owner = None


class NativeStream:
    """RecordStream on the native background reader (csrc/aqc_stream.cpp): available(k)/take(k) without copies.
    Batches of all streams opened with the same batch size have the same record boundaries, so lock-stepped consumers
    see the same k on every stream until the shortest file ends."""

    def __init__(self, path, batch_records, slots=4):
        import ctypes as C
        from . import _native, _abi
        self._L = _native.lib()
        self._h = C.c_void_p()
        self._rec_t = _abi.Records
        rc = self._L.aqc_reader_open(path.encode(), int(batch_records), int(slots), C.byref(self._h))
        if rc:
            self._h = None
            raise IOError("cannot open %s" % path)
        self._cur = None            # (slot, cols, off32, max_len, n)
        self._pos = 0
        self._ended = False
        self._out = {}              # slot -> outstanding views (+1 while it is the current batch)

    def _fetch(self):
        import ctypes as C
        r = self._rec_t()
        rc = self._L.aqc_reader_next(self._h, C.byref(r))
        if rc:
            raise ValueError(self._L.aqc_reader_error(self._h).decode() or "FASTQ reader failed (%d)" % rc)
        if r.n == 0:
            self._ended = True
            self._cur = None
            return
        n = int(r.n)
        offs = [_view(r.off[c], 8 * (n + 1), np.int64) for c in range(4)]
        tot = [int(o[-1]) for o in offs]
        cols = [(_view(r.bytes[0], tot[0]), offs[0]), (_view(r.bytes[1], tot[1] + 64), offs[1]),
                (_view(r.bytes[2], tot[2]), offs[2]), (_view(r.bytes[3], tot[1] + 64), offs[1])]
        self._cur = (int(r.slot), cols, _view(r.seq_off32, 4 * (n + 1), np.uint32), int(r.max_len), n)
        self._pos = 0
        self._out[int(r.slot)] = 1

    def available(self, k):
        if self._cur is None and not self._ended:
            self._fetch()
        if self._cur is None:
            return 0
        return min(k, self._cur[4] - self._pos)

    def take(self, k):
        k = self.available(k)
        if k == 0:
            return FastqRecords.empty()
        slot, cols, off32, max_len, n = self._cur
        v = RecordsView(self, slot, cols, off32, max_len, self._pos, self._pos + k)
        self._out[slot] += 1
        self._pos += k
        if self._pos >= n:
            self._cur = None
            self._view_done(slot)
        return v

    def _view_done(self, slot):
        self._out[slot] -= 1
        if self._out[slot] == 0:
            del self._out[slot]
            if self._h is not None:
                self._L.aqc_reader_release(self._h, slot)

    def close(self):
        if self._h is not None:
            self._L.aqc_reader_close(self._h)
            self._h = None
            self._cur = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def open_stream(path, batch_records, slots=4):
    """background native reader for plain/.gz files; the python reader for .bz2 (no bzip2 in the native library)"""
    if path.endswith(".bz2"):
        return RecordStream(path)
    return NativeStream(path, batch_records, slots)


def _parse_block(buf, final):
    """Parse complete records out of `buf` (bytes) with the native parser (csrc/aqc_fastq.cpp).
    Returns (records|None, consumed_bytes, hit_eof)."""
    import ctypes as C
    from . import _native
    L = _native.lib()
    n = len(buf)
    if n == 0:
        return None, 0, final
    max_rec = n // 8 + 1                      # a record needs at least 4 x (1 byte + newline)
    src = np.frombuffer(buf, dtype=np.uint8)
    cols = [np.empty(n, dtype=np.uint8) for _ in range(4)]
    offs = [np.zeros(max_rec + 1, dtype=np.uint64) for _ in range(4)]
    pb = (C.c_void_p * 4)(*[c.ctypes.data for c in cols])
    po = (C.c_void_p * 4)(*[o.ctypes.data for o in offs])
    nrec, consumed, bad = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    eof = C.c_int(0)
    rc = L.aqc_fastq_parse(src.ctypes.data, n, 1 if final else 0, max_rec, pb, po,
                           C.byref(nrec), C.byref(consumed), C.byref(eof), C.byref(bad))
    if rc:
        raise ValueError("FASTQ record %d: quality line length differs from sequence length" % bad.value)
    k = int(nrec.value)
    if k == 0:
        return None, int(consumed.value), bool(eof.value) or final
    out = []
    for c, o in zip(cols, offs):
        o = o[:k + 1].astype(np.int64)
        out.append(Column(c[:int(o[-1])].copy(), o))
    return FastqRecords(*out), int(consumed.value), bool(eof.value) or final


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
    return FastqRecords.concat(chunks) if chunks else FastqRecords.empty()


def to_batch(rec1, rec2, lo, hi, first_index=None):
    """PackedBatch of records [lo, hi) of rec1 (and rec2 if not None)."""
    def cut(rec):
        if isinstance(rec, RecordsView):      # zero copy: absolute offsets into the slot's columns
            return rec._cols[1][0], rec._cols[3][0], rec.off32[lo:hi + 1]
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


def barcode_transform(rec1, rec2, barcode_length, verify):
    """The barcode (UMI) pre-pass of the per-read loop (native aqc_barcode_pairs; preprocesser.py:435-452): returns
    (t1, t2, status, removed) -- the records in input order with barcodes moved into the names and barcode / verify /
    read-through bases removed, status per pair (0 ok, 1 BADBCD1, 2 BADBCD2: copied unchanged), removed bases per mate."""
    import ctypes as C
    from . import _native, _abi
    L = _native.lib()
    n = rec1.n

    def cin(rec):
        for c in (rec.names, rec.seqs):
            if c.off.dtype != np.int64 or not c.off.flags["C_CONTIGUOUS"]:
                c.off = np.ascontiguousarray(c.off, dtype=np.int64)
        return _abi.Columns(rec.names.data.ctypes.data, rec.names.off.ctypes.data, rec.seqs.data.ctypes.data,
                            rec.seqs.off.ctypes.data, rec.quals.data.ctypes.data)

    def cout(rec):
        nb = int(rec.names.off[n] - rec.names.off[0]) + n * (barcode_length + 3)
        sb = int(rec.seqs.off[n] - rec.seqs.off[0])
        arrs = (np.empty(nb, dtype=np.uint8), np.zeros(n + 1, dtype=np.int64), np.zeros(sb + 64, dtype=np.uint8),
                np.zeros(n + 1, dtype=np.int64), np.zeros(sb + 64, dtype=np.uint8))
        return arrs, _abi.Columns(*[a.ctypes.data for a in arrs])

    i1 = cin(rec1)
    a1, o1 = cout(rec1)
    i2 = o2 = a2 = None
    if rec2 is not None:
        i2 = cin(rec2)
        a2, o2 = cout(rec2)
    status = np.zeros(n, dtype=np.uint8)
    removed = (C.c_uint64 * 2)()
    rc = L.aqc_barcode_pairs(int(barcode_length), verify.encode("latin-1"), n, C.byref(i1), C.byref(i2) if i2 is not None else None,
                             C.byref(o1), C.byref(o2) if o2 is not None else None, status.ctypes.data, removed)
    if rc:
        raise ValueError("aqc_barcode_pairs failed (%d): barcode_length must be >= 1" % rc)

    def wrap(rec, arrs):
        names, noff, seqs, soff, quals = arrs
        lo, hi = int(rec.plus.off[0]), int(rec.plus.off[n])
        plus = Column(rec.plus.data[lo:hi].copy(), (rec.plus.off[:n + 1] - lo).astype(np.int64))
        return FastqRecords(Column(names[:int(noff[n])], noff), Column(seqs, soff), plus, Column(quals, soff))

    return wrap(rec1, a1), (wrap(rec2, a2) if rec2 is not None else None), status, (int(removed[0]), int(removed[1]))


def emit(rec, mate, which, rec_base, results):
    """FASTQ text (bytes) of records rec_base.. of one mate selected by `which` (0 good, 1 bad, 2 overlap tails),
    with the slices and edits of the aqc_result records applied (native, csrc/aqc_fastq.cpp)."""
    data, _ = emit_into(rec, mate, which, rec_base, results, None)
    return data.tobytes() if len(data) else b""


class TextSource:
    """read(n) over the text of a FASTQ file for callers that parse elsewhere (the device parser): plain and .gz files come
    from the native reader's sources (mapping / multi-threaded gzip decoder, csrc/aqc_stream.cpp) without its parser, .bz2
    from python's bz2.  read(n) returns n bytes (as a uint8 array) unless the file ends first."""

    BLOCK = 4 << 20

    def __init__(self, path):
        import ctypes as C
        self._py = None
        self._h = None
        self._keep = np.zeros(0, dtype=np.uint8)
        self._eof = False
        if path.endswith(".bz2"):
            import bz2
            self._py = bz2.open(path, "rb")
            return
        from . import _native
        self._L = _native.lib()
        self._h = C.c_void_p()
        if self._L.aqc_text_open(path.encode(), C.byref(self._h)):
            self._h = None
            raise IOError("cannot open %s" % path)

    def read(self, n):
        if self._py is not None:
            return np.frombuffer(self._py.read(n), dtype=np.uint8)
        parts, have = [], 0
        if len(self._keep):
            parts.append(self._keep); have = len(self._keep)
            self._keep = np.zeros(0, dtype=np.uint8)
        while have < n and not self._eof:
            cap = max(self.BLOCK, ((n - have + self.BLOCK - 1) // self.BLOCK) * self.BLOCK)
            buf = np.empty(cap, dtype=np.uint8)
            got = int(self._L.aqc_text_read(self._h, buf.ctypes.data, cap))
            if got < 0:
                raise ValueError(self._L.aqc_reader_error(self._h).decode() or "text source failed")
            if got == 0:
                self._eof = True
                break
            parts.append(buf[:got]); have += got
        data = np.concatenate(parts) if len(parts) > 1 else (parts[0] if parts else np.zeros(0, dtype=np.uint8))
        if len(data) > n:
            self._keep = data[n:]
            data = data[:n]
        return data

    def close(self):
        if self._py is not None:
            self._py.close(); self._py = None
        if self._h is not None:
            self._L.aqc_reader_close(self._h); self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TextRecords:
    """Records that were parsed on the device (Engine.parse_fastq): the host keeps the FASTQ text and the line table
    (start and length of the four rstrip()ped lines of every record); the writers format from those."""

    def __init__(self, text, line_start, line_len, n):
        self.text = text
        self.line_start = np.ascontiguousarray(line_start, dtype=np.uint32)
        self.line_len = np.ascontiguousarray(line_len, dtype=np.uint32)
        self.n = n

    def lengths(self):
        return self.line_len[1::4][:self.n].astype(np.int64)

    def done(self):
        pass


def _emit_lines_into(rec, mate, which, rec_base, results, scratch):
    import ctypes as C
    from . import _native
    L = _native.lib()
    n = len(results)
    if n == 0:
        return np.zeros(0, dtype=np.uint8), scratch
    ll = rec.line_len[4 * rec_base:4 * (rec_base + n)].astype(np.int64)
    cap = int(ll.sum()) + int(ll[1::4].sum()) + 20 * n + 64
    if scratch is None or len(scratch) < cap:
        scratch = np.empty(cap + cap // 8, dtype=np.uint8)
    olen = C.c_uint64(0)
    res = np.ascontiguousarray(results)
    rc = L.aqc_fastq_emit_lines(mate, which, rec.text.ctypes.data, rec.line_start.ctypes.data, rec.line_len.ctypes.data,
                                rec_base, res.ctypes.data, n, scratch.ctypes.data, cap, C.byref(olen))
    if rc:
        raise RuntimeError("aqc_fastq_emit_lines failed (%d)" % rc)
    return scratch[:olen.value], scratch


def emit_into(rec, mate, which, rec_base, results, scratch):
    """emit() into a reusable uint8 scratch array: returns (view of the text, scratch to pass next time)."""
    import ctypes as C
    from . import _native
    if isinstance(rec, TextRecords):
        return _emit_lines_into(rec, mate, which, rec_base, results, scratch)
    L = _native.lib()
    n = len(results)
    if n == 0:
        return np.zeros(0, dtype=np.uint8), scratch
    a, b = rec_base, rec_base + n
    cap = int((rec.names.off[b] - rec.names.off[a]) + (rec.plus.off[b] - rec.plus.off[a]) + 2 * (rec.seqs.off[b] - rec.seqs.off[a])) + 20 * n + 64
    if scratch is None or len(scratch) < cap:
        scratch = np.empty(cap + cap // 8, dtype=np.uint8)
    out = scratch
    olen = C.c_uint64(0)
    res = np.ascontiguousarray(results)
    cols = [rec.names, rec.seqs, rec.plus, rec.quals]
    for c in cols:
        if c.off.dtype != np.int64 or not c.off.flags["C_CONTIGUOUS"]:
            c.off = np.ascontiguousarray(c.off, dtype=np.int64)
    rc = L.aqc_fastq_emit(mate, which, rec.names.data.ctypes.data, rec.names.off.ctypes.data, rec.seqs.data.ctypes.data,
                          rec.seqs.off.ctypes.data, rec.plus.data.ctypes.data, rec.plus.off.ctypes.data, rec.quals.data.ctypes.data,
                          rec_base, res.ctypes.data, n, out.ctypes.data, cap, C.byref(olen))
    if rc:
        raise RuntimeError("aqc_fastq_emit failed (%d)" % rc)
    return out[:olen.value], scratch


class _ParallelGzip:
    """gzip output as a sequence of independent members (valid .gz: readers concatenate them), each block compressed
    by a worker thread (zlib releases the GIL).  Content after decompression is what the reference writes; the
    compressed bytes differ, as they already do between zlib builds (parity is on decompressed content, quirk Q14)."""

    BLOCK = 8 << 20

    def __init__(self, path, level, threads=None):
        import concurrent.futures
        if threads is None:                       # zlib level 2 deflates ~70 MB/s per thread: scale with the host, 4..12 per file
            threads = max(4, min(12, (os.cpu_count() or 8) // 8))
        self._f = open(path, "wb")
        self._level = level
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=threads)
        self._threads = threads
        self._pending = []
        self._buf = []
        self._size = 0

    @staticmethod
    def _member(data, level):
        import zlib
        c = zlib.compressobj(level, zlib.DEFLATED, 31)
        return c.compress(data) + c.flush()

    def write(self, data):
        if not len(data):
            return
        if not isinstance(data, bytes):
            data = bytes(data)            # callers may reuse their buffer
        self._buf.append(data)
        self._size += len(data)
        if self._size >= self.BLOCK:
            self._submit()

    def _submit(self):
        if self._size == 0:
            return
        data = b"".join(self._buf)
        self._buf, self._size = [], 0
        self._pending.append(self._pool.submit(self._member, data, self._level))
        while len(self._pending) > 2 * self._threads:  # bounded queue, in-order write-out
            self._f.write(self._pending.pop(0).result())

    def flush(self):
        pass

    def close(self):
        self._submit()
        for fut in self._pending:
            self._f.write(fut.result())
        self._pending = []
        if self._f.tell() == 0:                       # an empty .gz is still a gzip member
            self._f.write(self._member(b"", self._level))
        self._pool.shutdown()
        self._f.close()


class Writer:
    """fastq.Writer (fastq.py:57-104) on bytes."""

    def __init__(self, fname, force_gzip=False, gzip_compression=2):
        self.filename = fname
        if not self.filename.endswith(".gz") and force_gzip:
            self.filename = self.filename + ".gz"
        if self.filename.endswith(".gz"):
            self._f = _ParallelGzip(self.filename, gzip_compression)
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
