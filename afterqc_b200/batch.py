"""Packed SoA read batches: the data layout the kernels consume (DESIGN.md "data layout").

One uint8 column of bases, one of qualities and an (n+1) uint32 offsets column per mate; reads
are variable length.  Columns carry 16 bytes of zero slack because the device stages tiles with
16-byte granular bulk copies.
"""
import ctypes as C

import numpy as np

from . import _abi

SLACK = 16
MAX_COLUMN_BYTES = (1 << 32) - 64   # offsets are uint32


def _as_bytes(s):
    if isinstance(s, (bytes, bytearray)):
        return bytes(s)
    return s.encode("latin-1")


class _Pinned:
    """page-locked allocations of a pinned PackedBatch, freed with it"""

    def __init__(self, lib):
        self.lib, self.ptrs = lib, []

    def __del__(self):
        for p in self.ptrs:
            self.lib.aqc_host_free(p)
        self.ptrs = []


class PackedBatch:
    """Host-side packed batch.  seq2/qual2/off2 are None for single-end input."""

    def __init__(self, seq1, qual1, off1, seq2=None, qual2=None, off2=None, first_index=0):
        self.seq1, self.qual1, self.off1 = seq1, qual1, off1
        self.seq2, self.qual2, self.off2 = seq2, qual2, off2
        self.first_index = int(first_index)
        self.n = int(len(off1) - 1)
        assert off1.dtype == np.uint32 and seq1.dtype == np.uint8 and qual1.dtype == np.uint8
        assert len(seq1) >= int(off1[-1]) + SLACK and len(qual1) >= int(off1[-1]) + SLACK
        if seq2 is not None:
            assert len(off2) == len(off1)
            assert len(seq2) >= int(off2[-1]) + SLACK and len(qual2) >= int(off2[-1]) + SLACK

    @property
    def paired(self):
        return self.seq2 is not None

    @property
    def bytes1(self):
        return int(self.off1[-1])

    @property
    def bytes2(self):
        return int(self.off2[-1]) if self.paired else 0

    def max_len(self):
        m = int(np.diff(self.off1.astype(np.int64)).max()) if self.n else 0
        if self.paired and self.n:
            m = max(m, int(np.diff(self.off2.astype(np.int64)).max()))
        return m

    # ---- constructors -------------------------------------------------------------------
    @staticmethod
    def _pack(seqs, quals):
        n = len(seqs)
        lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=n)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        if off[-1] > MAX_COLUMN_BYTES:
            raise ValueError("batch column exceeds uint32 offsets; split the batch")
        total = int(off[-1])
        seq = np.zeros(total + SLACK, dtype=np.uint8)
        qual = np.zeros(total + SLACK, dtype=np.uint8)
        seq[:total] = np.frombuffer(b"".join(_as_bytes(s) for s in seqs), dtype=np.uint8)
        qb = b"".join(_as_bytes(q) for q in quals)
        if len(qb) != total:
            raise ValueError("quality/sequence length mismatch")
        qual[:total] = np.frombuffer(qb, dtype=np.uint8)
        return seq, qual, off.astype(np.uint32)

    @classmethod
    def from_reads(cls, reads1, reads2=None, first_index=0):
        """reads*: sequences of (seq, qual) str/bytes tuples."""
        for s, q in reads1:
            if len(s) != len(q):
                raise ValueError("quality/sequence length mismatch")
        s1, q1, o1 = cls._pack([r[0] for r in reads1], [r[1] for r in reads1])
        if reads2 is None:
            return cls(s1, q1, o1, first_index=first_index)
        if len(reads2) != len(reads1):
            raise ValueError("mate count mismatch")
        for s, q in reads2:
            if len(s) != len(q):
                raise ValueError("quality/sequence length mismatch")
        s2, q2, o2 = cls._pack([r[0] for r in reads2], [r[1] for r in reads2])
        return cls(s1, q1, o1, s2, q2, o2, first_index=first_index)

    def slice(self, lo, hi):
        """Sub-batch [lo, hi) (copies; offsets rebased)."""
        lo = max(0, int(lo)); hi = min(self.n, int(hi))
        def cut(seq, qual, off):
            a, b = int(off[lo]), int(off[hi])
            s = np.zeros(b - a + SLACK, dtype=np.uint8); q = np.zeros(b - a + SLACK, dtype=np.uint8)
            s[:b - a] = seq[a:b]; q[:b - a] = qual[a:b]
            return s, q, (off[lo:hi + 1].astype(np.int64) - a).astype(np.uint32)
        s1, q1, o1 = cut(self.seq1, self.qual1, self.off1)
        if not self.paired:
            return PackedBatch(s1, q1, o1, first_index=self.first_index + lo)
        s2, q2, o2 = cut(self.seq2, self.qual2, self.off2)
        return PackedBatch(s1, q1, o1, s2, q2, o2, first_index=self.first_index + lo)

    def pinned(self, engine):
        """Copy of this batch whose columns live in page-locked host memory (aqc_host_alloc), as the host-buffer entry wants
        them for asynchronous copies and for AQC_BATCH_QUAL2_IN_PLACE.  The memory is released with the returned object."""
        lib = engine._L
        owner = _Pinned(lib)

        def pin(arr):
            if arr is None:
                return None
            p = C.c_void_p()
            rc = lib.aqc_host_alloc(max(1, arr.nbytes), C.byref(p))
            if rc:
                raise MemoryError("aqc_host_alloc(%d) failed" % arr.nbytes)
            owner.ptrs.append(p)
            view = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(1, arr.nbytes),))[:arr.nbytes].view(arr.dtype)
            view[...] = arr
            return view
        out = PackedBatch(pin(self.seq1), pin(self.qual1), pin(self.off1), pin(self.seq2), pin(self.qual2), pin(self.off2), self.first_index)
        out._owner = owner
        return out

    # ---- views ---------------------------------------------------------------------------
    def read(self, mate, i):
        """(seq, qual) of record i as python str (latin-1)."""
        if mate == 1:
            a, b = int(self.off1[i]), int(self.off1[i + 1])
            return self.seq1[a:b].tobytes().decode("latin-1"), self.qual1[a:b].tobytes().decode("latin-1")
        a, b = int(self.off2[i]), int(self.off2[i + 1])
        return self.seq2[a:b].tobytes().decode("latin-1"), self.qual2[a:b].tobytes().decode("latin-1")

    def as_struct(self):
        """ctypes aqc_batch pointing at the host columns (keep self alive while it is used)."""
        b = _abi.Batch()
        b.first_index = self.first_index
        b.n = self.n
        b.flags = 0
        b.seq1 = self.seq1.ctypes.data; b.qual1 = self.qual1.ctypes.data; b.off1 = self.off1.ctypes.data
        if self.paired:
            b.seq2 = self.seq2.ctypes.data; b.qual2 = self.qual2.ctypes.data; b.off2 = self.off2.ctypes.data
        else:
            b.seq2 = None; b.qual2 = None; b.off2 = None
        return b
