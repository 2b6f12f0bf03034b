"""Engine: python face of the C-ABI (include/afterqc_b200.h).

Mirrors the reference's operator interface for the hot path: the batch methods are what the
pipeline uses; overlap()/hasPolyX()/lowQualityNum()/nNumber() are the per-read operators of
util.py:88 and preprocesser.py:30,61,70 routed through the GPU (for operator-level parity tests).
"""
import ctypes as C
import os

import numpy as np

from . import _abi, _native
from .batch import PackedBatch


class EngineError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("afterqc_b200 engine error %d (%s): %s" % (code, _abi.ERR_NAMES.get(code, "?"), msg))
        self.code = code


class DeviceBatch:
    """A batch resident in HBM (columns uploaded once)."""

    def __init__(self, engine, host):
        self.engine = engine
        self.n = host.n
        self.first_index = host.first_index
        self.paired = host.paired
        self.max_len = host.max_len()
        self._ptrs = []
        L = engine._L

        def up(arr, extra=64):
            p = C.c_void_p()
            nbytes = arr.nbytes
            engine._check(L.aqc_device_alloc(engine._h, nbytes + extra, C.byref(p)))
            engine._check(L.aqc_memcpy_h2d(engine._h, p, arr.ctypes.data, nbytes))
            self._ptrs.append(p)
            return p

        self.seq1, self.qual1, self.off1 = up(host.seq1), up(host.qual1), up(host.off1)
        if host.paired:
            self.seq2, self.qual2, self.off2 = up(host.seq2), up(host.qual2), up(host.off2)
        else:
            self.seq2 = self.qual2 = self.off2 = None
        self.results = C.c_void_p()
        engine._check(L.aqc_device_alloc(engine._h, max(1, self.n) * 32, C.byref(self.results)))
        self._ptrs.append(self.results)

    def as_struct(self):
        b = _abi.Batch()
        b.first_index = self.first_index
        b.n = self.n
        b.flags = self.max_len
        b.seq1, b.qual1, b.off1 = self.seq1, self.qual1, self.off1
        b.seq2, b.qual2, b.off2 = self.seq2, self.qual2, self.off2
        return b

    def free(self):
        for p in self._ptrs:
            self.engine._L.aqc_device_free(self.engine._h, p)
        self._ptrs = []


class ParsedText:
    """One mate's records parsed on the device (aqc_fastq_parse_device): the packed base / quality columns, the offsets and
    the line table live in HBM, owned by the engine, until the next parse into the same slot."""

    def __init__(self, engine, st, n_text):
        self.engine = engine
        self.n = int(st.n_records)
        self.consumed = int(st.consumed)
        self.hit_eof = bool(st.hit_eof)
        self.seq_bytes = int(st.seq_bytes)
        self.seq, self.qual, self.off = st.seq, st.qual, st.off
        self.line_start, self.line_len, self.text = st.line_start, st.line_len, st.text
        self.n_text = n_text

    def fetch_lines(self):
        """host copies of the line table only: (line_start, line_len)"""
        e = self.engine
        out = []
        for ptr in (self.line_start, self.line_len):
            a = np.zeros(4 * self.n, dtype=np.uint32)
            if self.n:
                e._check(e._L.aqc_memcpy_d2h(e._h, a.ctypes.data, ptr, a.nbytes))
            out.append(a)
        return tuple(out)

    def fetch(self):
        """host copies: dict(seq, qual, off, line_start, line_len)"""
        e = self.engine

        def down(ptr, count, dtype):
            a = np.zeros(count, dtype=dtype)
            if count:
                e._check(e._L.aqc_memcpy_d2h(e._h, a.ctypes.data, ptr, a.nbytes))
            return a
        if self.n == 0:
            z8, z32 = np.zeros(0, np.uint8), np.zeros(0, np.uint32)
            return dict(seq=z8, qual=z8, off=np.zeros(1, np.uint32), line_start=z32, line_len=z32)
        return dict(seq=down(self.seq, self.seq_bytes, np.uint8), qual=down(self.qual, self.seq_bytes, np.uint8),
                    off=down(self.off, self.n + 1, np.uint32), line_start=down(self.line_start, 4 * self.n, np.uint32),
                    line_len=down(self.line_len, 4 * self.n, np.uint32))


class ParsedDeviceBatch(DeviceBatch):
    """The columns of one or two ParsedText objects as a resident batch for stat_reads / filter_pairs (no copy)."""

    def __init__(self, engine, mate1, mate2=None, first_index=0):
        if mate2 is not None and mate2.n != mate1.n:
            raise ValueError("the two mates hold %d and %d records" % (mate1.n, mate2.n))
        self.engine = engine
        self.n = mate1.n
        self.first_index = first_index
        self.paired = mate2 is not None
        self.max_len = 0                          # no hint: the engine reduces the offsets column
        self._ptrs = []
        self.seq1, self.qual1, self.off1 = mate1.seq, mate1.qual, mate1.off
        self.seq2, self.qual2, self.off2 = (mate2.seq, mate2.qual, mate2.off) if mate2 is not None else (None, None, None)
        self.results = C.c_void_p()
        engine._check(engine._L.aqc_device_alloc(engine._h, max(1, self.n) * 32, C.byref(self.results)))
        self._ptrs.append(self.results)


class Engine:
    def __init__(self, params, device=-1):
        self._L = _native.lib()
        self.params = params
        self._h = C.c_void_p()
        rc = self._L.aqc_create(device, C.byref(params), C.byref(self._h))
        if rc:
            raise EngineError(rc, self._L.aqc_last_error(None).decode())
        if os.environ.get("AQC_TRACE_DEVICE"):
            print("[afterqc_b200] engine on device %d" % device, flush=True)

    # ---- lifecycle -----------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.aqc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise EngineError(rc, self._L.aqc_last_error(self._h).decode())

    def set_params(self, params):
        self.params = params
        self._check(self._L.aqc_set_params(self._h, C.byref(params)))

    def reset(self):
        self._check(self._L.aqc_reset(self._h))

    def reset_filter_counters(self):
        self._check(self._L.aqc_reset_filter(self._h))

    def sync(self):
        self._check(self._L.aqc_sync(self._h))

    def set_stream(self, cuda_stream_handle):
        """Run this engine's kernels on an external CUDA stream (int handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self._L.aqc_set_stream(self._h, C.c_void_p(cuda_stream_handle) if cuda_stream_handle else None))

    def device_ptr(self, what, slot=0):
        p = C.c_void_p()
        n = C.c_uint64()
        self._check(self._L.aqc_device_ptr(self._h, what, slot, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def upload(self, host_batch):
        return DeviceBatch(self, host_batch)

    # ---- hot path --------------------------------------------------------------------------
    def stat_reads(self, batch, qc1, qc2, stat_lo=0, stat_hi=(1 << 63), order_base=0):
        mem = _abi.MEM_DEVICE if isinstance(batch, DeviceBatch) else _abi.MEM_HOST
        b = batch.as_struct()
        self._check(self._L.aqc_stat_reads(self._h, C.byref(b), mem, qc1, qc2, stat_lo, stat_hi, order_base))

    def filter_pairs(self, batch, out=None, qual2_in_place=False):
        """Host batch: returns the result records (copies inside).  DeviceBatch: asynchronous, results stay in HBM.
        qual2_in_place: the host batch's qual2 column is page-locked (aqc_host_alloc / cudaHostAlloc): with the lane-per-pair
        kernel the engine may read it in place over PCIe instead of copying it (AQC_BATCH_QUAL2_IN_PLACE)."""
        if isinstance(batch, DeviceBatch):
            b = batch.as_struct()
            self._check(self._L.aqc_filter_pairs(self._h, C.byref(b), _abi.MEM_DEVICE, batch.results))
            return None
        res = out if out is not None else np.zeros(batch.n, dtype=_abi.RESULT_DTYPE)
        b = batch.as_struct()
        if qual2_in_place:
            b.flags |= _abi.BATCH_QUAL2_IN_PLACE
        self._check(self._L.aqc_filter_pairs(self._h, C.byref(b), _abi.MEM_HOST, res.ctypes.data))
        return res

    def fetch_results(self, dbatch):
        res = np.zeros(dbatch.n, dtype=_abi.RESULT_DTYPE)
        if dbatch.n:
            self._check(self._L.aqc_memcpy_d2h(self._h, res.ctypes.data, dbatch.results, res.nbytes))
        return res

    def ops_pairs(self, batch):
        res = np.zeros(batch.n, dtype=_abi.OPS_DTYPE)
        b = batch.as_struct()
        self._check(self._L.aqc_ops_pairs(self._h, C.byref(b), _abi.MEM_HOST, res.ctypes.data))
        return res

    # ---- fetch ---------------------------------------------------------------------------
    def counters(self):
        out = np.zeros(_abi.C_TOTAL, dtype=np.int64)
        self._check(self._L.aqc_get_counters(self._h, out.ctypes.data))
        return out

    def add_counters(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int64)
        self._check(self._L.aqc_add_counters(self._h, arr.ctypes.data))

    def qc(self, slot):
        out = np.zeros(1, dtype=_abi.QC_DTYPE)
        self._check(self._L.aqc_get_qc(self._h, slot, out.ctypes.data))
        return out[0]

    def kmers(self, slot):
        nk = 1 << (2 * self.params.qc_kmer)
        cnt = np.zeros(nk, dtype=np.uint64)
        first = np.zeros(nk, dtype=np.uint64)
        self._check(self._L.aqc_get_kmer_dense(self._h, slot, cnt.ctypes.data, first.ctypes.data))
        n = C.c_uint32(0)
        self._check(self._L.aqc_get_kmer_side(self._h, slot, None, None, None, 0, C.byref(n)))
        keys = np.zeros(n.value, dtype=np.uint64)
        sc = np.zeros(n.value, dtype=np.uint64)
        sf = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(self._L.aqc_get_kmer_side(self._h, slot, keys.ctypes.data, sc.ctypes.data, sf.ctypes.data, n.value, C.byref(n)))
        order = np.argsort(keys, kind="stable")
        return cnt, first, keys[order], sc[order], sf[order]

    def edit_distances(self, pairs):
        """Levenshtein distance of every (a, b) pair of byte / str strings on the GPU (aqc_edit_distance_batch)"""
        def col(strs):
            bs = [x if isinstance(x, (bytes, bytearray)) else x.encode("latin-1") for x in strs]
            off = np.zeros(len(bs) + 1, dtype=np.uint32)
            np.cumsum([len(x) for x in bs], out=off[1:])
            data = np.frombuffer(b"".join(bs) + b"\0" * 16, dtype=np.uint8).copy()
            return data, off
        a, ao = col([p[0] for p in pairs])
        b, bo = col([p[1] for p in pairs])
        out = np.zeros(len(pairs), dtype=np.int32)
        if len(pairs):
            self._check(self._L.aqc_edit_distance_batch(self._h, a.ctypes.data, ao.ctypes.data, b.ctypes.data, bo.ctypes.data, len(pairs), _abi.MEM_HOST, out.ctypes.data))
        return out

    def parse_fastq(self, text, slot=0, final=True, max_records=(1 << 62)):
        """FASTQ text (bytes / uint8 array, host) -> ParsedText in HBM (aqc_fastq_parse_device; csrc/aqc_parse.cuh).
        Raises ValueError for a record whose quality line is not as long as its sequence line, like the host parser."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, dtype=np.uint8)
        st = _abi.Parsed()
        rc = self._L.aqc_fastq_parse_device(self._h, slot, buf.ctypes.data if buf.size else None, buf.size, _abi.MEM_HOST,
                                            1 if final else 0, max_records, C.byref(st))
        if rc == _abi.ERR_INVALID and "quality line" in self._L.aqc_last_error(self._h).decode():
            raise ValueError("FASTQ record %d: quality line length differs from sequence length" % st.bad_record)
        self._check(rc)
        return ParsedText(self, st, buf.size)

    def parse_fastq_resident(self, dev_ptr, nbytes, slot=0, final=True, max_records=(1 << 62)):
        """the same for text that already lies in HBM (16-byte aligned, 16 bytes of slack): no copy unless the last line lacks its newline"""
        st = _abi.Parsed()
        self._check(self._L.aqc_fastq_parse_device(self._h, slot, dev_ptr, nbytes, _abi.MEM_DEVICE, 1 if final else 0, max_records, C.byref(st)))
        return ParsedText(self, st, nbytes)

    def launch_count(self):
        return int(self._L.aqc_launch_count(self._h))

    def kmer_side_raw(self, slot):
        """(keys, counts, first_direct, first_seed) of the side table as the device holds it (aqc_get_kmer_side_raw)"""
        n = C.c_uint32(0)
        self._check(self._L.aqc_get_kmer_side_raw(self._h, slot, None, None, None, None, 0, C.byref(n)))
        arrs = [np.zeros(n.value, dtype=np.uint64) for _ in range(4)]
        if n.value:
            self._check(self._L.aqc_get_kmer_side_raw(self._h, slot, *[a.ctypes.data for a in arrs], n.value, C.byref(n)))
        return tuple(arrs)

    def last_phase_ms(self, phase):
        """device time of the last call's kernels by phase: 0 filter kernel, 1 list mode, 2 statistics launches, -1 all"""
        return float(self._L.aqc_last_phase_ms(self._h, phase))

    def last_kernel_ms(self):
        return float(self._L.aqc_last_kernel_ms(self._h))

    # ---- reference operator interface (single reads; util.py:88, preprocesser.py:30,61,70) ----
    def _ops1(self, r1, q1, r2=None, q2=None):
        b = PackedBatch.from_reads([(r1, q1)], [(r2, q2)] if r2 is not None else None)
        return self.ops_pairs(b)[0]

    def overlap(self, r1, r2):
        """util.overlap(r1, r2) -> (offset, overlap_len, distance)"""
        o = self._ops1(r1, "I" * len(r1), r2, "I" * len(r2))
        return (int(o["ov_offset"]), int(o["ov_len"]), int(o["ov_diff"]))

    def hasPolyX(self, seq):
        """hasPolyX(seq, maxPoly, mismatch) with the engine's -p/-a parameters; returns base or None"""
        o = self._ops1(seq, "I" * len(seq))
        return chr(int(o["poly1"])) if o["poly1"] else None

    def lowQualityNum(self, read):
        return int(self._ops1(read[1], read[3])["lowq1"])

    def nNumber(self, read):
        return int(self._ops1(read[1], read[3])["n1"])
