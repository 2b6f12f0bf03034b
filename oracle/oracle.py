"""ctypes binding of the CPU oracle (oracle/aqc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; afterqc_b200/ never does.  The class mirrors afterqc_b200.engine.Engine so
parity tests call both with the same arguments.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from afterqc_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libaqc_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "aqc_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "afterqc_b200.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.aqo_create.argtypes = [C.POINTER(_abi.Params), C.POINTER(C.c_void_p)]
        L.aqo_destroy.argtypes = [C.c_void_p]
        L.aqo_destroy.restype = None
        L.aqo_set_params.argtypes = [C.c_void_p, C.POINTER(_abi.Params)]
        L.aqo_last_error.argtypes = [C.c_void_p]
        L.aqo_last_error.restype = C.c_char_p
        L.aqo_stat_reads.argtypes = [C.c_void_p, C.POINTER(_abi.Batch), C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
        L.aqo_filter_pairs.argtypes = [C.c_void_p, C.POINTER(_abi.Batch), C.c_void_p]
        L.aqo_ops_pairs.argtypes = [C.c_void_p, C.POINTER(_abi.Batch), C.c_void_p]
        L.aqo_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.aqo_get_qc.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.aqo_get_kmer_dense.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.aqo_get_kmer_side.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.aqo_edit_distance.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]
        L.aqo_edit_distance.restype = C.c_int
        L.aqo_reset.argtypes = [C.c_void_p]
        L.aqo_reset_filter.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def edit_distance(a, b):
    """Levenshtein distance of two byte strings by the oracle's dynamic programme"""
    a = a if isinstance(a, (bytes, bytearray)) else a.encode("latin-1")
    b = b if isinstance(b, (bytes, bytearray)) else b.encode("latin-1")
    return int(lib().aqo_edit_distance(bytes(a), len(a), bytes(b), len(b)))


_REF_ED = os.path.join(_HERE, "_ref", "libed_ref.so")


def build_ref(reference="/root/reference"):
    """oracle/_ref/libed_ref.so: the reference's editdistance/_editdistance.cpp compiled where it lies (build container only)"""
    if not os.path.isfile(os.path.join(reference, "editdistance", "_editdistance.cpp")):
        return None
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref", "REFERENCE=" + reference])
    return _REF_ED


class OracleError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("oracle error %d (%s) %s" % (code, _abi.ERR_NAMES.get(code, "?"), msg))
        self.code = code


class Oracle:
    """CPU oracle with the Engine interface (stat_reads / filter_pairs / ops_pairs / counters / qc / kmers)."""

    def __init__(self, params):
        self.params = params
        self._h = C.c_void_p()
        rc = lib().aqo_create(C.byref(params), C.byref(self._h))
        if rc:
            raise OracleError(rc, "create")

    def close(self):
        if self._h:
            lib().aqo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise OracleError(rc, lib().aqo_last_error(self._h).decode())

    def set_params(self, params):
        self.params = params
        self._check(lib().aqo_set_params(self._h, C.byref(params)))

    def reset(self):
        self._check(lib().aqo_reset(self._h))

    def reset_filter_counters(self):
        self._check(lib().aqo_reset_filter(self._h))

    def stat_reads(self, batch, qc1, qc2, stat_lo=0, stat_hi=(1 << 63), order_base=0):
        b = batch.as_struct()
        self._check(lib().aqo_stat_reads(self._h, C.byref(b), qc1, qc2, stat_lo, stat_hi, order_base))

    def filter_pairs(self, batch):
        res = np.zeros(batch.n, dtype=_abi.RESULT_DTYPE)
        b = batch.as_struct()
        self._check(lib().aqo_filter_pairs(self._h, C.byref(b), res.ctypes.data))
        return res

    def ops_pairs(self, batch):
        res = np.zeros(batch.n, dtype=_abi.OPS_DTYPE)
        b = batch.as_struct()
        self._check(lib().aqo_ops_pairs(self._h, C.byref(b), res.ctypes.data))
        return res

    def counters(self):
        out = np.zeros(_abi.C_TOTAL, dtype=np.int64)
        self._check(lib().aqo_get_counters(self._h, out.ctypes.data))
        return out

    def qc(self, slot):
        out = np.zeros(1, dtype=_abi.QC_DTYPE)
        self._check(lib().aqo_get_qc(self._h, slot, out.ctypes.data))
        return out[0]

    def kmers(self, slot):
        """(dense_counts, dense_first, side_keys, side_counts, side_first)"""
        nk = 1 << (2 * self.params.qc_kmer)
        cnt = np.zeros(nk, dtype=np.uint64)
        first = np.zeros(nk, dtype=np.uint64)
        self._check(lib().aqo_get_kmer_dense(self._h, slot, cnt.ctypes.data, first.ctypes.data))
        n = C.c_uint32(0)
        lib().aqo_get_kmer_side(self._h, slot, None, None, None, 0, C.byref(n))
        keys = np.zeros(n.value, dtype=np.uint64); sc = np.zeros(n.value, dtype=np.uint64); sf = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(lib().aqo_get_kmer_side(self._h, slot, keys.ctypes.data, sc.ctypes.data, sf.ctypes.data, n.value, C.byref(n)))
        order = np.argsort(keys, kind="stable")
        return cnt, first, keys[order], sc[order], sf[order]
