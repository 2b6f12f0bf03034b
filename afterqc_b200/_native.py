"""ctypes loader of libafterqc_b200.so.  Fails loudly: there is no CPU fallback in the product."""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AQC_LIB_PATH") or os.path.join(_HERE, "libafterqc_b200.so")   # override: kernel tuning experiments

# every symbol include/afterqc_b200.h declares
SYMBOLS = [
    "aqc_abi_version", "aqc_device_count", "aqc_create", "aqc_destroy", "aqc_set_params", "aqc_reset", "aqc_reset_filter",
    "aqc_last_error", "aqc_host_alloc", "aqc_host_free", "aqc_device_alloc", "aqc_device_free",
    "aqc_memcpy_h2d", "aqc_memcpy_d2h", "aqc_stat_reads", "aqc_filter_pairs", "aqc_ops_pairs", "aqc_sync",
    "aqc_get_counters", "aqc_add_counters", "aqc_get_qc", "aqc_get_kmer_dense", "aqc_get_kmer_side", "aqc_get_kmer_side_raw", "aqc_last_phase_ms",
    "aqc_edit_distance_batch", "edit_distance", "seek_overlap", "aqc_launch_count", "aqc_last_kernel_ms", "aqc_set_stream", "aqc_device_ptr", "aqc_fastq_parse", "aqc_fastq_parse_device", "aqc_fastq_emit", "aqc_fastq_emit_lines",
    "aqc_barcode_pairs", "aqc_gunzip_buffer", "aqc_gunzip_buffer_mt", "aqc_reader_open", "aqc_reader_next", "aqc_reader_release", "aqc_reader_error", "aqc_reader_close", "aqc_text_open", "aqc_text_read",
]

_lib = None


class NativeLibraryMissing(ImportError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "%s is missing: build it with `python -m afterqc_b200.build` (needs nvcc). "
            "The B200 engine has no CPU fallback." % LIB_PATH)
    _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def bind(L):
    """Attach the argument/return types of include/afterqc_b200.h to a loaded library and check its ABI version."""
    vp, i32, u32, u64, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_size_t
    PB, PP = C.POINTER(_abi.Batch), C.POINTER(_abi.Params)
    sig = {
        "aqc_abi_version": (i32, []),
        "aqc_device_count": (i32, []),
        "aqc_create": (i32, [i32, PP, C.POINTER(vp)]),
        "aqc_destroy": (None, [vp]),
        "aqc_set_params": (i32, [vp, PP]),
        "aqc_reset": (i32, [vp]),
        "aqc_reset_filter": (i32, [vp]),
        "aqc_last_error": (C.c_char_p, [vp]),
        "aqc_host_alloc": (i32, [sz, C.POINTER(vp)]),
        "aqc_host_free": (None, [vp]),
        "aqc_device_alloc": (i32, [vp, sz, C.POINTER(vp)]),
        "aqc_device_free": (None, [vp, vp]),
        "aqc_memcpy_h2d": (i32, [vp, vp, vp, sz]),
        "aqc_memcpy_d2h": (i32, [vp, vp, vp, sz]),
        "aqc_stat_reads": (i32, [vp, PB, i32, i32, i32, u64, u64, u64]),
        "aqc_filter_pairs": (i32, [vp, PB, i32, vp]),
        "aqc_ops_pairs": (i32, [vp, PB, i32, vp]),
        "aqc_sync": (i32, [vp]),
        "aqc_get_counters": (i32, [vp, vp]),
        "aqc_add_counters": (i32, [vp, vp]),
        "aqc_get_qc": (i32, [vp, i32, vp]),
        "aqc_get_kmer_dense": (i32, [vp, i32, vp, vp]),
        "aqc_get_kmer_side": (i32, [vp, i32, vp, vp, vp, u32, C.POINTER(u32)]),
        "aqc_get_kmer_side_raw": (i32, [vp, i32, vp, vp, vp, vp, u32, C.POINTER(u32)]),
        "aqc_set_stream": (i32, [vp, vp]),
        "aqc_device_ptr": (i32, [vp, i32, i32, C.POINTER(vp), C.POINTER(u64)]),
        "aqc_fastq_parse": (i32, [vp, u64, i32, u64, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), C.POINTER(u64), C.POINTER(i32), C.POINTER(u64)]),
        "aqc_fastq_emit": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, vp, u64, vp, u64, vp, u64, C.POINTER(u64)]),
        "aqc_barcode_pairs": (i32, [i32, C.c_char_p, u64, C.POINTER(_abi.Columns), C.POINTER(_abi.Columns), C.POINTER(_abi.Columns), C.POINTER(_abi.Columns), vp, C.POINTER(u64)]),
        "aqc_gunzip_buffer": (i32, [vp, u64, vp, u64, C.POINTER(u64), C.c_char_p, u64]),
        "aqc_gunzip_buffer_mt": (i32, [vp, u64, vp, u64, C.POINTER(u64), i32, C.POINTER(u64), C.c_char_p, u64]),
        "aqc_reader_open": (i32, [C.c_char_p, u64, u32, C.POINTER(vp)]),
        "aqc_reader_next": (i32, [vp, C.POINTER(_abi.Records)]),
        "aqc_reader_release": (i32, [vp, u32]),
        "aqc_reader_error": (C.c_char_p, [vp]),
        "aqc_reader_close": (None, [vp]),
        "aqc_text_open": (i32, [C.c_char_p, C.POINTER(vp)]),
        "aqc_text_read": (C.c_int64, [vp, vp, u64]),
        "aqc_launch_count": (u64, [vp]),
        "aqc_edit_distance_batch": (i32, [vp, vp, vp, vp, vp, u32, i32, vp]),
        "aqc_fastq_parse_device": (i32, [vp, i32, vp, u64, i32, i32, u64, vp]),
        "aqc_fastq_emit_lines": (i32, [i32, i32, vp, vp, vp, u64, vp, u64, vp, u64, C.POINTER(u64)]),
        "edit_distance": (C.c_uint, [C.c_char_p, C.c_uint, C.c_char_p, C.c_uint]),
        "seek_overlap": (i32, [C.c_char_p, i32, C.c_char_p, i32, i32, i32, i32]),
        "aqc_last_kernel_ms": (C.c_float, [vp]),
        "aqc_last_phase_ms": (C.c_float, [vp, i32]),
    }
    for name in SYMBOLS:
        fn = getattr(L, name)     # AttributeError if the library does not export it
        fn.restype, fn.argtypes = sig[name]
    if L.aqc_abi_version() != _abi.ABI_VERSION:
        raise ImportError("libafterqc_b200.so ABI %d != python ABI %d" % (L.aqc_abi_version(), _abi.ABI_VERSION))
    return L
