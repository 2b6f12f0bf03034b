"""TEST INFRASTRUCTURE: the engine's CUDA sources compiled for the host and run under a SIMT emulator.

`build()` compiles afterqc_b200/csrc/*.cu|cpp with g++ (-DAQC_EMU, tests/emu/cuda_runtime.h standing in for the CUDA
runtime) into tests/emu/_build/libafterqc_b200_emu.so, which exports the same C-ABI; `EmuEngine` is the python Engine
bound to that library.  Purpose: check the DEVICE code (warp collectives, tile ring, per-lane arithmetic) against the
oracle on a machine without a GPU, e.g. while developing a kernel.  Nothing in the product imports this package and the
emulated library is never what `afterqc_b200._native` loads.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "afterqc_b200", "csrc")
# AQC_EMU_DEFINES="-DAQC_LANE_IMAD_SHIFT,..." builds (and loads) a tuning variant of the kernels under the emulator
_DEFINES = [d for d in os.environ.get("AQC_EMU_DEFINES", "").split(",") if d]
_TAG = ("_" + "".join(c if c.isalnum() else "_" for c in "".join(_DEFINES))) if _DEFINES else ""
OUT = os.path.join(HERE, "_build", "libafterqc_b200_emu%s.so" % _TAG)
SOURCES = ["aqc_engine.cu", "aqc_fastq.cpp", "aqc_stream.cpp", "aqc_inflate.cpp", "aqc_pinflate.cpp"]

_lib = None


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d += [os.path.join(HERE, "cuda_runtime.h"), os.path.join(HERE, "simt_emu.cpp"), os.path.join(ROOT, "include", "afterqc_b200.h")]
    return d


def build(force=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(p) <= os.path.getmtime(OUT) for p in _deps()):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-DAQC_EMU"] + _DEFINES + ["-I" + HERE, "-fPIC", "-shared", "-x", "c++"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "simt_emu.cpp"), "-o", OUT, "-lz", "-lpthread"]
    subprocess.check_call(cmd)
    return OUT


def lib():
    global _lib
    if _lib is None:
        from afterqc_b200 import _native
        _lib = _native.bind(C.CDLL(build()))
    return _lib


def EmuEngine(params, device=0):
    """afterqc_b200.engine.Engine whose kernels run under the emulator."""
    from afterqc_b200 import engine as _engine

    class _Emu(_engine.Engine):
        def __init__(self, p, d):
            self._L = lib()
            self.params = p
            self._h = C.c_void_p()
            rc = self._L.aqc_create(d, C.byref(p), C.byref(self._h))
            if rc:
                raise _engine.EngineError(rc, self._L.aqc_last_error(None).decode())

    return _Emu(params, device)
