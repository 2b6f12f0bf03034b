"""In-tree build of libafterqc_b200.so (hand-written sm_100a kernels + C-ABI) with nvcc."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libafterqc_b200.so")
SOURCES = ["aqc_engine.cu", "aqc_fastq.cpp", "aqc_stream.cpp", "aqc_inflate.cpp", "aqc_pinflate.cpp"]
DEPS = ["aqc_stat_kernel.cuh", "aqc_engine.cu", "aqc_fastq.cpp", "aqc_stream.cpp", "aqc_inflate.cpp", "aqc_inflate.hpp", "aqc_pinflate.cpp", "aqc_pinflate.hpp", "aqc_kernel.cuh", "aqc_device.cuh", "aqc_lane_kernel.cuh", os.path.join("..", "..", "include", "afterqc_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-lz", "-lpthread"]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
