"""The host packer of the packed transport (afterqc_b200/csrc/aqc_pack.cpp) on its own: tests/pack_stress.cpp packs random base
and quality columns (lengths 0..200 k, up to four columns per dispatch, pools of 1..12 threads, all-random-byte columns that
must fall back) and decodes them again; run once with the AVX2 path (where the CPU has it) and once with the portable one."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "afterqc_b200", "csrc")


@pytest.mark.parametrize("scalar", [False, True])
def test_pack_columns_round_trip(tmp_path, scalar):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    exe = str(tmp_path / "pack_stress")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "pack_stress.cpp"),
                           os.path.join(CSRC, "aqc_pack.cpp"), "-o", exe, "-lpthread"])
    env = dict(os.environ)
    if scalar:
        env["AQC_PACK_SCALAR"] = "1"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "stress ok" in r.stdout, r.stdout + r.stderr
