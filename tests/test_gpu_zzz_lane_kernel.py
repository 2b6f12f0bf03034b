"""GPU parity of the lane-per-pair filter kernel (lane_kernel, aqc_params.filter_kernel = 2).

The kernel was developed in a session that had no GPU minutes left: before its first run on hardware it had only been
checked under the SIMT emulator (tests/test_emu_parity.py).  It is therefore opt-in in the engine, its check runs in a
child process with a timeout (a hang must not take the suite down) and the tests are non-strict xfail until a round has
seen them pass on a B200 -- an XPASS here is the signal to make them plain tests and flip the engine default."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REASON = "lane_kernel not yet run on hardware (round 1 ended without GPU minutes); emulator-verified only"


def _run(args, timeout):
    try:
        r = subprocess.run([sys.executable, os.path.join(HERE, "lane_gpu_check.py")] + args, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        pytest.fail("lane_gpu_check %s timed out after %d s (kernel hang?)" % (args, timeout))
    assert r.returncode == 0, "lane_gpu_check %s failed:\n%s\n%s" % (args, r.stdout[-2000:], r.stderr[-4000:])
    return r.stdout


@pytest.mark.xfail(reason=REASON, strict=False)
def test_lane_kernel_parity_vs_oracle():
    out = _run(["parity"], 600)
    assert "lane kernel parity ok" in out


@pytest.mark.xfail(reason=REASON, strict=False)
def test_lane_kernel_equals_warp_kernel_at_bench_size():
    out = _run(["full", "2000000"], 600)
    assert '"identical": true' in out
