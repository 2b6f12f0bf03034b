"""GPU parity of the lane-per-pair filter kernel (lane_kernel, aqc_params.filter_kernel = 2).

The kernel was developed under the SIMT emulator (tests/test_emu_parity.py) with the round's GPU budget all but spent: the
last seconds of it showed records, counters and postfilter QC identical to pair_kernel on 400 k PE150 pairs and 2.5x its
throughput on 2 M pairs (profiles/r01_lane_kernel_first_gpu_run.json, r01_lane_kernel_timing_2M.json).  That is what the
plain test below repeats.  The full parameter/length matrix against the oracle has not run on hardware yet, so that test is
a non-strict xfail (an XPASS is the signal to make it plain and flip the engine default), and both run in a child process
with a timeout so that a hang cannot take the suite down."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REASON = "the full lane_kernel matrix has not run on hardware yet (round 1 ran out of GPU minutes); emulator-verified"


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


def test_lane_kernel_equals_warp_kernel_at_bench_size():
    out = _run(["full", "2000000", "lane", "noinplace"], 600)
    assert '"identical": true' in out


@pytest.mark.xfail(reason="AQC_BATCH_QUAL2_IN_PLACE (the kernel reads mate-2 qualities from page-locked host memory) has not run on hardware yet; emulator-verified", strict=False)
def test_lane_kernel_host_path_with_qual2_in_place():
    out = _run(["full", "2000000", "lane"], 600)
    assert '"in_place_ok": true' in out


_GROUP = {}
GROUP_CANDIDATES = ["lane2", "warp_st2", "lane_st2", "lane_st3", "lane2_st3"]


def _group_verdicts():
    """ONE child process for all never-on-hardware candidates at bench size (one torch import, one workload, pair_kernel once);
    a verdict line is printed per candidate as soon as it is known, so a later crash or hang costs only the candidates after it"""
    import json
    if not _GROUP:
        _GROUP["ran"] = True
        cmd = [sys.executable, os.path.join(HERE, "lane_gpu_check.py"), "full", "2000000", ",".join(GROUP_CANDIDATES), "noinplace"]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=480)
            out = r.stdout or ""
        except subprocess.TimeoutExpired as e:
            out = e.stdout or ""
            if isinstance(out, bytes):
                out = out.decode("utf-8", "replace")
        for ln in out.splitlines():
            if ln.startswith("{"):
                try:
                    j = json.loads(ln)
                except ValueError:
                    continue
                _GROUP[j.get("candidate")] = j
    return _GROUP


@pytest.mark.xfail(reason="lane2_kernel and the lane-per-read statistics (stat_kernel = 2 / 3) were written after the round's GPU budget was spent; "
                          "emulator-verified", strict=False)
@pytest.mark.parametrize("cand", GROUP_CANDIDATES)
def test_candidate_equals_warp_kernel_at_bench_size(cand):
    v = _group_verdicts().get(cand)
    assert v is not None, "no verdict for %s (the child crashed or hung before it)" % cand
    assert v.get("identical") is True, v


@pytest.mark.xfail(reason="lane2_kernel has not run on hardware yet (written after the round's GPU budget was spent); emulator-verified", strict=False)
def test_lane2_kernel_parity_vs_oracle():
    out = _run(["parity", "lane2"], 300)
    assert "lane2 kernel parity ok" in out


ST2_REASON = "stat_tile / stat_lane_kernel (aqc_params.stat_kernel = 2 / 3) were written after the round's GPU budget was spent; emulator-verified"


@pytest.mark.xfail(reason=ST2_REASON, strict=False)
@pytest.mark.parametrize("cand", ["lane_st2", "lane_st3"])
def test_stat2_parity_vs_oracle(cand):
    out = _run(["parity", cand], 300)
    assert cand + " kernel parity ok" in out


@pytest.mark.xfail(reason="AQC_BATCH_PACK_BASES (host threads pack the bases to 2 bits, unpack_bases_kernel restores them) was written after the "
                          "round's GPU budget was spent; emulator-verified", strict=False)
def test_packed_base_transport_parity_vs_oracle(monkeypatch):
    monkeypatch.setenv("AQC_CHUNK_PAIRS", "3000")       # several chunks per batch: both staging slots, aligned chunk origins
    out = _run(["pack"], 300)
    assert "pack_bases parity ok" in out
