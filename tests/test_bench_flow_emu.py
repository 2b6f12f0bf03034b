"""bench.py's control flow without a GPU (tools/bench_on_emulator.py: torch.cuda stubbed, engine = SIMT emulator, child checks
in-process with made-up kernel times).  Checks the kernel auto-selection and its fall-backs, that the host-buffer path returns
the resident path's records, and the keys of the JSON line.  It measures nothing."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

ARGS = ["--pairs", "2000", "--qc-sample", "600", "--steps", "1", "--warmup", "1", "--cpu-sample", "1000"]
KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline")


@pytest.mark.parametrize("fail,want_kernel,want_stat", [
    ((), "lane", "lane_post"),                         # every candidate identical: fastest filter kernel + statistics in their own launch
    (("lane", "lane2"), "warp", "lane ("),             # no lane-per-pair kernel: pair_kernel, prefilter statistics may still switch
    (("lane_st3",), "lane", "lane ("),                 # the deferred form fails: statistics with one lane per read inside the filter kernel
    (("lane2", "lane_st2", "lane_st3"), "lane", "warp"),   # failing statistics candidates do not cost the filter kernel its place
])
def test_bench_flow_on_emulator(oracle_lib, fail, want_kernel, want_stat):
    import bench_on_emulator
    line = bench_on_emulator.run(ARGS, fail)
    assert line is not None
    for k in KEYS:
        assert k in line, k
    sel = line["config"]["filter_kernel"]
    assert sel["used"] == want_kernel, sel
    assert sel["stat_kernel_used"].startswith(want_stat), sel
    assert line["e2e"]["results_match_resident"] is True
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] == 32 * 2000
    assert line["gpu_launches"] >= 2
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["achieved"] > 0 and 0 < r["frac"] and r["unit"] == "GB/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    assert line["config"]["workload"]
