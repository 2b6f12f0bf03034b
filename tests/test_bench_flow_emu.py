"""bench.py's control flow without a GPU (tools/bench_on_emulator.py: torch.cuda stubbed, engine = SIMT emulator, made-up
event times).  Checks every --config: the resident and the host-buffer steps return the same records, the full pipeline resolves
its trims between the two engine calls, the roofline bookkeeping and the keys of the JSON line.  It measures nothing."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline")


@pytest.mark.parametrize("config,extra", [
    ("pe150", ["--qc-sample", "600"]),
    ("pe150", ["--qc-sample", "600", "--filter-kernel", "warp"]),
    ("se100", []),
    ("pe250_full", []),
    ("pe150_err3", ["--qc-sample", "900"]),
])
def test_bench_flow_on_emulator(oracle_lib, config, extra):
    import bench_on_emulator
    n = 1500 if config == "pe250_full" else 2000
    line = bench_on_emulator.run(["--config", config, "--pairs", str(n), "--steps", "1", "--warmup", "1", "--cpu-sample", "800"] + extra)
    assert line is not None
    for k in KEYS:
        assert k in line, k
    assert line["config"]["config"] == config and line["config"]["workload"]
    assert line["e2e"]["results_match_resident"] is True, line["e2e"]
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] == 32 * n
    assert line["gpu_launches"] >= 2
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["achieved"] > 0 and 0 < r["frac"] and r["unit"] == "GB/s"
    assert len(r["phases"]) >= 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    if config == "pe250_full":
        assert line["autotrim_resolved"] is not None
    if config == "se100":
        assert line["unit"] == "M reads/s"


def test_reference_arm_line(oracle_lib, capsys):
    """--impl reference: the oracle on host threads, same config / steps / warm-up keys as the GPU arm"""
    import bench
    out = []
    old_emit, old_argv = bench.emit_json, sys.argv
    bench.emit_json = out.append
    sys.argv = ["bench.py", "--impl", "reference", "--config", "pe250_full", "--pairs", "3000", "--cpu-sample", "600", "--steps", "2", "--warmup", "1"]
    try:
        bench.run_reference_arm(bench.parse_args())
    finally:
        bench.emit_json, sys.argv = old_emit, old_argv
    line = out[0]
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["config"] == "pe250_full" and line["cpu_baseline"]["kind"] == "port"
    # the GPU arm of the same command line carries the same config dictionary
    import bench_on_emulator
    ours = bench_on_emulator.run(["--config", "pe250_full", "--pairs", "3000", "--cpu-sample", "600", "--steps", "2", "--warmup", "1"])
    assert ours["config"] == line["config"]
