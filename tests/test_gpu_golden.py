"""GPU end-to-end: the full drop-in pipeline on the CUDA engine reproduces the reference's golden outputs
(JSON + decompressed good/bad/overlap files), and the after.py command line runs."""
import os
import subprocess
import sys

import pytest

import golden_util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", golden_util.CASES)
def test_pipeline_on_engine_matches_reference_golden(name, tmp_path):
    from afterqc_b200.engine import Engine
    problems = golden_util.run_case(name, tmp_path, lambda p: Engine(p))
    assert not problems, problems


def test_after_py_cli_runs_on_gpu(tmp_path):
    import shutil
    src = os.path.join(golden_util.GOLD, "pe150_default")
    for fn in ("x_R1.fq.gz", "x_R2.fq.gz"):
        shutil.copy(os.path.join(src, fn), str(tmp_path / fn))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "after.py"), "-1", str(tmp_path / "x_R1.fq.gz"), "-2", str(tmp_path / "x_R2.fq.gz"),
                          "-g", str(tmp_path / "good")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Time used:" in out.stdout
    for rel in ("good/x_R1.good.fq.gz", "bad/x_R2.bad.fq.gz", "QC/x_R1.fq.gz.json"):
        assert os.path.exists(str(tmp_path / rel)), rel


def test_directory_mode_on_gpu(tmp_path):
    """after.py -d DIR: one job per *R1* file (after.py:101-171); outputs land in good/ bad/ QC/ next to the default names."""
    import shutil
    d = tmp_path / "run"
    d.mkdir()
    for tag, case in (("a", "pe150_default"), ("b", "pe150_small_head_fallback")):
        src = os.path.join(golden_util.GOLD, case)
        shutil.copy(os.path.join(src, "x_R1.fq.gz"), str(d / ("s%s_R1.fq.gz" % tag)))
        shutil.copy(os.path.join(src, "x_R2.fq.gz"), str(d / ("s%s_R2.fq.gz" % tag)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "after.py"), "-d", str(d), "-g", str(d / "good")],
                         capture_output=True, text=True, timeout=900, cwd=str(d))
    assert out.returncode == 0, out.stderr[-2000:]
    for tag in ("a", "b"):
        for rel in ("good/s%s_R1.good.fq.gz" % tag, "good/s%s_R2.good.fq.gz" % tag, "bad/s%s_R1.bad.fq.gz" % tag, "QC/s%s_R1.fq.gz.json" % tag,
                    "QC/s%s_R1.fq.gz.html" % tag):
            assert os.path.exists(str(d / rel)), rel


def test_directory_mode_spreads_jobs_over_gpus(tmp_path):
    """after.py:168-171 starts one process per R1 file; here job i runs on GPU i mod (GPUs of the box) and every job reproduces
    the single-GPU outputs of its case"""
    import shutil
    from afterqc_b200 import cli
    if cli.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = tmp_path / "run"
    d.mkdir()
    cases_ = (("a", "pe150_default"), ("b", "pe150_small_head_fallback"), ("c", "testdata"))
    for tag, case in cases_:
        src = os.path.join(golden_util.GOLD, case)
        shutil.copy(os.path.join(src, "x_R1.fq.gz"), str(d / ("s%s_R1.fq.gz" % tag)))
        shutil.copy(os.path.join(src, "x_R2.fq.gz"), str(d / ("s%s_R2.fq.gz" % tag)))
    env = dict(os.environ, AQC_TRACE_DEVICE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "after.py"), "-d", str(d), "-g", str(d / "good")],
                         capture_output=True, text=True, timeout=900, cwd=str(d), env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    used = sorted(set(ln.split()[-1] for ln in out.stdout.splitlines() if ln.startswith("[afterqc_b200] engine on device")))
    assert len(used) >= 2, out.stdout[-2000:]
    import json
    for tag, case in cases_:
        with open(os.path.join(golden_util.GOLD, case, "expected.json")) as f:
            exp = json.load(f)
        for rel, digest in exp["outputs_sha256"].items():
            p = str(d / rel.replace("x_R", "s%s_R" % tag))
            assert os.path.exists(p) and golden_util.sha(p) == digest, rel
