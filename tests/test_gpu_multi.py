"""2-GPU NCCL path of the drop-in pipeline (after.py --gpus 2): must reproduce the single-GPU outputs.
Skipped on boxes with fewer than 2 GPUs."""
import gzip
import json
import os
import subprocess
import sys

import pytest

import golden_util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_cli_two_gpus_equals_one(tmp_path):
    import shutil
    src = os.path.join(golden_util.GOLD, "pe150_default")
    outs = {}
    for tag, extra in (("one", []), ("two", ["--gpus", "2"])):
        d = tmp_path / tag
        d.mkdir()
        for fn in ("x_R1.fq.gz", "x_R2.fq.gz"):
            shutil.copy(os.path.join(src, fn), str(d / fn))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "after.py"), "-1", str(d / "x_R1.fq.gz"), "-2", str(d / "x_R2.fq.gz"),
                            "-g", str(d / "good")] + extra, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-3000:]
        with open(str(d / "QC" / "x_R1.fq.gz.json")) as f:
            j = json.load(f)
        for k in ("read1_file", "read2_file", "good_output_folder"):
            j["command"][k] = None
        outs[tag] = (j, {rel: gzip.open(str(d / rel)).read() for rel in
                         ("good/x_R1.good.fq.gz", "good/x_R2.good.fq.gz", "bad/x_R1.bad.fq.gz", "bad/x_R2.bad.fq.gz")})
    assert outs["one"][0] == outs["two"][0]
    assert outs["one"][1] == outs["two"][1]
