"""aqc_fastq_parse_device on a B200: the cases of tests/test_parse_device.py through the CUDA library, and a text above 1 GiB
(more than 256 K blocks of 4 KB: every level of the newline scan)."""
import numpy as np
import pytest

import cases
import test_parse_device as T
from afterqc_b200 import _abi

pytestmark = pytest.mark.gpu


def test_parse_device_cases():
    from afterqc_b200.engine import Engine
    T.run_cases(lambda p: Engine(p))


def test_parse_device_large_text():
    from afterqc_b200.engine import Engine
    batch = cases.synthetic("pe150", 4000)
    unit = T.fastq_text(batch, 1)
    reps = (1100 << 20) // len(unit) + 1
    text = np.tile(np.frombuffer(unit, dtype=np.uint8), reps)
    eng = Engine(_abi.Params.defaults())
    p = eng.parse_fastq(text)
    assert p.n == 4000 * reps and p.consumed == text.size and p.hit_eof
    d = p.fetch()
    one = eng.parse_fastq(unit).fetch()
    L = one["seq"].size
    assert d["seq"].size == L * reps
    for r in (0, 1, reps // 2, reps - 1):
        assert np.array_equal(d["seq"][r * L:(r + 1) * L], one["seq"])
        assert np.array_equal(d["qual"][r * L:(r + 1) * L], one["qual"])
        assert np.array_equal(d["off"][r * 4000:(r + 1) * 4000 + 1].astype(np.int64) - r * L, one["off"].astype(np.int64))
    ls = d["line_start"].astype(np.int64)
    assert np.all(np.diff(ls) > 0)
    eng.close()
