"""Barcoded (UMI) files through the full pipeline on the CUDA engine, against the reference's golden outputs
(tests/golden_barcode/, made by oracle/make_golden.py).  The barcode pre-pass is host code; the device sees compacted
sub-batches as column views.  Named to run after the parity tests."""
import pytest

import golden_util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_util.BARCODE_CASES)
def test_barcode_pipeline_on_engine_matches_reference_golden(name, tmp_path):
    from afterqc_b200.engine import Engine
    problems = golden_util.run_barcode_case(name, tmp_path, lambda p: Engine(p), 211)
    assert not problems, problems
