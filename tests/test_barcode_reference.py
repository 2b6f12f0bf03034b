"""Barcode (UMI) files: the unmodified reference vs this repo's host pipeline (native
barcode pre-pass + the oracle as the device).  Build container only."""
import os

import pytest

import barcode_cases
import refcmp
from afterqc_b200 import cli
from afterqc_b200.pipeline import seqFilter

pytestmark = pytest.mark.reference


def _run(d, paired, extra, oracle_lib, batch_records, stem="x_barcode"):
    from oracle import ref_loader
    def args(sub):
        a = ["-1", os.path.join(d, sub, stem + "_R1.fq")]
        if paired:
            a += ["-2", os.path.join(d, sub, stem + "_R2.fq")]
        return a + ["-g", os.path.join(d, sub, "good")] + list(extra)
    ropt = ref_loader.run_cli(args("ref"))
    assert ropt.barcode is True
    opts, _ = cli.parseCommand(args("new"))
    cli.normalize_options(opts)
    opts.barcode = True; opts.trim_front = 0; opts.trim_front2 = 0          # after.py:215-219 (cli.main does the same)
    sf = seqFilter(opts, backend_factory=lambda p: oracle_lib.Oracle(p), batch_records=batch_records)
    sf.run()
    a = refcmp.load_json(d, "ref", stem + "_R1.fq")
    b = refcmp.load_json(d, "new", stem + "_R1.fq")
    diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
    assert not diffs, diffs[:5]
    files = ["good/%s_R1.good.fq" % stem, "bad/%s_R1.bad.fq" % stem]
    if paired:
        files += ["good/%s_R2.good.fq" % stem, "bad/%s_R2.bad.fq" % stem]
    for f in files:
        assert open(os.path.join(d, "ref", f), "rb").read() == open(os.path.join(d, "new", f), "rb").read(), f
    return a


@pytest.mark.parametrize("batch_records", [1 << 18, 37])
def test_barcoded_pairs(tmp_path, oracle_lib, batch_records):
    d = str(tmp_path)
    r1, r2 = barcode_cases.make(600, 1)
    barcode_cases.write(d, ("ref", "new"), r1, r2)
    a = _run(d, True, [], oracle_lib, batch_records)
    s = a["afterqc_main_summary"]
    assert s["bad_reads_with_bad_barcode"] > 20 and s["good_reads"] > 200
    bad = open(os.path.join(d, "new", "bad", "x_barcode_R1.bad.fq")).read()
    assert "@BADBCD1" in bad and "@BADBCD2" in bad


def test_barcoded_pairs_qc_sample_gate_inside_the_file(tmp_path, oracle_lib):
    d = str(tmp_path)
    r1, r2 = barcode_cases.make(500, 2, colon=False)
    barcode_cases.write(d, ("ref", "new"), r1, r2)
    _run(d, True, ["--qc_sample", "120", "-f", "0", "-t", "3"], oracle_lib, 64)


def test_barcoded_single_end_strips_the_design_length(tmp_path, oracle_lib):
    d = str(tmp_path)
    r1, _ = barcode_cases.make(400, 3, paired=False)
    barcode_cases.write(d, ("ref", "new"), r1, None)
    a = _run(d, False, ["--barcode_length", "11", "--barcode_verify", "CAGT"], oracle_lib, 50)
    assert a["afterqc_main_summary"]["bad_reads_with_bad_barcode"] > 0
