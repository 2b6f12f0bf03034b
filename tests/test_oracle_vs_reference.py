"""Live differential test against the UNMODIFIED reference (build container only; skipped elsewhere)."""
import os
import shutil

import pytest

import refcmp
from afterqc_b200 import synth

pytestmark = pytest.mark.reference

CASES = [
    ("pe150_f0", "pe150", 1500, 0, ["-f", "0", "-t", "0"]),
    ("pe150_tps_false", "pe150", 1300, 0, ["--trim_pair_same", "false"]),
    ("pe150_noov", "pe150", 1100, 0, ["--no_overlap"]),
    ("pe150_qconly", "pe150", 1800, 0, ["--qc_only", "--qc_sample", "1200"]),
    ("pe150_window", "pe150", 1250, 30, ["--qc_sample", "100", "-z"]),
]


@pytest.mark.parametrize("name,cfg,n,jitter,extra", CASES)
def test_pipeline_on_oracle_equals_reference(name, cfg, n, jitter, extra, tmp_path, oracle_lib, capsys):
    from oracle import ref_loader
    d = str(tmp_path)
    batch = synth.generate(cfg, n, len_jitter=jitter)
    refcmp.prepare_case(d, batch)
    ref_loader.run_cli(refcmp.cli_args(d, "ref", batch.paired, extra))
    refcmp.run_ours(d, "new", batch.paired, extra, lambda p: oracle_lib.Oracle(p))
    a, b = refcmp.load_json(d, "ref"), refcmp.load_json(d, "new")
    diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
    assert not diffs, diffs[:5]
    if "--qc_only" not in extra:
        import gzip
        for f in refcmp.output_files(batch.paired, extra):
            fa, fb = os.path.join(d, "ref", f), os.path.join(d, "new", f)
            if "-z" in extra:
                assert gzip.open(fa + ".gz").read() == gzip.open(fb + ".gz").read(), f
            else:
                assert open(fa, "rb").read() == open(fb, "rb").read(), f


def test_testdata_matches_survey_numbers(tmp_path, oracle_lib):
    """The repo-shipped 250-pair sample (BASELINE configs[0]); expected counters from SURVEY.md section 4."""
    d = str(tmp_path)
    for f in ("R1.fq.gz", "R2.fq.gz"):
        shutil.copy(os.path.join("/root/reference/testdata", f), os.path.join(d, f))
    from afterqc_b200 import cli
    from afterqc_b200.pipeline import seqFilter
    opts, _ = cli.parseCommand(["-1", os.path.join(d, "R1.fq.gz"), "-2", os.path.join(d, "R2.fq.gz"), "-g", os.path.join(d, "good")])
    cli.normalize_options(opts); opts.barcode = False
    stat = seqFilter(opts, backend_factory=lambda p: oracle_lib.Oracle(p)).run()
    s = stat["afterqc_main_summary"]
    assert (s["total_reads"], s["good_reads"], s["bad_reads_with_polyX"], s["bad_reads_with_too_many_N"], s["bad_reads_with_bad_overlap"]) == (250, 236, 11, 1, 2)
    assert (s["total_bases"], s["good_bases"], s["readlen"]) == (36819, 29271, 151)
    o = stat["afterqc_overlap"]
    assert (o["overlapped_pairs"], o["corrected_reads"], o["corrected_bases"], o["skipped_correction_bases"]) == (183, 26, 32, 10)
    assert (o["trimmed_adapter_reads"], o["trimmed_adapter_bases"]) == (35, 506)
    assert o["edit_distance_histogram"] == [208, 21, 5, 2, 1, 0, 1, 0, 0, 0]
    assert o["error_rate"] == 0.0011543021151806327
    assert (opts.trim_front, opts.trim_tail) == (15, 7)
