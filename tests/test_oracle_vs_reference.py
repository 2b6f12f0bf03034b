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


@pytest.mark.parametrize("with_i2", [False, True])
def test_index_files_pass_through(with_i2, tmp_path, oracle_lib):
    """-7/-5 index reads are carried to good/bad(/overlap) untouched; an index2 file switches on the R2 base counting
    quirk Q1 (preprocesser.py:426-431,622-623); the loop ends at the shortest file."""
    from oracle import ref_loader
    from afterqc_b200 import cli
    from afterqc_b200.pipeline import seqFilter
    import random
    d = str(tmp_path)
    batch = synth.generate("pe150", 1300)
    refcmp.prepare_case(d, batch)
    rng = random.Random(5)
    n_idx = {"I1": 1300, "I2": 1250}           # I2 is shorter: the loop must stop there
    for sub in ("ref", "new"):
        rng = random.Random(5)
        for tag in ("I1", "I2"):
            with open(os.path.join(d, sub, "x_%s.fq" % tag), "w") as f:
                for i in range(n_idx[tag]):
                    f.write("@SYN:1:FC:1:1101:%d:%d 1:N:0:A\n%s\n+\n%s\n" % (i, i, "".join(rng.choice("ACGT") for _ in range(8)), "I" * 8))
    extra = ["--store_overlap", "on", "-7", "{sub}/x_I1.fq"] + (["-5", "{sub}/x_I2.fq"] if with_i2 else [])

    def args(sub):
        return refcmp.cli_args(d, sub, True, [a.format(sub=os.path.join(d, sub)) for a in extra])
    ref_loader.run_cli(args("ref"))
    opts, _ = cli.parseCommand(args("new"))
    cli.normalize_options(opts); opts.barcode = False
    seqFilter(opts, backend_factory=lambda p: oracle_lib.Oracle(p)).run()
    a, b = refcmp.load_json(d, "ref"), refcmp.load_json(d, "new")
    diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
    assert not diffs, diffs[:5]
    files = refcmp.output_files(True, ["--store_overlap"]) + ["good/x_I1.good.fq", "bad/x_I1.bad.fq", "overlap/x_I1.overlap.fq"]
    if with_i2:
        files += ["good/x_I2.good.fq", "bad/x_I2.bad.fq", "overlap/x_I2.overlap.fq"]
    for f in files:
        assert open(os.path.join(d, "ref", f), "rb").read() == open(os.path.join(d, "new", f), "rb").read(), f
    assert a["afterqc_main_summary"]["total_reads"] == (1250 if with_i2 else 1300)
