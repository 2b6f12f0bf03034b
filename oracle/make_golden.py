#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ by running the UNMODIFIED reference
(oracle/ref_loader.py over /root/reference) in the build container.  TEST INFRASTRUCTURE.

  tests/golden/<case>/x_R1.fq.gz [x_R2.fq.gz]   seeded synthetic input (afterqc_b200.synth)
  tests/golden/<case>/expected.json             the reference's QC/<R1>.json + sha256 of its decompressed outputs
  tests/golden/ops_adversarial.json             util.overlap / hasPolyX / lowQualityNum / nNumber of the reference on
                                                the adversarial pairs of tests/cases.py for three parameter sets
Run:  python oracle/make_golden.py
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from afterqc_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, config, n, len_jitter, extra CLI args
    ("pe150_default", "pe150", 1400, 0, []),
    ("pe150_err3_mask_overlap", "pe150_err3", 1200, 0, ["-f", "0", "-t", "0", "--mask_mismatch", "--store_overlap", "on"]),
    ("pe150_jitter_nocorr", "pe150_err3", 1200, 40, ["-f", "2", "-t", "3", "--no_correction", "--qc_sample", "500"]),
    ("se100_f0", "se100", 1500, 0, ["-f", "0", "-t", "0", "--qc_sample", "0"]),
    ("pe250_k5_strict", "pe250", 700, 25, ["--qc_sample", "0", "--qc_kmer", "5", "-p", "20", "-a", "1", "-q", "20", "-u", "30", "-n", "1", "-s", "60"]),
    ("pe150_small_head_fallback", "pe150", 600, 0, []),
    # BASELINE configs[0]: the 250 NextSeq pairs the reference ships (testdata/R1.fq.gz + R2.fq.gz, copied as data), default run:
    # auto-trim resolves to 15 / 7, 236 good, BADPOL 11 + BADNCT 1 + BADDIFF 2
    ("testdata", "reference:testdata", 250, 0, []),
]


def sha(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


BARCODE_GOLD = os.path.join(ROOT, "tests", "golden_barcode")
BARCODE_CASES = [
    # name, pairs, seed, paired, names with ':', extra CLI args
    ("pe100_umi12", 900, 11, True, True, []),
    ("pe100_umi12_gate_nocolon", 700, 12, True, False, ["--qc_sample", "150", "-t", "3", "--store_overlap", "on"]),
    ("se100_umi11", 600, 13, False, True, ["--barcode_length", "11", "--barcode_verify", "CAGT"]),
]


def make_barcode_golden():
    """tests/golden_barcode/<case>/: barcoded (UMI) inputs of tests/barcode_cases.py + the reference's outputs."""
    import barcode_cases
    os.makedirs(BARCODE_GOLD, exist_ok=True)
    for name, n, seed, paired, colon, extra in BARCODE_CASES:
        d = os.path.join(BARCODE_GOLD, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        r1s, r2s = barcode_cases.make(n, seed, paired=paired, colon=colon)
        paths = []
        for m, recs in ((1, r1s), (2, r2s)):
            if recs is None:
                continue
            p = os.path.join(d, "x_barcode_R%d.fq.gz" % m)
            with gzip.open(p, "wb", compresslevel=6) as f:
                for nm, s, q in recs:
                    f.write(("%s\n%s\n+\n%s\n" % (nm, s, q)).encode())
            paths.append(p)
        work = tempfile.mkdtemp()
        args = ["-1", paths[0]] + (["-2", paths[1]] if paired else []) + ["-g", os.path.join(work, "good")] + extra
        opt = ref_loader.run_cli(args)
        assert opt.barcode is True
        with open(os.path.join(work, "QC", "x_barcode_R1.fq.gz.json")) as f:
            stat = json.load(f)
        for k in ("read1_file", "read2_file", "good_output_folder"):
            stat["command"][k] = None
        outs = {}
        for sub in ("good", "bad", "overlap"):
            p = os.path.join(work, sub)
            if os.path.isdir(p):
                for fn in sorted(os.listdir(p)):
                    outs[sub + "/" + fn] = sha(os.path.join(p, fn))
        with open(os.path.join(d, "expected.json"), "w") as f:
            json.dump({"args": extra, "stat": stat, "outputs_sha256": outs}, f, sort_keys=True, indent=1)
        shutil.rmtree(work)
        sm = stat["afterqc_main_summary"]
        print(name, sm["good_reads"], "/", sm["total_reads"], "bad barcode", sm["bad_reads_with_bad_barcode"], list(outs))


def main():
    assert ref_loader.available(), "needs /root/reference"
    if "--barcode-only" in sys.argv:
        return make_barcode_golden()
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    if not only:
        make_barcode_golden()
    os.makedirs(GOLD, exist_ok=True)
    for name, cfg, n, jitter, extra in CASES:
        if only and name not in only:
            continue
        d = os.path.join(GOLD, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        r1 = os.path.join(d, "x_R1.fq.gz")
        if cfg == "reference:testdata":
            r2 = os.path.join(d, "x_R2.fq.gz")
            shutil.copy(os.path.join(ref_loader.REFERENCE_DIR, "testdata", "R1.fq.gz"), r1)
            shutil.copy(os.path.join(ref_loader.REFERENCE_DIR, "testdata", "R2.fq.gz"), r2)
        else:
            batch = synth.generate(cfg, n, len_jitter=jitter)
            r2 = os.path.join(d, "x_R2.fq.gz") if batch.paired else None
            synth.write_fastq(batch, r1, r2)
        work = tempfile.mkdtemp()
        args = ["-1", r1] + (["-2", r2] if r2 else []) + ["-g", os.path.join(work, "good")] + extra
        ref_loader.run_cli(args)
        with open(os.path.join(work, "QC", "x_R1.fq.gz.json")) as f:
            stat = json.load(f)
        for k in ("read1_file", "read2_file", "good_output_folder"):
            stat["command"][k] = None          # machine-specific paths
        outs = {}
        for sub in ("good", "bad", "overlap"):
            p = os.path.join(work, sub)
            if os.path.isdir(p):
                for fn in sorted(os.listdir(p)):
                    outs[sub + "/" + fn] = sha(os.path.join(p, fn))
        with open(os.path.join(d, "expected.json"), "w") as f:
            json.dump({"args": extra, "stat": stat, "outputs_sha256": outs}, f, sort_keys=True, indent=1)
        shutil.rmtree(work)
        print(name, stat["afterqc_main_summary"]["good_reads"], "/", stat["afterqc_main_summary"]["total_reads"], list(outs))

    if only:
        return
    # operator-level goldens on the adversarial pairs
    import cases
    mods = ref_loader.load()
    util, pre = mods["util"], mods["preprocesser"]
    batch = cases.adversarial_batch()
    out = {"n": batch.n, "sets": {}}
    for pname in ("default_f0", "strict", "poly_wide"):
        p = cases.make_params(pname)
        rows = []
        for i in range(batch.n):
            s1, q1 = batch.read(1, i)
            s2, q2 = batch.read(2, i)
            ov = util.overlap(s1, s2)
            px1 = pre.hasPolyX(s1, p.poly_size_limit, p.allow_mismatch_in_poly)
            px2 = pre.hasPolyX(s2, p.poly_size_limit, p.allow_mismatch_in_poly)
            rows.append([ov[0], ov[1], ov[2], ord(px1) if px1 else 0, ord(px2) if px2 else 0,
                         pre.lowQualityNum(["", s1, "+", q1], p.qualified_quality_phred),
                         pre.lowQualityNum(["", s2, "+", q2], p.qualified_quality_phred),
                         pre.nNumber(["", s1, "+", q1]), pre.nNumber(["", s2, "+", q2])])
        out["sets"][pname] = rows
    with open(os.path.join(GOLD, "ops_adversarial.json"), "w") as f:
        json.dump(out, f)
    print("ops_adversarial", batch.n)


if __name__ == "__main__":
    main()
