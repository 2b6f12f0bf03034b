import gzip
import hashlib
import json
import os
import shutil

import refcmp
from afterqc_b200 import cli
from afterqc_b200.pipeline import seqFilter

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(d for d in os.listdir(GOLD) if os.path.isdir(os.path.join(GOLD, d)))


def sha(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def run_case(name, tmp, backend_factory):
    """Run this repo's pipeline on a golden case; return list of discrepancies vs the reference's outputs."""
    src = os.path.join(GOLD, name)
    with open(os.path.join(src, "expected.json")) as f:
        exp = json.load(f)
    work = os.path.join(str(tmp), name)
    os.makedirs(work)
    paired = os.path.exists(os.path.join(src, "x_R2.fq.gz"))
    for fn in ("x_R1.fq.gz", "x_R2.fq.gz"):
        if os.path.exists(os.path.join(src, fn)):
            shutil.copy(os.path.join(src, fn), os.path.join(work, fn))
    args = ["-1", os.path.join(work, "x_R1.fq.gz")] + (["-2", os.path.join(work, "x_R2.fq.gz")] if paired else [])
    args += ["-g", os.path.join(work, "good")] + exp["args"]
    opts, _ = cli.parseCommand(args)
    cli.normalize_options(opts); opts.barcode = False
    seqFilter(opts, backend_factory=backend_factory).run()
    with open(os.path.join(work, "QC", "x_R1.fq.gz.json")) as f:
        got = json.load(f)
    for k in ("read1_file", "read2_file", "good_output_folder"):
        got["command"][k] = None
    problems = [str(x) for x in refcmp.json_diff(exp["stat"], got)][:10]
    for rel, digest in exp["outputs_sha256"].items():
        p = os.path.join(work, rel)
        if not os.path.exists(p):
            problems.append("missing output " + rel)
        elif sha(p) != digest:
            problems.append("content differs: " + rel)
    return problems


BARCODE_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_barcode")
BARCODE_CASES = sorted(d for d in os.listdir(BARCODE_GOLD) if os.path.isdir(os.path.join(BARCODE_GOLD, d))) if os.path.isdir(BARCODE_GOLD) else []


def run_barcode_case(name, tmp, backend_factory, batch_records=1 << 18):
    """Barcoded (UMI) golden case: file names carry the barcode flag, so the CLI rules of after.py:215-221 switch the
    barcode pre-pass on (oracle/make_golden.py made the expected outputs with the reference)."""
    src = os.path.join(BARCODE_GOLD, name)
    with open(os.path.join(src, "expected.json")) as f:
        exp = json.load(f)
    work = os.path.join(str(tmp), name)
    os.makedirs(work)
    r1 = os.path.join(work, "x_barcode_R1.fq.gz")
    r2 = os.path.join(work, "x_barcode_R2.fq.gz")
    shutil.copy(os.path.join(src, "x_barcode_R1.fq.gz"), r1)
    paired = os.path.exists(os.path.join(src, "x_barcode_R2.fq.gz"))
    if paired:
        shutil.copy(os.path.join(src, "x_barcode_R2.fq.gz"), r2)
    args = ["-1", r1] + (["-2", r2] if paired else []) + ["-g", os.path.join(work, "good")] + exp["args"]
    opts, _ = cli.parseCommand(args)
    cli.normalize_options(opts)
    assert opts.barcode_flag in opts.read1_file and cli.parseBool(opts.barcode)
    opts.barcode = True; opts.trim_front = 0; opts.trim_front2 = 0
    seqFilter(opts, backend_factory=backend_factory, batch_records=batch_records).run()
    with open(os.path.join(work, "QC", "x_barcode_R1.fq.gz.json")) as f:
        got = json.load(f)
    for k in ("read1_file", "read2_file", "good_output_folder"):
        got["command"][k] = None
    problems = [str(x) for x in refcmp.json_diff(exp["stat"], got)][:10]
    for rel, digest in exp["outputs_sha256"].items():
        p = os.path.join(work, rel)
        if not os.path.exists(p):
            problems.append("missing output " + rel)
        elif sha(p) != digest:
            problems.append("content differs: " + rel)
    return problems
