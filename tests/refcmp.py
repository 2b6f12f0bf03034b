"""Helpers for differential runs: reference (via oracle/ref_loader) vs this repo's host pipeline on a backend."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT,):
    if p not in sys.path:
        sys.path.insert(0, p)


def json_diff(x, y, path="", out=None):
    out = [] if out is None else out
    if type(x) != type(y):
        out.append(("TYPE", path)); return out
    if isinstance(x, dict):
        for k in sorted(set(x) | set(y)):
            if k not in x or k not in y:
                out.append(("KEY", path, k)); continue
            json_diff(x[k], y[k], path + "/" + k, out)
    elif isinstance(x, list):
        if len(x) != len(y):
            out.append(("LEN", path, len(x), len(y))); return out
        for i, (p, q) in enumerate(zip(x, y)):
            json_diff(p, q, path + "/%d" % i, out)
    elif x != y:
        out.append(("DIFF", path, x, y))
    return out


def cli_args(d, sub, paired, extra):
    a = ["-1", os.path.join(d, sub, "x_R1.fq")]
    if paired:
        a += ["-2", os.path.join(d, sub, "x_R2.fq")]
    return a + ["-g", os.path.join(d, sub, "good")] + list(extra)


def run_ours(d, sub, paired, extra, backend_factory):
    from afterqc_b200 import cli
    from afterqc_b200.pipeline import seqFilter
    opts, _ = cli.parseCommand(cli_args(d, sub, paired, extra))
    cli.normalize_options(opts)
    opts.barcode = False
    sf = seqFilter(opts, backend_factory=backend_factory)
    sf.run()
    return sf


def output_files(paired, extra):
    files = ["good/x_R1.good.fq", "bad/x_R1.bad.fq"]
    if paired:
        files += ["good/x_R2.good.fq", "bad/x_R2.bad.fq"]
    if "--store_overlap" in extra and paired:
        files += ["overlap/x_R1.overlap.fq", "overlap/x_R2.overlap.fq"]
    return files


def prepare_case(d, batch, subs=("ref", "new")):
    from afterqc_b200 import synth
    shutil.rmtree(d, ignore_errors=True)
    for sub in subs:
        os.makedirs(os.path.join(d, sub))
        synth.write_fastq(batch, os.path.join(d, sub, "x_R1.fq"), os.path.join(d, sub, "x_R2.fq") if batch.paired else None)


def load_json(d, sub, r1_name="x_R1.fq"):
    with open(os.path.join(d, sub, "QC", r1_name + ".json")) as f:
        return json.load(f)
