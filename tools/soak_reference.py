"""Randomised differential soak: the UNMODIFIED reference (oracle/ref_loader, build container only) against this repo's
host pipeline with the oracle as the device, on random synthetic inputs x random command lines (trims, thresholds,
QC window, k, flags, gz, odd batch sizes, barcoded files).  Any difference in the JSON or in a good/bad/overlap file is
printed with the seed that reproduces it.  TEST INFRASTRUCTURE.

  python tools/soak_reference.py [--cases N] [--seed S]
"""
import argparse
import contextlib
import gzip
import io
import os
import random
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import barcode_cases            # noqa: E402
import refcmp                   # noqa: E402
from afterqc_b200 import cli, synth     # noqa: E402
from afterqc_b200.pipeline import seqFilter   # noqa: E402
from oracle import oracle as oracle_lib, ref_loader   # noqa: E402


def random_case(rng):
    c = {"barcode": rng.random() < 0.2}
    c["paired"] = rng.random() < 0.8
    c["gz"] = rng.random() < 0.25
    c["n"] = rng.choice([150, 400, 900, 1300, 2500])
    extra = []
    if rng.random() < 0.6:
        extra += ["-f", str(rng.choice([0, 0, 1, 3, 8])), "-t", str(rng.choice([0, 0, 2, 5, 11]))]
    if rng.random() < 0.3:
        extra += ["--trim_pair_same", "false"]
    if rng.random() < 0.4:
        extra += ["--qc_sample", str(rng.choice([0, 1, 50, 300, 1000, 2000]))]
    if rng.random() < 0.3:
        extra += ["--qc_kmer", str(rng.choice([3, 4, 5, 6, 7, 8]))]
    if rng.random() < 0.3:
        extra += ["-q", str(rng.choice([2, 10, 20, 30])), "-u", str(rng.choice([0, 5, 30, 100]))]
    if rng.random() < 0.3:
        extra += ["-p", str(rng.choice([8, 15, 25, 40])), "-a", str(rng.choice([0, 1, 3, 6]))]
    if rng.random() < 0.3:
        extra += ["-n", str(rng.choice([0, 1, 3, 10])), "-s", str(rng.choice([1, 20, 60, 90]))]
    for flag, p in (("--no_correction", 0.2), ("--mask_mismatch", 0.2), ("--no_overlap", 0.1), ("--qc_only", 0.1), ("-z", 0.1)):
        if rng.random() < p:
            extra.append(flag)
    if rng.random() < 0.25:
        extra += ["--store_overlap", "on"]
    if c["barcode"] and rng.random() < 0.4:
        extra += ["--barcode_length", str(rng.choice([8, 11, 12, 13])), "--barcode_verify", rng.choice(["CAGTA", "CAGT", "ACGTAC"])]
    c["extra"] = extra
    c["index"] = rng.choice([0, 0, 0, 0, 0, 1, 2]) if not c["barcode"] else 0      # 1: -7 file, 2: -7 and -5 files (quirk Q1)
    c["index_short"] = rng.random() < 0.3                                            # index file with fewer records than R1
    c["batch_records"] = rng.choice([1 << 18, 1 << 18, 97, 256, 1001])
    c["cfg"] = rng.choice(["pe150", "pe150_err3", "pe250", "pe150"]) if c["paired"] else "se100"
    c["jitter"] = rng.choice([0, 0, 17, 60])
    c["flavour"] = rng.choice(["synth"] * 6 + ["adversarial", "long"]) if c["paired"] and not c["barcode"] else "synth"
    return c


def run_case(c, seed, d):
    stem = "x_barcode" if c["barcode"] else "x"
    ext = ".fq.gz" if c["gz"] else ".fq"
    for sub in ("ref", "new"):
        os.makedirs(os.path.join(d, sub))
    if c["barcode"]:
        r1s, r2s = barcode_cases.make(c["n"], seed, L=random.Random(seed).choice([80, 100, 150]), paired=c["paired"], colon=seed % 3 != 0)
        for sub in ("ref", "new"):
            for m, recs in ((1, r1s), (2, r2s)):
                if recs is None:
                    continue
                p = os.path.join(d, sub, "%s_R%d%s" % (stem, m, ext))
                with (gzip.open(p, "wt") if c["gz"] else open(p, "w")) as f:
                    for nm, s, q in recs:
                        f.write("%s\n%s\n+\n%s\n" % (nm, s, q))
    else:
        if c.get("flavour") == "adversarial":
            import cases
            batch = cases.adversarial_batch(seed=seed)            # pairs on the decision boundaries of overlap / polyX / filters
        elif c.get("flavour") == "long":
            import cases
            batch = cases.long_read_batch(seed=seed, n=min(c["n"], 400))      # 200-1000 bp reads
        else:
            batch = synth.generate(c["cfg"], c["n"], seed=seed, len_jitter=c["jitter"])
        for sub in ("ref", "new"):
            synth.write_fastq(batch, os.path.join(d, sub, stem + "_R1" + ext), os.path.join(d, sub, stem + "_R2" + ext) if batch.paired else None)
    paired = c["paired"]
    if c["index"]:
        rng = random.Random(seed + 7)
        n_idx = c["n"] - (rng.randint(1, 40) if c["index_short"] else 0)
        for k in range(c["index"]):
            text = "".join("@idx%d:%d\n%s\n+\n%s\n" % (k, i, "".join(rng.choice("ACGT") for _ in range(8)), "IIIIIIII") for i in range(n_idx))
            for sub in ("ref", "new"):
                p = os.path.join(d, sub, "%s_I%d%s" % (stem, k + 1, ext))
                with (gzip.open(p, "wt") if c["gz"] else open(p, "w")) as f:
                    f.write(text)

    def args(sub):
        a = ["-1", os.path.join(d, sub, stem + "_R1" + ext)]
        if paired:
            a += ["-2", os.path.join(d, sub, stem + "_R2" + ext)]
        for k in range(c["index"]):
            a += [("-7", "-5")[k], os.path.join(d, sub, "%s_I%d%s" % (stem, k + 1, ext))]
        return a + ["-g", os.path.join(d, sub, "good")] + c["extra"]
    sink = io.StringIO()
    ref_err = new_err = None
    with contextlib.redirect_stdout(sink):
        try:
            ref_loader.run_cli(args("ref"))
        except BaseException as e:                 # the reference raises on some inputs (documented domain limits)
            ref_err = "%s: %s" % (type(e).__name__, e)
        try:
            opts, _ = cli.parseCommand(args("new"))
            cli.normalize_options(opts)
            if opts.barcode_flag in opts.read1_file and cli.parseBool(opts.barcode):
                opts.barcode = True; opts.trim_front = 0; opts.trim_front2 = 0
            else:
                opts.barcode = False
            seqFilter(opts, backend_factory=lambda p: oracle_lib.Oracle(p), batch_records=c["batch_records"]).run()
        except BaseException as e:
            new_err = "%s: %s" % (type(e).__name__, e)
    r1name = stem + "_R1" + ext
    note = "ok"
    if ref_err and new_err:
        return "both_raised", (ref_err, new_err)           # e.g. reads of 1-4 bases reaching statRead, SE + --store_overlap
    if new_err:
        return "only_new_raised", new_err
    if ref_err:
        # the reference's HTML stage raises on empty postfilter statistics (max() of an empty list,
        # qualitycontrol.py:234) AFTER the JSON and the read files are complete: those still have to match
        if not os.path.exists(os.path.join(d, "ref", "QC", r1name + ".json")):
            return "only_ref_raised", ref_err
        note = "ok_ref_report_crashed"
    a, b = refcmp.load_json(d, "ref", r1name), refcmp.load_json(d, "new", r1name)
    diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
    if diffs:
        return "json", diffs[:4]
    for sub in ("good", "bad", "overlap"):
        pr = os.path.join(d, "ref", sub)
        if not os.path.isdir(pr):
            continue
        for fn in sorted(os.listdir(pr)):
            pn = os.path.join(d, "new", sub, fn)
            if not os.path.exists(pn):
                return "missing", sub + "/" + fn
            op = gzip.open if fn.endswith(".gz") else open
            with op(os.path.join(pr, fn), "rb") as f1, op(pn, "rb") as f2:
                if f1.read() != f2.read():
                    return "file", sub + "/" + fn
    return note, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=50)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--replay", type=int, nargs="*", help="case seeds (as printed) to run again")
    a = ap.parse_args()
    assert ref_loader.available(), "needs /root/reference"
    tally = {}
    for seed in (a.replay if a.replay else [a.seed * 100000 + i for i in range(a.cases)]):
        c = random_case(random.Random(seed))
        d = tempfile.mkdtemp(prefix="aqc_soak_")
        try:
            kind, info = run_case(c, seed, d)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        tally[kind] = tally.get(kind, 0) + 1
        if not kind.startswith("ok") and kind != "both_raised":
            print("case seed=%d %s -> %s %s" % (seed, c, kind, info), flush=True)
    print("soak:", tally)


if __name__ == "__main__":
    main()
