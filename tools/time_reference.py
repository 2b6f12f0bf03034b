#!/usr/bin/env python
"""Time the reference's OWN Python implementation (unmodified source under oracle/ref_loader.py, CPython 3.12) on a slice of
every bench config: one process on one core, and directory mode = one process per R1 file on all cores (after.py:168-171;
BASELINE.md section 3).  The reference is Python 2 and cannot travel to the GPU box, so this runs in the BUILD CONTAINER and
its result is committed as profiles/r02_reference_python_timing.json; bench.py reports it as `cpu_baseline_python` next to
the C port it times on the box itself.  TEST / MEASUREMENT INFRASTRUCTURE (uses oracle/).

  python tools/time_reference.py [--records 20000]
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {     # synth config, records of the slice, CLI flags of the bench config
    "pe150": ("pe150", 1.0, ["-f", "0", "-t", "0"]),
    "se100": ("se100", 1.0, ["-f", "0", "-t", "0", "--qc_sample", "0"]),
    "pe250_full": ("pe250", 0.4, ["--qc_sample", "0"]),
    "pe150_err3": ("pe150_err3", 1.0, ["-f", "0", "-t", "0"]),
}


def _job(args):
    from oracle import ref_loader
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        ref_loader.run_cli(args)


def run_one(argv):
    t0 = time.perf_counter()
    _job(argv)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=20000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_reference_python_timing.json"))
    a = ap.parse_args()
    from afterqc_b200 import synth
    from oracle import ref_loader
    assert ref_loader.available(), "needs /root/reference"
    nproc = len(os.sched_getaffinity(0))
    out = {}
    for name, (cfg, scale, flags) in CONFIGS.items():
        n = int(a.records * scale)
        work = tempfile.mkdtemp(prefix="aqc_reftime_")
        batch = synth.generate(cfg, n)
        # uncompressed in/out keeps zlib out of the measurement
        r1, r2 = os.path.join(work, "s_R1.fq"), os.path.join(work, "s_R2.fq") if batch.paired else None
        synth.write_fastq(batch, r1, r2)
        base = ["-1", r1] + (["-2", r2] if r2 else []) + ["-g", os.path.join(work, "good")] + flags
        one = run_one(base)
        # directory mode: the slice split into nproc R1/R2 file pairs, one process each
        d = os.path.join(work, "dir"); os.makedirs(d)
        per = (n + nproc - 1) // nproc
        jobs = []
        for i in range(nproc):
            part = batch.slice(i * per, min(n, (i + 1) * per))
            if part.n == 0:
                continue
            p1, p2 = os.path.join(d, "p%d_R1.fq" % i), os.path.join(d, "p%d_R2.fq" % i) if batch.paired else None
            synth.write_fastq(part, p1, p2)
            jobs.append(["-1", p1] + (["-2", p2] if p2 else []) + ["-g", os.path.join(d, "good")] + flags)
        t0 = time.perf_counter()
        procs = [mp.Process(target=_job, args=(j,)) for j in jobs]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        allc = time.perf_counter() - t0
        shutil.rmtree(work, ignore_errors=True)
        unit = "read-pairs/s" if batch.paired else "reads/s"
        out[name] = {"kind": "reference", "interpreter": "CPython %s + oracle/ref_loader.py (the reference is Python 2; PyPy is not installable here)" % platform.python_version(),
                     "where": "build container (%d cores visible), NOT the GPU box" % nproc, "records": n, "flags": flags, "unit": unit,
                     "one_core": {"seconds": round(one, 2), "value": round(n / one, 1)},
                     "all_cores_directory_mode": {"processes": len(jobs), "seconds": round(allc, 2), "value": round(n / allc, 1)}}
        print(name, out[name]["one_core"], out[name]["all_cores_directory_mode"], flush=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
