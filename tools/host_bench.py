"""Host-side ceiling of the streaming CLI pipeline (FASTQ ingest/egress).

Runs afterqc_b200.pipeline.seqFilter on a synthetic FASTQ pair with a STUB device backend whose filter/stat calls
return at once (every pair "good", untouched), so the time left is what the host spends around the device calls:
inflate/parse -> pack -> (device) -> emit -> (deflate) -> write.  It is a tool for sizing the host pipeline on a box
without a GPU; it is not a product path and its numbers are not bench.py numbers.

  python tools/host_bench.py [--pairs N] [--gz] [--config pe150]
"""
import argparse
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from afterqc_b200 import _abi, cli, synth                 # noqa: E402
from afterqc_b200.pipeline import seqFilter               # noqa: E402


class StubBackend:
    """No device: filter_pairs marks every pair GOOD with its full length."""

    def __init__(self, params):
        self.t_filter = 0.0

    def set_params(self, p): pass
    def stat_reads(self, *a, **k): pass
    def close(self): pass

    def filter_pairs(self, batch):
        res = np.zeros(batch.n, dtype=_abi.RESULT_DTYPE)
        res["len1"] = np.diff(batch.off1.astype(np.int64))
        if batch.off2 is not None:
            res["len2"] = np.diff(batch.off2.astype(np.int64))
        return res

    def counters(self):
        return np.zeros(_abi.C_TOTAL, dtype=np.uint64)

    def qc(self, slot):
        return np.zeros((), dtype=_abi.QC_DTYPE)

    def kmers(self, slot):
        z = np.zeros(0, dtype=np.uint64)
        return (np.zeros(4 ** 8, dtype=np.uint64), np.full(4 ** 8, _abi.KMER_NEVER, dtype=np.uint64), z, z, z)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1000000)
    ap.add_argument("--config", default="pe150")
    ap.add_argument("--gz", action="store_true")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else None, help="scratch directory (tmpfs by default: disks vary)")
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix="aqc_host_bench_", dir=a.dir)
    try:
        ext = ".fq.gz" if a.gz else ".fq"
        r1, r2 = os.path.join(d, "s_R1" + ext), os.path.join(d, "s_R2" + ext)
        t = time.time()
        unit = min(a.pairs, 100000)                  # one synthetic unit, repeated (names repeat; irrelevant here)
        reps = max(1, a.pairs // unit)
        a.pairs = unit * reps
        u1, u2 = os.path.join(d, "u_R1.fq"), os.path.join(d, "u_R2.fq")
        synth.write_fastq(synth.generate(a.config, unit, seed=7), u1, u2)
        import gzip
        for u, r in ((u1, r1), (u2, r2)):
            text = open(u, "rb").read()
            with (gzip.open(r, "wb", compresslevel=1) if a.gz else open(r, "wb")) as f:
                for _ in range(reps):
                    f.write(text)
            os.unlink(u)
        in_bytes = os.path.getsize(r1) + os.path.getsize(r2)
        print("wrote %d pairs (%.1f MB on disk) in %.1fs" % (a.pairs, in_bytes / 1e6, time.time() - t), file=sys.stderr)
        opts, _ = cli.parseCommand(["-1", r1, "-2", r2, "-f", "0", "-t", "0", "-g", os.path.join(d, "good")])     # all outputs inside the scratch dir
        cli.normalize_options(opts); opts.barcode = False
        sf = seqFilter(opts, backend_factory=StubBackend)
        t = time.time()
        so = sys.stdout
        sys.stdout = open(os.devnull, "w")
        try:
            sf.run()
        finally:
            sys.stdout = so
        dt = time.time() - t
        print('{"tool": "host_bench", "pairs": %d, "gz": %s, "seconds": %.3f, "pairs_per_s": %.0f, "threads_note": "stub device backend"}'
              % (a.pairs, "true" if a.gz else "false", dt, a.pairs / dt))
    finally:
        if not a.keep:
            shutil.rmtree(d)


if __name__ == "__main__":
    main()
