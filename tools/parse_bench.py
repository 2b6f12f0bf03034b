#!/usr/bin/env python
"""FASTQ text -> packed columns: the device parser (aqc_fastq_parse_device, text resident in HBM and from host memory) next
to the host parser (aqc_fastq_parse, one thread) on the same bytes.  Prints one JSON line.
usage: python tools/parse_bench.py [--mb 1024] [--reps 10]"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    import torch
    from afterqc_b200 import _abi, fastq_io, synth
    from afterqc_b200.engine import Engine
    import test_parse_device as T
    unit = T.fastq_text(synth.generate("pe150", 20000), 1)
    reps = (a.mb << 20) // len(unit) + 1
    text = np.tile(np.frombuffer(unit, dtype=np.uint8), reps)
    eng = Engine(_abi.Params.defaults())
    L = eng._L
    dev = C.c_void_p()
    eng._check(L.aqc_device_alloc(eng._h, text.size + 64, C.byref(dev)))
    eng._check(L.aqc_memcpy_h2d(eng._h, dev, text.ctypes.data, text.size))
    p = eng.parse_fastq_resident(dev, text.size)          # warm-up, buffers grown
    n0 = eng.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        p = eng.parse_fastq_resident(dev, text.size)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.reps
    launches = (eng.launch_count() - n0) // a.reps
    pinned = torch.empty(text.size, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = text
    hp = pinned.numpy()
    eng.parse_fastq(hp)
    t0 = time.perf_counter()
    for _ in range(3):
        eng.parse_fastq(hp)
    dth = (time.perf_counter() - t0) / 3
    small = bytes(text[:min(text.size, 256 << 20)])
    t0 = time.perf_counter()
    rec, consumed, eof = fastq_io._parse_block(small, True)
    dtc = time.perf_counter() - t0
    print(json.dumps({
        "what": "FASTQ text -> packed base / quality columns + line table", "text_GB": text.size / 1e9, "records": p.n,
        "device_resident": {"GBps": text.size / dt / 1e9, "M_records_per_s": p.n / dt / 1e6, "ms": dt * 1e3, "launches": launches,
                            "note": "text in HBM; wall clock around the call (three small device-to-host reads inside)"},
        "device_from_pinned_host": {"GBps": text.size / dth / 1e9, "M_records_per_s": p.n / dth / 1e6, "ms": dth * 1e3},
        "host_parser_1_thread": {"GBps": len(small) / dtc / 1e9, "M_records_per_s": (len(rec.seqs.off) - 1) / dtc / 1e6,
                                 "sample_GB": len(small) / 1e9},
    }))
    eng.close()


if __name__ == "__main__":
    main()
