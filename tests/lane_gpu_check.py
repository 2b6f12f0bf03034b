"""GPU parity of the lane-per-pair filter kernel (aqc_params.filter_kernel = 2), run as a script so that callers
(tests/test_gpu_zzz_lane_kernel.py, bench.py) can put it in a child process with a timeout: the kernel was written in a
round without GPU time left and had only been run under the SIMT emulator (tests/emu) when it was committed.

  python tests/lane_gpu_check.py parity          lane engine vs the CPU oracle, bit-exact, on the parity-test batches
  python tests/lane_gpu_check.py full [pairs] [lane|lane2|warp_st2|lane_st2|lane_st3|...][,more] [noinplace]
                                                 that kernel vs the warp-per-pair kernel on the bench workload
                                                 (HBM-resident entry); prints one JSON line with both kernel times;
                                                 *_st2 = with aqc_params.stat_kernel = 2 (also compares aqc_stat_reads)
  python tests/lane_gpu_check.py pack            the 2-bit base transport of the host-buffer entry (AQC_BATCH_PACK_BASES) vs the oracle
Exit code 0 = identical everywhere.
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def kernel_ids(candidate):
    """'lane' | 'lane2' | 'warp', optionally with '_st2' / '_st3' (aqc_params.stat_kernel = 2 / 3: statRead with one lane per read, in the
    filter kernel / in its own launch after it)"""
    from afterqc_b200 import _abi
    base, _, st2 = candidate.partition("_")
    fk = {"warp": _abi.KERNEL_WARP, "lane": _abi.KERNEL_LANE, "lane2": _abi.KERNEL_LANE2}[base]
    if st2 not in ("", "st2", "st3"):
        raise SystemExit("unknown candidate %r" % candidate)
    return fk, {"": _abi.STAT_DEFAULT, "st2": _abi.STAT_LANE, "st3": _abi.STAT_LANE_POST}[st2]


def parity(candidate="lane"):
    import cases
    import compare
    from afterqc_b200 import _abi
    from afterqc_b200.engine import Engine
    from oracle import oracle
    oracle.build()
    fk, sk = kernel_ids(candidate)
    batches = {
        "adversarial": (cases.adversarial_batch(), list(cases.PARAM_SETS)),
        "pe150": (cases.synthetic("pe150", 20000), list(cases.PARAM_SETS)),
        "pe150_err3": (cases.synthetic("pe150_err3", 12000), ["default_f0", "trim", "strict"]),
        "pe150_jitter": (cases.synthetic("pe150", 8000, len_jitter=60), ["default_f0", "trim", "mask"]),
        "pe250": (cases.synthetic("pe250", 6000), ["default_f0", "trim", "poly_wide"]),
        "long": (cases.long_read_batch(), ["default_f0"]),           # > 256 bases: must fall through to pair_kernel
    }
    n_cases = 0
    for bname, (batch, pnames) in batches.items():
        for pname in pnames:
            p = cases.make_params(pname)
            p.filter_kernel = fk; p.stat_kernel = sk
            orc, eng = oracle.Oracle(p), Engine(p)
            a = orc.filter_pairs(batch)
            b = eng.filter_pairs(batch)
            what = "%s %s/%s" % (candidate, bname, pname)
            compare.assert_records_equal(batch, a, b, what)
            compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), what)
            # the HBM-resident entry
            eng.reset()
            d = eng.upload(batch)
            eng.filter_pairs(d)
            compare.assert_records_equal(batch, a, eng.fetch_results(d), what + " resident")
            compare.assert_counters_equal(orc.counters(), eng.counters(), what + " resident")
            if sk and pname == pnames[0]:                 # the prefilter statistics entry (stat_lane_kernel), window inside the batch
                lo, hi = batch.n // 10, batch.n - batch.n // 7
                for be, bb in ((orc, batch), (eng, d)):
                    be.stat_reads(bb, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=lo, stat_hi=hi, order_base=3)
                compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), what + " prefilter statistics")
            d.free(); orc.close(); eng.close()
            n_cases += 1
    se = cases.synthetic("se100", 30000)
    for pname in ("default_f0", "trim", "loose"):
        p = cases.make_params(pname, paired=False)
        p.filter_kernel = fk; p.stat_kernel = sk
        orc, eng = oracle.Oracle(p), Engine(p)
        compare.assert_records_equal(se, orc.filter_pairs(se), eng.filter_pairs(se), "lane se100 %s" % pname)
        compare.compare_backends(orc, eng, (_abi.QC_R1_POST,), "lane se100 %s" % pname)
        orc.close(); eng.close()
        n_cases += 1
    print("%s kernel parity ok: %d cases" % (candidate, n_cases))


def full(pairs, candidates="lane", try_in_place=True):
    """pair_kernel (+ stat_read) once, then every candidate of the comma-separated list on the bench workload, resident in HBM:
    records, counters, the postfilter QC slots and -- for the *_st2 / *_st3 candidates -- the prefilter slots after aqc_stat_reads
    must match.  One JSON line per candidate as soon as its verdict is known (callers take the LAST line of a candidate)."""
    import torch
    from afterqc_b200 import _abi, synth
    from afterqc_b200.batch import PackedBatch
    from afterqc_b200.engine import Engine
    t = synth.generate_device("pe150", pairs, device="cuda")
    host = PackedBatch(t["seq1"].cpu().numpy(), t["qual1"].cpu().numpy(), t["off1"].cpu().numpy().astype(np.uint32),
                       t["seq2"].cpu().numpy(), t["qual2"].cpu().numpy(), t["off2"].cpu().numpy().astype(np.uint32))
    del t
    torch.cuda.empty_cache()
    qs = max(1000, pairs // 50)                                     # bench.py's mix: --qc_sample 200000 on 10 M pairs = 2 %
    s_lo, s_hi = 999, 999 + min(qs, max(1, pairs // 2))            # its prefilter window ...
    window = host.slice((s_lo // 4) * 4, s_hi)                     # ... and, as there, a batch of just those records
    post = (_abi.QC_R1_POST, _abi.QC_R2_POST)
    pre = (_abi.QC_R1_PRE, _abi.QC_R2_PRE)

    def run(k, sk, with_stat):
        eng = Engine(_abi.Params.defaults(filter_kernel=k, stat_kernel=sk, qc_sample=qs))
        d = eng.upload(host)
        r = {"eng": eng, "d": d}
        if with_stat:                                        # aqc_stat_reads: pair_kernel<MODE_STAT> or stat_lane_kernel
            dw = eng.upload(window)
            eng.stat_reads(dw, pre[0], pre[1], s_lo, s_hi, 0); eng.sync()
            eng.reset()
            eng.stat_reads(dw, pre[0], pre[1], s_lo, s_hi, 0); eng.sync()
            r["stat_ms"] = round(eng.last_kernel_ms(), 4)
            r["pre"] = ([eng.qc(x) for x in pre], [eng.kmers(x) for x in pre])
            dw.free()
        eng.filter_pairs(d); eng.sync()                      # warm-up
        eng.reset()
        eng.filter_pairs(d); eng.sync()
        r["ms"] = round(eng.last_kernel_ms(), 4)
        r["res"] = eng.fetch_results(d)
        r["cnt"] = eng.counters()
        r["post"] = ([eng.qc(x) for x in post], [eng.kmers(x) for x in post])
        return r

    def same_qc(a, b, what):
        for x, y in zip(a[0], b[0]):
            for f in x.dtype.names:
                assert np.array_equal(x[f], y[f]), "%s QC field %s differs" % (what, f)
        for x, y in zip(a[1], b[1]):
            for u, v in zip(x, y):
                assert np.array_equal(u, v), "%s k-mer tables differ" % what

    ref = run(_abi.KERNEL_WARP, _abi.STAT_DEFAULT, True)
    ref["d"].free(); ref["eng"].close()
    for candidate in [c for c in candidates.split(",") if c]:
        out = {"candidate": candidate, "pairs": pairs, "warp_ms": ref["ms"], "stat_warp_ms": ref["stat_ms"]}
        cand_id, cand_stat = kernel_ids(candidate)
        eng = d = None
        try:
            r = run(cand_id, cand_stat, bool(cand_stat))
            eng, d = r["eng"], r["d"]
            out["lane_ms"] = r["ms"]                         # key "lane_ms" = the candidate's filter launches
            if cand_stat:
                out["stat_ms"] = r["stat_ms"]
                same_qc(r["pre"], ref["pre"], "prefilter")
            assert r["res"].tobytes() == ref["res"].tobytes(), "records differ between the kernels"
            assert np.array_equal(r["cnt"], ref["cnt"]), "counters differ between the kernels"
            same_qc(r["post"], ref["post"], "postfilter")
            out["identical"] = True
        except AssertionError as e:
            out["identical"] = False
            out["why"] = str(e)[:200]
        # the verdict on the resident kernels stands whatever happens below (callers take the LAST line of a candidate they can parse)
        print(json.dumps(out), flush=True)
        if out["identical"] and cand_id != _abi.KERNEL_WARP and not cand_stat and try_in_place:
            # host-buffer entry, mate-2 qualities copied vs left in page-locked host memory (AQC_BATCH_QUAL2_IN_PLACE)
            try:
                import ctypes as C
                p = C.c_void_p()
                eng._check(eng._L.aqc_host_alloc(host.qual2.nbytes, C.byref(p)))
                pinned = np.frombuffer((C.c_uint8 * host.qual2.nbytes).from_address(p.value), dtype=np.uint8)
                pinned[:] = host.qual2
                h2 = PackedBatch(host.seq1, host.qual1, host.off1, host.seq2, pinned, host.off2)
                for key, kw in (("host_ms", {}), ("host_in_place_ms", {"qual2_in_place": True})):
                    eng.reset()
                    t0 = time.perf_counter()
                    rr = eng.filter_pairs(h2, **kw)
                    out[key] = round(1e3 * (time.perf_counter() - t0), 2)
                    assert rr.tobytes() == ref["res"].tobytes(), "host path (%s): records differ" % key
                    assert np.array_equal(eng.counters(), ref["cnt"]), "host path (%s): counters differ" % key
                out["in_place_ok"] = True
                eng._L.aqc_host_free(p)
            except Exception as e:      # noqa: BLE001  (the resident verdict stands; the in-place mode is simply not used)
                out["in_place_ok"] = False
                out["in_place_why"] = repr(e)[:200]
            print(json.dumps(out), flush=True)
        if d is not None:
            d.free()
        if eng is not None:
            eng.close()


def pack():
    """AQC_BATCH_PACK_BASES (2-bit base transport of the host-buffer entry) against the oracle: both filter kernels, the
    prefilter statistics entry, many chunks, a batch whose bases are mostly exceptions (falls back to bytes)."""
    import cases
    import compare
    from afterqc_b200 import _abi
    from afterqc_b200.batch import PackedBatch
    from afterqc_b200.engine import Engine
    from oracle import oracle
    oracle.build()
    odd = PackedBatch.from_reads([("N" * 100, "I" * 100)] * 50 + [("ACGT" * 25, "I" * 100)] * 50, [("acgtn" * 20, "I" * 100)] * 100)
    n_cases = 0
    for bname, batch in (("adversarial", cases.adversarial_batch()), ("pe150", cases.synthetic("pe150", 20000)),
                         ("pe150_jitter", cases.synthetic("pe150", 8000, len_jitter=60)), ("mostly_exceptions", odd), ("se100", cases.synthetic("se100", 20000))):
        for cand in ("warp", "lane"):
            fk, sk = kernel_ids(cand)
            p = cases.make_params("default_f0", paired=batch.paired); p.filter_kernel = fk; p.stat_kernel = sk; p.qc_sample = batch.n // 2
            orc, eng = oracle.Oracle(p), Engine(p)
            a = orc.filter_pairs(batch)
            b = eng.filter_pairs(batch, pack_bases=True, pack_quals=(cand == "lane"))
            what = "pack_bases %s/%s" % (bname, cand)
            compare.assert_records_equal(batch, a, b, what)
            slots = (_abi.QC_R1_POST, _abi.QC_R2_POST) if batch.paired else (_abi.QC_R1_POST,)
            compare.compare_backends(orc, eng, slots, what)
            for be, kw in ((orc, {}), (eng, {"pack_bases": True, "pack_quals": True})):
                be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE if batch.paired else -1, stat_lo=10, stat_hi=batch.n - 3, order_base=0, **kw)
            compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE) if batch.paired else (_abi.QC_R1_PRE,), what + " prefilter")
            orc.close(); eng.close()
            n_cases += 1
    print("pack_bases parity ok: %d cases" % n_cases)


def smoke():
    """lane_kernel and lane2_kernel against the oracle on two small batches (adversarial reads incl. foreign bytes -> list
    mode, synthetic PE150); called from __graft_entry__.smoke() in a child process."""
    import cases
    import compare
    from afterqc_b200 import _abi
    from afterqc_b200.engine import Engine
    from oracle import oracle
    oracle.build()
    for cand in ("lane", "lane2", "lane_st2", "lane_st3"):
        kid, sk = kernel_ids(cand)
        try:
            for bname, batch in (("adversarial", cases.adversarial_batch()), ("pe150", cases.synthetic("pe150", 4096))):
                p = cases.make_params("default_f0"); p.qc_sample = 3000; p.filter_kernel = kid; p.stat_kernel = sk
                orc, eng = oracle.Oracle(p), Engine(p)
                a = orc.filter_pairs(batch); b = eng.filter_pairs(batch)
                compare.assert_records_equal(batch, a, b, "%s smoke %s" % (cand, bname))
                compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "%s smoke %s" % (cand, bname))
                if sk:
                    for be in (orc, eng):
                        be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=7, stat_hi=batch.n - 5, order_base=0)
                    compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "%s smoke %s prefilter" % (cand, bname))
                orc.close(); eng.close()
            print("%s kernel smoke ok (bit-exact vs the oracle: adversarial + pe150)" % cand)
        except Exception as e:      # noqa: BLE001
            print("%s kernel smoke FAILED: %s" % (cand, str(e)[:300]))
    try:        # the packed transport of the host-buffer entry (host threads + unpack kernels), several chunks
        os.environ["AQC_CHUNK_PAIRS"] = "1000"
        for bname, batch in (("adversarial", cases.adversarial_batch()), ("pe150", cases.synthetic("pe150", 4096))):
            p = cases.make_params("default_f0"); p.qc_sample = 3000
            orc, eng = oracle.Oracle(p), Engine(p)
            a = orc.filter_pairs(batch); b = eng.filter_pairs(batch, pack_bases=True, pack_quals=True)
            compare.assert_records_equal(batch, a, b, "pack smoke %s" % bname)
            compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "pack smoke %s" % bname)
            orc.close(); eng.close()
        print("packed transport smoke ok (AQC_BATCH_PACK_BASES | AQC_BATCH_PACK_QUALS, bit-exact vs the oracle)")
    except Exception as e:      # noqa: BLE001
        print("packed transport smoke FAILED: %s" % str(e)[:300])


if __name__ == "__main__":
    t0 = time.time()
    mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if mode == "smoke":
        smoke()
    elif mode == "pack":
        pack()
    elif mode == "parity":
        parity(sys.argv[2] if len(sys.argv) > 2 else "lane")
    else:
        full(int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000, sys.argv[3] if len(sys.argv) > 3 else "lane",
             try_in_place="noinplace" not in sys.argv[4:])
    sys.stderr.write("lane_gpu_check %s: %.1f s\n" % (mode, time.time() - t0))
