"""Device code on the CPU: the engine's CUDA sources compiled for the host (tests/emu, a SIMT emulator with fibers as
CUDA threads) and checked bit-exact against the oracle through the same C-ABI.  This is test infrastructure for
machines without a GPU -- it exercises warp collectives, the TMA/mbarrier tile ring (modelled transaction counts) and
the per-lane arithmetic of the kernels; the `-m gpu` tests remain the parity tests proper."""
import numpy as np
import pytest

import cases
import compare
from afterqc_b200 import _abi


KERNELS = {"warp": (_abi.KERNEL_WARP, _abi.STAT_WARP), "lane": (_abi.KERNEL_DEFAULT, _abi.STAT_DEFAULT)}


@pytest.fixture(scope="module", params=["warp", "lane"])
def backends(oracle_lib, request):
    """warp = pair_kernel everywhere (one warp per pair, fused stat_read: aqc_params.filter_kernel = stat_kernel = 1);
    lane = the engine's default path: lane_kernel (one lane per pair) + pair_kernel's list mode + stat_kernel (shared-memory
    histograms; also aqc_stat_reads) for batches of short reads"""
    import emu
    name = request.param

    def make(params):
        params.filter_kernel, params.stat_kernel = KERNELS[name]
        return oracle_lib.Oracle(params), emu.EmuEngine(params)
    make.kernel = name
    return make


BATCHES = {
    "adversarial": lambda: cases.adversarial_batch(),
    "pe150": lambda: cases.synthetic("pe150", 3000),
    "pe150_err3": lambda: cases.synthetic("pe150_err3", 2000),
    "pe150_jitter": lambda: cases.synthetic("pe150", 2000, len_jitter=60),
    "pe250": lambda: cases.synthetic("pe250", 1000),
    "long": lambda: cases.long_read_batch().slice(0, 150),
}


@pytest.mark.parametrize("bname", ["adversarial", "pe150_jitter"])
@pytest.mark.parametrize("pname", ["default_f0", "trim", "poly_wide"])
def test_emu_ops_parity(backends, bname, pname):
    if backends.kernel != "warp":
        pytest.skip("the operator entry always runs pair_kernel")
    batch = BATCHES[bname]()
    orc, eng = backends(cases.make_params(pname))
    compare.assert_records_equal(batch, orc.ops_pairs(batch), eng.ops_pairs(batch), "emu ops %s/%s" % (bname, pname))
    orc.close(); eng.close()


@pytest.mark.parametrize("bname", list(BATCHES))
@pytest.mark.parametrize("pname", list(cases.PARAM_SETS))
def test_emu_filter_parity(backends, bname, pname):
    if bname not in ("adversarial", "pe150") and pname not in ("default_f0", "trim", "strict"):
        pytest.skip("reduced matrix on the emulator")
    batch = BATCHES[bname]()
    orc, eng = backends(cases.make_params(pname))
    a = orc.filter_pairs(batch)
    b = eng.filter_pairs(batch)
    compare.assert_records_equal(batch, a, b, "emu filter %s/%s" % (bname, pname))
    compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "emu filter %s/%s" % (bname, pname))
    orc.close(); eng.close()


def test_emu_stat_parity(backends):
    """aqc_stat_reads: pair_kernel<MODE_STAT> (warp) and stat_kernel (lane); reads > 256 bases take pair_kernel on both"""
    for bname, kmer in (("pe150_jitter", 8), ("pe150_jitter", 4), ("adversarial", 8), ("pe250", 8), ("long", 8), ("adversarial", 1), ("pe150", 2)):
        batch = BATCHES[bname]()
        orc, eng = backends(_abi.Params.defaults(qc_kmer=kmer))
        lo, hi = batch.n // 10, batch.n - batch.n // 7
        for be in (orc, eng):
            be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=lo, stat_hi=hi, order_base=0)
        compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "emu stat %s k=%d" % (bname, kmer))
        # a second call on top (the stamp bitmap of the second launch is built from the first one's stamps), one mate only
        for be in (orc, eng):
            be.stat_reads(batch, -1, _abi.QC_R2_PRE, stat_lo=0, stat_hi=lo, order_base=hi - lo)
        compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "emu stat twice %s k=%d" % (bname, kmer))
        orc.close(); eng.close()


def test_emu_stat_short_head(backends, monkeypatch):
    """stat_kernel with a head of 40 records (AQC_STAT_HEAD): most dense k-mers are first met beyond the head and take the
    kernel's own stamp path (bitmap bit clear -> load, atomicMin); several calls in a row rebuild the bitmap from the table"""
    monkeypatch.setenv("AQC_STAT_HEAD", "40")
    for bname, kmer in (("pe150", 8), ("adversarial", 5), ("pe150_jitter", 3)):
        batch = BATCHES[bname]()
        p = cases.make_params("default_f0"); p.qc_kmer = kmer; p.qc_sample = batch.n // 2
        orc, eng = backends(p)
        third = batch.n // 3
        for be in (orc, eng):
            be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=7, stat_hi=third, order_base=0)
            be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=third, stat_hi=batch.n, order_base=third - 7)
        compare.assert_records_equal(batch, orc.filter_pairs(batch), eng.filter_pairs(batch), "emu short head %s" % bname)
        compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE, _abi.QC_R1_POST, _abi.QC_R2_POST), "emu short head %s k=%d" % (bname, kmer))
        orc.close(); eng.close()


def test_emu_stat_counter_spill(backends, monkeypatch):
    """stat_kernel's 16-bit shared-memory k-mer counters pass 0x4000 (the lane that fills a half moves it to the global table),
    for unflagged k-mers and for k-mers first met beyond the stamped head (bit 15 of the counter set)"""
    if backends.kernel != "lane":
        pytest.skip("stat_kernel only")
    monkeypatch.setenv("AQC_STAT_HEAD", "40")
    batch = cases.homopolymer_batch(1200)
    orc, eng = backends(_abi.Params.defaults(qc_kmer=8))
    for be in (orc, eng):
        be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=0, stat_hi=batch.n, order_base=0)
    compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "emu counter spill")
    k = eng.kmers(_abi.QC_R1_PRE)
    assert int(k[0].max()) > 3 * 0x4000          # the case does what it says
    orc.close(); eng.close()


def test_emu_single_end_and_resident(backends):
    batch = cases.synthetic("se100", 3000)
    for pname in ("default_f0", "trim"):
        orc, eng = backends(cases.make_params(pname, paired=False))
        a = orc.filter_pairs(batch)
        d = eng.upload(batch)                      # the HBM-resident entry (maxlen_kernel + one launch)
        eng.filter_pairs(d)
        b = eng.fetch_results(d)
        compare.assert_records_equal(batch, a, b, "emu se100 %s" % pname)
        compare.compare_backends(orc, eng, (_abi.QC_R1_POST,), "emu se100 %s" % pname)
        d.free(); orc.close(); eng.close()


def test_emu_lane_batch_split_and_sample_gate(backends):
    """first_index gates the postfilter sample (quirk Q10) and counters add up over batches, whichever kernel runs"""
    batch = cases.synthetic("pe150", 2500)
    p = cases.make_params("default_f0"); p.qc_sample = 1500
    orc, eng = backends(p)
    orc.filter_pairs(batch)
    for part in (batch.slice(0, 700), batch.slice(700, 1733), batch.slice(1733, 2500)):
        eng.filter_pairs(part)
    compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "emu split")
    orc.close(); eng.close()


def test_emu_lane_short_and_power_of_two_lengths(backends):
    """reads of 32/64/128/256 bases (bank-conflict lengths for the lane kernel), very short mates, empty tail tile"""
    import random
    from afterqc_b200.batch import PackedBatch
    rng = random.Random(11)
    r1s, r2s = [], []
    for L in (32, 33, 64, 100, 128, 129, 160, 161, 255, 256):
        for _ in range(40):
            frag = rng.randint(20, 2 * L + 40)
            f = cases._rand_seq(rng, frag)
            a = (f + cases._rand_seq(rng, L))[:L]
            b = (cases.revcomp(f) + cases._rand_seq(rng, L))[:rng.choice([L, L, max(5, L - 7)])]
            b = cases._mutate(rng, b, rng.choice([0, 0, 1, 2, 3, 5]), "ACGTN")
            r1s.append((a, cases._rand_qual(rng, len(a)))); r2s.append((b, cases._rand_qual(rng, len(b))))
    for L in (255, 256):                                    # every quality low: the count does not fit in a byte
        r1s.append((cases._rand_seq(rng, L), "#" * L)); r2s.append((cases._rand_seq(rng, L), "#" * L))
    batch = PackedBatch.from_reads(r1s, r2s)
    for pname in ("default_f0", "trim", "loose"):
        orc, eng = backends(cases.make_params(pname))
        a = orc.filter_pairs(batch); b = eng.filter_pairs(batch)
        compare.assert_records_equal(batch, a, b, "emu lengths %s" % pname)
        compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "emu lengths %s" % pname)
        orc.close(); eng.close()


def test_emu_lane_qual2_in_place(backends, monkeypatch):
    """AQC_BATCH_QUAL2_IN_PLACE: mate-2 qualities stay in (page-locked) host memory; identical results, also through the
    general kernel's list mode (foreign bytes) and when the pointer turns out not to be device-accessible (copy fall-back)."""
    if backends.kernel == "warp":
        pytest.skip("only the lane-per-pair path leaves the column in place")
    for bname in ("adversarial", "pe150_err3"):
        batch = BATCHES[bname]()
        for pageable in (False, True):
            if pageable:
                monkeypatch.setenv("AQC_EMU_PAGEABLE", "1")
            else:
                monkeypatch.delenv("AQC_EMU_PAGEABLE", raising=False)
            p = cases.make_params("default_f0"); p.qc_sample = 700
            orc, eng = backends(p)
            a = orc.filter_pairs(batch)
            b = eng.filter_pairs(batch, qual2_in_place=True)
            compare.assert_records_equal(batch, a, b, "emu in-place %s" % bname)
            compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "emu in-place %s" % bname)
            orc.close(); eng.close()


def _empty_mate_batch():
    import random
    from afterqc_b200.batch import PackedBatch
    rng = random.Random(5)
    r1s, r2s = [], []
    for i in range(40):
        a = cases._rand_seq(rng, 100)
        b = "" if i % 3 == 0 else cases._rand_seq(rng, 100)
        r1s.append((a, "I" * len(a))); r2s.append((b, "I" * len(b)))
    return PackedBatch.from_reads(r1s, r2s)


def test_emu_empty_mate_reaches_statread(backends):
    """an empty mate 2 in a good pair (R2 is never length-checked, quirk Q3) still counts in statRead: gcHistogram[0] += 1"""
    batch = _empty_mate_batch()
    orc, eng = backends(cases.make_params("default_f0"))
    a = orc.filter_pairs(batch); b = eng.filter_pairs(batch)
    compare.assert_records_equal(batch, a, b, "emu empty mate")
    compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "emu empty mate")
    assert int(orc.qc(_abi.QC_R2_POST)["gcHistogram"][0]) >= 13
    orc.close(); eng.close()


@pytest.mark.parametrize("kernel", ["warp", "lane"])
@pytest.mark.parametrize("name", ["pe150_default", "pe150_err3_mask_overlap", "pe250_k5_strict", "se100_f0", "testdata"])
def test_emu_pipeline_matches_reference_golden(name, kernel, tmp_path):
    """the whole drop-in pipeline (readers, packed columns, engine calls, writers, JSON) on the emulated engine reproduces
    the reference's golden outputs -- the same check tests/test_gpu_golden.py makes on the GPU"""
    import emu
    import golden_util
    if name not in golden_util.CASES:
        pytest.skip("golden case %s not present" % name)

    def factory(p):
        p.filter_kernel, p.stat_kernel = KERNELS[kernel]
        return emu.EmuEngine(p)
    problems = golden_util.run_case(name, tmp_path, factory)
    assert not problems, problems


def test_emu_resident_columns_need_only_16_bytes_of_slack(backends):
    """the emulator's device buffers end at guard pages: with resident columns that carry exactly the 16 bytes of slack the
    ABI asks for, neither the bulk copies nor the kernels' word loads may touch memory 16 bytes or more past a buffer"""
    import ctypes as C

    def upload_exact(eng, host):
        from afterqc_b200.engine import DeviceBatch
        d = DeviceBatch.__new__(DeviceBatch)
        d.engine = eng; d.n = host.n; d.first_index = host.first_index; d.paired = host.paired
        d.max_len = host.max_len(); d._ptrs = []

        def up(arr):
            p = C.c_void_p()
            eng._check(eng._L.aqc_device_alloc(eng._h, arr.nbytes, C.byref(p)))
            eng._check(eng._L.aqc_memcpy_h2d(eng._h, p, arr.ctypes.data, arr.nbytes))
            d._ptrs.append(p)
            return p
        d.seq1, d.qual1, d.off1 = up(host.seq1), up(host.qual1), up(host.off1)
        if host.paired:
            d.seq2, d.qual2, d.off2 = up(host.seq2), up(host.qual2), up(host.off2)
        else:
            d.seq2 = d.qual2 = d.off2 = None
        d.results = C.c_void_p()
        eng._check(eng._L.aqc_device_alloc(eng._h, max(1, d.n) * 32, C.byref(d.results)))
        d._ptrs.append(d.results)
        return d

    for batch in (cases.synthetic("pe150", 1000), cases.synthetic("pe150", 777, len_jitter=60), cases.adversarial_batch(), cases.synthetic("se100", 999)):
        orc, eng = backends(cases.make_params("default_f0", paired=batch.paired))
        a = orc.filter_pairs(batch)
        d = upload_exact(eng, batch)
        eng.filter_pairs(d)
        compare.assert_records_equal(batch, a, eng.fetch_results(d), "emu exact slack")
        d.free(); orc.close(); eng.close()


def test_emu_host_path_many_chunks(backends, monkeypatch):
    """the chunked H2D -> kernels -> D2H pipeline of the host-buffer entry (double-buffered staging, per-chunk first_index,
    16-byte aligned chunk origins) with chunks of a few hundred pairs, also with mate-2 qualities left in place"""
    monkeypatch.setenv("AQC_CHUNK_PAIRS", "372")
    batch = cases.synthetic("pe150", 2600, len_jitter=37)
    p = cases.make_params("trim"); p.qc_sample = 1500
    for in_place in (False, True):
        orc, eng = backends(p)
        a = orc.filter_pairs(batch)
        b = eng.filter_pairs(batch, qual2_in_place=in_place)
        compare.assert_records_equal(batch, a, b, "emu chunks")
        compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "emu chunks")
        lo, hi = 100, 2100
        for be in (orc, eng):
            be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=lo, stat_hi=hi, order_base=0)
        compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "emu chunks stat")
        orc.close(); eng.close()


def test_emu_host_path_with_length_hint(backends, monkeypatch):
    """aqc_batch.flags bits 0-15 (the longest read of the batch) on a host batch: the engine takes the hint instead of scanning
    the offsets of every chunk; same results"""
    import ctypes as C
    monkeypatch.setenv("AQC_CHUNK_PAIRS", "500")
    batch = cases.synthetic("pe150", 1800, len_jitter=40)
    orc, eng = backends(cases.make_params("default_f0"))
    a = orc.filter_pairs(batch)
    res = np.zeros(batch.n, dtype=_abi.RESULT_DTYPE)
    b = batch.as_struct()
    b.flags |= batch.max_len()
    eng._check(eng._L.aqc_filter_pairs(eng._h, C.byref(b), _abi.MEM_HOST, res.ctypes.data))
    compare.assert_records_equal(batch, a, res, "emu length hint")
    compare.assert_counters_equal(orc.counters(), eng.counters(), "emu length hint")
    orc.close(); eng.close()
