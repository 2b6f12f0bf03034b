"""GPU parity: hand-written sm_100a kernels (through the C-ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

import cases
import compare
from afterqc_b200 import _abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["default", "warp"])
def backends(oracle_lib, request):
    """default = the engine as shipped (lane_kernel + pair_kernel's list mode + stat_kernel for batches of short reads);
    warp = aqc_params.filter_kernel = stat_kernel = 1: pair_kernel with the fused stat_read everywhere"""
    from afterqc_b200.engine import Engine

    def make(params):
        if request.param == "warp":
            params.filter_kernel, params.stat_kernel = _abi.KERNEL_WARP, _abi.STAT_WARP
        return oracle_lib.Oracle(params), Engine(params)
    make.kernel = request.param
    return make


BATCHES = {
    "adversarial": lambda: cases.adversarial_batch(),
    "pe150": lambda: cases.synthetic("pe150", 20000),
    "pe150_err3": lambda: cases.synthetic("pe150_err3", 12000),
    "pe150_jitter": lambda: cases.synthetic("pe150", 8000, len_jitter=60),
    "pe250": lambda: cases.synthetic("pe250", 6000),
    "long": lambda: cases.long_read_batch(),
}


@pytest.mark.parametrize("bname", list(BATCHES))
@pytest.mark.parametrize("pname", ["default_f0", "trim", "strict", "poly_wide"])
def test_ops_parity(backends, bname, pname):
    if backends.kernel == "warp":
        pytest.skip("the operator entry always runs pair_kernel")
    batch = BATCHES[bname]()
    orc, eng = backends(cases.make_params(pname))
    a = orc.ops_pairs(batch)
    b = eng.ops_pairs(batch)
    compare.assert_records_equal(batch, a, b, "ops %s/%s" % (bname, pname))
    orc.close(); eng.close()


@pytest.mark.parametrize("bname", list(BATCHES))
@pytest.mark.parametrize("pname", list(cases.PARAM_SETS))
def test_filter_parity(backends, bname, pname):
    batch = BATCHES[bname]()
    orc, eng = backends(cases.make_params(pname))
    a = orc.filter_pairs(batch)
    b = eng.filter_pairs(batch)
    compare.assert_records_equal(batch, a, b, "filter %s/%s" % (bname, pname))
    compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "filter %s/%s" % (bname, pname))
    orc.close(); eng.close()


@pytest.mark.parametrize("bname", ["adversarial", "pe150", "pe150_jitter", "long"])
def test_stat_parity(backends, bname):
    batch = BATCHES[bname]()
    if bname == "adversarial":   # statRead needs >= 5 bases (the reference raises below that)
        keep = [i for i in range(batch.n) if batch.off1[i + 1] - batch.off1[i] >= 5 and batch.off2[i + 1] - batch.off2[i] >= 5]
        from afterqc_b200.batch import PackedBatch
        batch = PackedBatch.from_reads([batch.read(1, i) for i in keep], [batch.read(2, i) for i in keep])
    for kmer in (8, 4):
        orc, eng = backends(_abi.Params.defaults(qc_kmer=kmer))
        lo, hi = batch.n // 10, batch.n - batch.n // 7
        for be in (orc, eng):
            be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=lo, stat_hi=hi, order_base=0)
            be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=0, stat_hi=lo, order_base=1 << 40)
        compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "stat %s k=%d" % (bname, kmer))
        orc.close(); eng.close()


def test_stat_counter_spill(backends, monkeypatch):
    """stat_kernel's 16-bit shared-memory k-mer counters pass 0x4000 in every CTA (reads that pile onto a handful of k-mers),
    for k-mers met in the stamped head and for flagged ones first met beyond it (short head: AQC_STAT_HEAD)"""
    batch = cases.homopolymer_batch(60000)
    for head in ("40", None):
        if head: monkeypatch.setenv("AQC_STAT_HEAD", head)
        else: monkeypatch.delenv("AQC_STAT_HEAD", raising=False)
        orc, eng = backends(_abi.Params.defaults(qc_kmer=8))
        for be in (orc, eng):
            be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=0, stat_hi=batch.n, order_base=0)
        compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "counter spill head=%s" % head)
        assert int(eng.kmers(_abi.QC_R1_PRE)[0].max()) > 148 * 0x4000
        orc.close(); eng.close()


def test_single_end_parity(backends):
    batch = cases.synthetic("se100", 30000)
    for pname in ("default_f0", "trim", "loose"):
        orc, eng = backends(cases.make_params(pname, paired=False))
        a = orc.filter_pairs(batch); b = eng.filter_pairs(batch)
        compare.assert_records_equal(batch, a, b, "se100 %s" % pname)
        compare.compare_backends(orc, eng, (_abi.QC_R1_POST,), "se100 %s" % pname)
        orc.close(); eng.close()


def test_device_resident_equals_host(backends):
    """The HBM-resident entry (what bench.py's kernel number uses) gives the same records and counters."""
    batch = cases.synthetic("pe150", 50000)
    orc, eng = backends(cases.make_params("default_f0"))
    a = orc.filter_pairs(batch)
    d = eng.upload(batch)
    eng.filter_pairs(d)
    b = eng.fetch_results(d)
    compare.assert_records_equal(batch, a, b, "device-resident")
    compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "device-resident")
    d.free(); orc.close(); eng.close()


def test_batch_split_invariance(backends):
    """Counters are additive over batches and first_index gates the postfilter sample (quirk Q10)."""
    batch = cases.synthetic("pe150", 9000)
    p = cases.make_params("default_f0"); p.qc_sample = 5000
    orc, eng = backends(p)
    orc.filter_pairs(batch)
    parts = [batch.slice(0, 2500), batch.slice(2500, 6100), batch.slice(6100, 9000)]
    for part in parts:
        eng.filter_pairs(part)
    compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "split")
    orc.close(); eng.close()


def test_empty_mate_reaches_statread(backends):
    """an empty mate 2 in a good pair (R2 is never length-checked, quirk Q3) still counts in statRead: gcHistogram[0] += 1"""
    import random
    from afterqc_b200.batch import PackedBatch
    rng = random.Random(5)
    r1s, r2s = [], []
    for i in range(40):
        a = cases._rand_seq(rng, 100)
        b = "" if i % 3 == 0 else cases._rand_seq(rng, 100)
        r1s.append((a, "I" * len(a))); r2s.append((b, "I" * len(b)))
    batch = PackedBatch.from_reads(r1s, r2s)
    orc, eng = backends(cases.make_params("default_f0"))
    a = orc.filter_pairs(batch); b = eng.filter_pairs(batch)
    compare.assert_records_equal(batch, a, b, "empty mate")
    compare.compare_backends(orc, eng, (_abi.QC_R1_POST, _abi.QC_R2_POST), "empty mate")
    for be in (orc, eng):
        be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE)
    compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "empty mate, prefilter")
    orc.close(); eng.close()


def test_operator_interface(backends):
    """util.overlap / hasPolyX / lowQualityNum / nNumber through the GPU; KATs from the reference."""
    from afterqc_b200.engine import Engine
    eng = Engine(_abi.Params.defaults())
    assert eng.overlap(
        "CAGCGCCTACGGGCCCCTTTTTCTGCGCGACCGCGTGGCTGTGGGCGCGGATGCCTTTGAGCGCGGTGACTTCTCACTGCGTATCGAGCCGCTGGAGGTCTCCC",
        "ACCTCCAGCGGCTCGATACGCAGTGAGAAGTCACCGCGCTCAAAGGCATCCGCGCCCACAGCCACGCGGTCGCGCAGAAAAAGGGGCCCGTAGGCGCGGCTCCC") == (-5, 99, 1)
    assert eng.overlap(
        "CAGCGCCTACGGGCCCCTTTTTCTGCGCGACCGCGTGGCTGTGGGCGCGGATGCCTTTGAGCGCGGTGACTTCTCACTGCGTATCGAGC",
        "ACCTCCAGCGGCTCGATACGCAGTGAGAAGTCACCGCGCTCAAAGGCATCCGCGCCCACAGCCACGCGGTCGCGCAGAAAAAGGGGTCC") == (10, 79, 1)
    assert eng.hasPolyX("ACGT" * 5 + "G" * 33 + "AT" + "ACGT" * 5) == "G"
    assert eng.hasPolyX("ACGT" * 30) is None
    assert eng.nNumber(["@x", "ANNACNGT" * 5, "+", "I" * 40]) == 15
    assert eng.lowQualityNum(["@x", "A" * 40, "+", "#" * 7 + "I" * 33]) == 7
    eng.close()


def test_full_scale_properties():
    """BASELINE-size invariants that need no oracle: class counts partition the batch, base sums match the
    records, and running the same resident batch twice doubles every counter."""
    import torch
    from afterqc_b200 import synth
    from afterqc_b200.engine import Engine
    n = 2_000_000
    t = synth.generate_device("pe150", n, device="cuda")
    from afterqc_b200.batch import PackedBatch
    host = PackedBatch(t["seq1"].cpu().numpy(), t["qual1"].cpu().numpy(), t["off1"].cpu().numpy().astype(np.uint32),
                       t["seq2"].cpu().numpy(), t["qual2"].cpu().numpy(), t["off2"].cpu().numpy().astype(np.uint32))
    del t
    torch.cuda.empty_cache()
    eng = Engine(_abi.Params.defaults())
    d = eng.upload(host)
    eng.filter_pairs(d)
    res = eng.fetch_results(d)
    c1 = eng.counters()
    ci = _abi.CIDX
    assert c1[ci["TOTAL_READS"]] == n
    cls_counts = np.bincount(res["cls"], minlength=9)
    assert cls_counts.sum() == n and cls_counts[0] == c1[ci["GOOD_READS"]]
    for k, name in enumerate(["BADTRIM1", "BADTRIM2", "BADLEN", "BADPOL", "BADLQC", "BADNCT", "BADDIFF", "BADMISMATCH"], start=1):
        assert cls_counts[k] == c1[ci[name]]
    good = res["cls"] == 0
    assert int(res["len1"][good].sum()) == c1[ci["GOOD_BASES_R1"]]
    assert int(res["len2"][good].sum()) == c1[ci["GOOD_BASES_R2"]]
    assert c1[_abi.C_DISTANCE_HIST:_abi.C_DISTANCE_HIST + 1001].sum() + cls_counts[1:7].sum() - \
        ((res["cls"] == 3) & (res["ov_len"] > 0)).sum() * 0 >= c1[ci["GOOD_READS"]]
    assert c1[_abi.C_OVERLAP_HIST:_abi.C_OVERLAP_HIST + 1001].sum() == n - cls_counts[1:7].sum() + ((res["cls"] == 3) & (res["ov_len"] > 0)).sum()
    eng.filter_pairs(d)
    c2 = eng.counters()
    assert np.array_equal(c2, 2 * c1)
    res2 = eng.fetch_results(d)
    assert res2.tobytes() == res.tobytes()
    d.free(); eng.close()


def test_domain_errors_are_loud(backends, oracle_lib):
    """Situations where the reference raises: the engine reports the same sticky error codes as the oracle."""
    from afterqc_b200.batch import PackedBatch
    from afterqc_b200.engine import Engine, EngineError
    # statRead on a 3-base read (IndexError in the reference, qualitycontrol.py:106-108)
    short = PackedBatch.from_reads([("ACG", "III"), ("ACGTACGTAC", "I" * 10)], [("ACGTA", "IIIII"), ("ACGTACGTAC", "I" * 10)])
    for make in (lambda: oracle_lib.Oracle(_abi.Params.defaults()), lambda: Engine(_abi.Params.defaults())):
        be = make()
        with pytest.raises(Exception) as ei:
            be.stat_reads(short, _abi.QC_R1_PRE, _abi.QC_R2_PRE)
            be.counters()
        assert getattr(ei.value, "code", None) == _abi.ERR_TOO_SHORT_STAT
        be.close()
    # a read longer than MAX_LEN
    long_ = PackedBatch.from_reads([("A" * 1001, "I" * 1001)], [("ACGTA", "IIIII")])
    eng = Engine(_abi.Params.defaults())
    with pytest.raises(EngineError) as ei:
        eng.filter_pairs(long_)
    assert ei.value.code == _abi.ERR_TOO_LONG
    eng.close()
    # side-table overflow: 16-entry table, hundreds of distinct N-containing k-mers
    import random
    rng = random.Random(3)
    reads = [("".join(rng.choice("ACGTN") for _ in range(100)), "I" * 100) for _ in range(64)]
    eng = Engine(_abi.Params.defaults(kmer_side_log2=4))
    with pytest.raises(EngineError) as ei:
        eng.stat_reads(PackedBatch.from_reads(reads, reads), _abi.QC_R1_PRE, _abi.QC_R2_PRE)
        eng.counters()
    assert ei.value.code == _abi.ERR_KMER_TABLE_FULL
    eng.close()
