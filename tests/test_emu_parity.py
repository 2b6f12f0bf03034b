"""Device code on the CPU: the engine's CUDA sources compiled for the host (tests/emu, a SIMT emulator with fibers as
CUDA threads) and checked bit-exact against the oracle through the same C-ABI.  This is test infrastructure for
machines without a GPU -- it exercises warp collectives, the TMA/mbarrier tile ring (modelled transaction counts) and
the per-lane arithmetic of the kernels; the `-m gpu` tests remain the parity tests proper."""
import numpy as np
import pytest

import cases
import compare
from afterqc_b200 import _abi


@pytest.fixture(scope="module")
def backends(oracle_lib):
    import emu

    def make(params):
        return oracle_lib.Oracle(params), emu.EmuEngine(params)
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
    batch = BATCHES["pe150_jitter"]()
    for kmer in (8, 4):
        orc, eng = backends(_abi.Params.defaults(qc_kmer=kmer))
        lo, hi = batch.n // 10, batch.n - batch.n // 7
        for be in (orc, eng):
            be.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=lo, stat_hi=hi, order_base=0)
        compare.compare_backends(orc, eng, (_abi.QC_R1_PRE, _abi.QC_R2_PRE), "emu stat k=%d" % kmer)
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
