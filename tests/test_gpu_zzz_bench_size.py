"""At bench size (2 M PE150 pairs of bench.py's workload) the engine's default path (lane_kernel + list mode + stat_kernel)
and the warp-per-pair path (pair_kernel with the fused stat_read) must agree on every record, counter, per-cycle array and
both k-mer tables -- resident in HBM and through the host-buffer entry with mate-2 qualities left in page-locked memory."""
import numpy as np
import pytest

import compare
from afterqc_b200 import _abi

pytestmark = pytest.mark.gpu

PAIRS = 2_000_000


@pytest.fixture(scope="module")
def workload():
    import torch
    from afterqc_b200 import synth
    from afterqc_b200.batch import PackedBatch
    t = synth.generate_device("pe150", PAIRS, device="cuda")
    host = PackedBatch(t["seq1"].cpu().numpy(), t["qual1"].cpu().numpy(), t["off1"].cpu().numpy().astype(np.uint32),
                       t["seq2"].cpu().numpy(), t["qual2"].cpu().numpy(), t["off2"].cpu().numpy().astype(np.uint32))
    del t
    torch.cuda.empty_cache()
    return host


def _run(host, warp, qc_sample, in_place=False):
    from afterqc_b200.engine import Engine
    p = _abi.Params.defaults(qc_sample=qc_sample)
    if warp:
        p.filter_kernel, p.stat_kernel = _abi.KERNEL_WARP, _abi.STAT_WARP
    eng = Engine(p)
    lo, hi = 999, 999 + (qc_sample if qc_sample > 0 else host.n)
    if in_place:
        pinned = host.pinned(eng)
        eng.stat_reads(pinned, _abi.QC_R1_PRE, _abi.QC_R2_PRE, lo, hi, 0)
        res = eng.filter_pairs(pinned, qual2_in_place=True)
    else:
        d = eng.upload(host)
        eng.stat_reads(d, _abi.QC_R1_PRE, _abi.QC_R2_PRE, lo, hi, 0)
        eng.filter_pairs(d)
        res = eng.fetch_results(d)
        d.free()
    return eng, res


@pytest.mark.parametrize("qc_sample", [40000, 0])
def test_default_path_equals_warp_path(workload, qc_sample):
    ew, rw = _run(workload, True, qc_sample)
    ed, rd = _run(workload, False, qc_sample)
    assert rw.tobytes() == rd.tobytes(), "records differ"
    compare.compare_backends(ew, ed, (_abi.QC_R1_PRE, _abi.QC_R2_PRE, _abi.QC_R1_POST, _abi.QC_R2_POST), "bench size qc_sample=%d" % qc_sample)
    ew.close(); ed.close()


def test_host_entry_with_qual2_in_place(workload):
    from afterqc_b200.batch import PackedBatch
    if not hasattr(PackedBatch, "pinned"):
        pytest.skip("no pinned-copy helper")
    ed, rd = _run(workload, False, 40000)
    eh, rh = _run(workload, False, 40000, in_place=True)
    assert rd.tobytes() == rh.tobytes(), "records differ"
    compare.compare_backends(ed, eh, (_abi.QC_R1_PRE, _abi.QC_R2_PRE, _abi.QC_R1_POST, _abi.QC_R2_POST), "host entry, qual2 in place")
    ed.close(); eh.close()
