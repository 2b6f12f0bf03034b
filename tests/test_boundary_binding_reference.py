"""The boundary of INTEGRATION.md section 2, executed: the UNMODIFIED reference (oracle/ref_loader.py) runs its own
seqFilter.run() loop with its operator functions -- util.overlap (util.py:88-89), preprocesser.hasPolyX / lowQualityNum /
nNumber (preprocesser.py:30-76) -- rebound to this library's entry points (Engine.overlap / hasPolyX / lowQualityNum / nNumber
over aqc_ops_pairs), and must write the same JSON and the same good/bad files as the reference with its own operators.
Build container only (needs /root/reference); the device code runs under the SIMT emulator there, on a GPU box the same
binding runs on the CUDA library (tests/test_gpu_parity.py::test_operator_interface covers the operator face on hardware)."""
import gzip
import json
import os
import shutil
import sys

import pytest

pytestmark = pytest.mark.reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_reference(args):
    from oracle import ref_loader
    ref_loader.run_cli(args)


def _outputs(work):
    out = {}
    for sub in ("good", "bad", "QC"):
        d = os.path.join(work, sub)
        for fn in sorted(os.listdir(d)):
            if fn.endswith(".html"):
                continue
            p = os.path.join(d, fn)
            data = gzip.open(p, "rb").read() if fn.endswith(".gz") else open(p, "rb").read()
            if fn.endswith(".json"):
                j = json.loads(data)
                for k in ("read1_file", "read2_file", "good_output_folder"):
                    j["command"][k] = None
                data = json.dumps(j, sort_keys=True).encode()
            out[sub + "/" + fn] = data
    return out


@pytest.mark.parametrize("case,extra", [("testdata", []), ("pe150_default", ["-f", "0", "-t", "0", "-p", "20", "-a", "1"])])
def test_reference_loop_with_operators_bound_to_the_engine(tmp_path, case, extra):
    import emu
    import golden_util
    from afterqc_b200 import _abi
    from oracle import ref_loader
    mods = ref_loader.load()
    util, pre = mods["util"], mods["preprocesser"]
    runs = {}
    for mode in ("reference", "bound"):
        work = str(tmp_path / mode)
        os.makedirs(work)
        for fn in ("x_R1.fq.gz", "x_R2.fq.gz"):
            shutil.copy(os.path.join(golden_util.GOLD, case, fn), os.path.join(work, fn))
        args = ["-1", os.path.join(work, "x_R1.fq.gz"), "-2", os.path.join(work, "x_R2.fq.gz"), "-g", os.path.join(work, "good")] + extra
        if mode == "reference":
            _run_reference(args)
        else:
            # the binding a maintainer would add, with the reference's own signatures
            p = _abi.Params.defaults()
            for i, a in enumerate(extra):
                if a == "-p":
                    p.poly_size_limit = int(extra[i + 1])
                if a == "-a":
                    p.allow_mismatch_in_poly = int(extra[i + 1])
            eng = emu.EmuEngine(p)
            calls = {"overlap": 0, "hasPolyX": 0, "lowQualityNum": 0, "nNumber": 0}
            saved = (util.overlap, pre.hasPolyX, pre.lowQualityNum, pre.nNumber)

            def overlap(r1, r2):
                calls["overlap"] += 1
                return eng.overlap(r1, r2)

            def hasPolyX(seq, maxPoly, mismatch):
                calls["hasPolyX"] += 1
                assert (maxPoly, mismatch) == (p.poly_size_limit, p.allow_mismatch_in_poly)
                return eng.hasPolyX(seq)

            def lowQualityNum(read, qual):
                calls["lowQualityNum"] += 1
                assert qual == p.qualified_quality_phred
                return eng.lowQualityNum(read)

            def nNumber(read):
                calls["nNumber"] += 1
                return eng.nNumber(read)
            util.overlap, pre.hasPolyX, pre.lowQualityNum, pre.nNumber = overlap, hasPolyX, lowQualityNum, nNumber
            try:
                _run_reference(args)
            finally:
                util.overlap, pre.hasPolyX, pre.lowQualityNum, pre.nNumber = saved
                eng.close()
            assert all(v > 100 for v in calls.values()), calls
        runs[mode] = _outputs(work)
    assert runs["reference"].keys() == runs["bound"].keys()
    for k in runs["reference"]:
        assert runs["reference"][k] == runs["bound"][k], k
