"""N > 1 path on CPU: world_size-2 gloo run of the sharded pipeline (per-rank backend = the CPU oracle through the
test hook) must reproduce the single-process outputs byte for byte and the same JSON."""
import json
import os
import socket

import pytest

import refcmp
from afterqc_b200 import cli, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("extra", [[], ["-f", "2", "-t", "3", "--qc_sample", "700", "-z"]])
def test_two_rank_gloo_equals_single(tmp_path, oracle_lib, extra):
    import torch.multiprocessing as mp
    from afterqc_b200 import multigpu
    from oracle.oracle import Oracle
    d = str(tmp_path)
    batch = synth.generate("pe150", 2600)
    refcmp.prepare_case(d, batch, subs=("one", "two"))
    refcmp.run_ours(d, "one", True, extra, lambda p: Oracle(p))
    opts, _ = cli.parseCommand(refcmp.cli_args(d, "two", True, extra))
    cli.normalize_options(opts); opts.barcode = False
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=multigpu._worker, args=(r, 2, opts, port, "gloo", None, "oracle.oracle:Oracle")) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs)
    a, b = refcmp.load_json(d, "one"), refcmp.load_json(d, "two")
    diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
    assert not diffs, diffs[:5]
    import gzip
    gz = "-z" in extra
    for f in refcmp.output_files(True, extra):
        fa, fb = os.path.join(d, "one", f), os.path.join(d, "two", f)
        if gz:
            fa += ".gz"; fb += ".gz"
            assert gzip.open(fa).read() == gzip.open(fb).read(), f
        else:
            assert open(fa, "rb").read() == open(fb, "rb").read(), f


def _side_worker(rank, world, port, out_path):
    """two emulated engines (the real device code, tests/emu) over the two halves of a batch full of foreign bytes"""
    import os as _os
    import sys as _sys
    here = _os.path.dirname(_os.path.abspath(__file__))
    for p in (here, _os.path.dirname(here)):
        if p not in _sys.path:
            _sys.path.insert(0, p)
    import numpy as np
    import torch.distributed as dist
    import cases
    import emu
    from afterqc_b200 import _abi, multigpu
    _os.environ["MASTER_ADDR"] = "127.0.0.1"; _os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = cases.adversarial_batch()
    lo, hi = multigpu.shard_range(batch.n, rank, world)
    be = multigpu.DistBackend(emu.EmuEngine(_abi.Params.defaults()))
    be.stat_reads(batch.slice(lo, hi), _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=0, stat_hi=1 << 62, order_base=0)
    merged = [be.kmers(s) for s in (_abi.QC_R1_PRE, _abi.QC_R2_PRE)]
    if rank == 0:
        one = emu.EmuEngine(_abi.Params.defaults())
        one.stat_reads(batch, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=0, stat_hi=1 << 62, order_base=0)
        ok = True
        for s, m in zip((_abi.QC_R1_PRE, _abi.QC_R2_PRE), merged):
            for x, y in zip(one.kmers(s), m):
                ok = ok and np.array_equal(x, y)
        n_side = len(merged[0][2])
        with open(out_path, "w") as f:
            f.write("ok %d" % n_side if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_side_table_first_seen_is_exact(tmp_path):
    """quirk Q12 with shards: the first-seen stamps of k-mers that hold bytes outside util.COMP (lower case, IUPAC, '-') are
    merged from the engines' unresolved side tables and resolved afterwards -- equal to one engine over the whole batch"""
    import torch.multiprocessing as mp
    out = str(tmp_path / "verdict.txt")
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_side_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert all(p.exitcode == 0 for p in procs)
    verdict = open(out).read()
    assert verdict.startswith("ok") and int(verdict.split()[1]) > 1000, verdict
