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
