"""bench.clocks_of_timed_region at world_size 2 (gloo, CPU): when only ONE rank's sampler caught the timed region, both ranks
must take the same decision and repeat the same number of steps -- a step of the multi-GPU bench holds a collective, so ranks
that repeated different numbers of steps would hang (they did, once, on 8 GPUs)."""
import os
import socket
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class FakeSampler:
    def __init__(self, n_inside):
        self.n_inside = n_inside
        self.stopped = None

    def samples(self, t0, t1):
        return [(1965.0, 1965.0, [])] * self.n_inside

    def stop(self, t0, t1, note=None):
        self.stopped = (t0, t1, note)
        out = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3}
        if note:
            out["window"] = note
        return out


def _worker(rank, port, inside, ms_totals, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    steps_done = [0]

    def step():
        t = torch.ones(1)
        dist.all_reduce(t)                      # the collective inside a bench step
        steps_done[0] += 1
    s = FakeSampler(inside[rank])
    out = bench.clocks_of_timed_region(s, 100.0, 100.02, ms_totals[rank], 5, 2, torch.device("cpu"), step, dist.barrier)
    q.put((rank, steps_done[0], out.get("window") is not None))
    dist.destroy_process_group()


def _run(inside, ms_totals):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, inside, ms_totals, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    return sorted(q.get(timeout=5) for _ in range(2))


def test_ranks_repeat_together_when_one_sampler_missed():
    got = _run(inside=(3, 0), ms_totals=(20.0, 21.0))          # rank 1 has no sample in the region
    assert got[0][1] == got[1][1] > 0 and got[0][2] and got[1][2]
    assert got[0][1] == 96                                        # ceil(400 / (21 / 5)) steps, sized by the slowest rank


def test_no_repeat_when_every_sampler_caught_the_region():
    got = _run(inside=(2, 4), ms_totals=(80.0, 81.0))
    assert got[0][1] == got[1][1] == 0 and not got[0][2] and not got[1][2]
