"""TEST INFRASTRUCTURE: bench.py's `run_ours` on a machine without a GPU.

torch.cuda is stubbed (streams, events, pinned memory -> plain host memory), the engine is the SIMT-emulated library
(tests/emu) and the child-process kernel checks run tests/lane_gpu_check.py in-process.  Nothing here measures anything:
the emulator has no clock, so the child checks are given made-up kernel times.  What it checks is the CONTROL FLOW of
the benchmark -- kernel auto-selection, the fall-backs when a candidate fails, the resident and host-buffer steps, the
identity checks between them and the one JSON line -- which otherwise only runs on the GPU box at round end.

  python tools/bench_on_emulator.py [--fail lane,lane2,lane_st2,...] [bench.py options]
prints the JSON line bench.py would print.  Used by tests/test_bench_flow_emu.py."""
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(argv, fail=()):
    import torch
    import emu
    import afterqc_b200.engine as E
    from afterqc_b200 import synth
    real_engine = E.Engine

    class Shim:
        def __new__(cls, p, device=0):
            E.Engine = real_engine
            try:
                e = emu.EmuEngine(p)
            finally:
                E.Engine = Shim
            e.last_kernel_ms = lambda: 1.0
            return e

    class Stream:
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

    class Event:
        def __init__(self, **k):
            self.t = 0.0

        def record(self, s=None):
            self.t = time.time()

        def elapsed_time(self, o):
            return max(1e-3, (o.t - self.t) * 1e3)

    saved = {"Engine": E.Engine, "Stream": torch.cuda.Stream, "Event": torch.cuda.Event, "empty": torch.empty, "device": torch.device,
             "gen": synth.generate_device}
    for f in ("set_device", "synchronize", "empty_cache", "set_stream"):
        saved[f] = getattr(torch.cuda, f)
    real_device, real_empty, real_gen = torch.device, torch.empty, synth.generate_device
    try:
        E.Engine = Shim
        torch.cuda.Stream, torch.cuda.Event = Stream, Event
        for f in ("set_device", "synchronize", "empty_cache", "set_stream"):
            setattr(torch.cuda, f, lambda *a, **k: None)

        def empty(*a, **k):
            k.pop("pin_memory", None)
            return real_empty(*a, **k)
        torch.empty = empty
        synth.generate_device = lambda name, n, device="cuda", **kw: real_gen(name, n, device="cpu", **kw)
        import bench
        import lane_gpu_check

        def child(local_rank, pairs, timeout_s=300, candidates="lane"):
            names = [c for c in candidates.split(",") if c]
            out = {c: {"ok": False, "why": "made to fail by the harness"} for c in names}
            run = [c for c in names if c not in fail]
            if run:
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    lane_gpu_check.full(pairs, ",".join(run))
                for ln in buf.getvalue().splitlines():
                    if ln.startswith("{"):
                        j = json.loads(ln)
                        j["ok"] = bool(j.get("identical"))
                        out[j["candidate"]] = j
            for c in run:       # no clock under the emulator: made-up times with lane < lane2 < warp, tile statistics faster than stat_read
                out[c].update({"warp_ms": 3.0, "lane_ms": 2.0 if c.startswith("lane2") else (1.0 if c.startswith("lane") else 3.0),
                               "stat_warp_ms": 1.0, "stat_ms": 0.5})
                if c.endswith("_st3"):
                    out[c]["lane_ms"] -= 0.1       # ... and the filter kernel without statistics code a little faster still
            return out
        saved["child"], saved["emit"] = bench.lane_child_check, bench.emit_json
        bench.lane_child_check = child
        out = []
        bench.emit_json = out.append
        old_argv = sys.argv
        sys.argv = ["bench.py"] + list(argv)
        try:
            args = bench.parse_args()
        finally:
            sys.argv = old_argv
        torch.device = lambda *a, **k: real_device("cpu")
        try:
            bench.run_ours(args)
        finally:
            torch.device = real_device
            bench.lane_child_check, bench.emit_json = saved["child"], saved["emit"]
        return out[0] if out else None
    finally:
        E.Engine = saved["Engine"]
        torch.cuda.Stream, torch.cuda.Event, torch.empty = saved["Stream"], saved["Event"], saved["empty"]
        synth.generate_device = saved["gen"]
        for f in ("set_device", "synchronize", "empty_cache", "set_stream"):
            setattr(torch.cuda, f, saved[f])


if __name__ == "__main__":
    av = sys.argv[1:]
    fail = ()
    if "--fail" in av:
        i = av.index("--fail")
        fail = tuple(av[i + 1].split(","))
        del av[i:i + 2]
    if not any(a == "--pairs" for a in av):
        av = ["--pairs", "3000", "--qc-sample", "1500", "--steps", "2", "--warmup", "1", "--cpu-sample", "2000"] + av
    print(json.dumps(run(av, fail)))
