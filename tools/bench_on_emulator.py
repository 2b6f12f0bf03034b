"""TEST INFRASTRUCTURE: bench.py's `run_ours` on a machine without a GPU.

torch.cuda is stubbed (streams, events, pinned memory -> plain host memory) and the engine is the SIMT-emulated library
(tests/emu).  Nothing here measures anything: the emulator has no clock, so the engine's event times are made up.  What it
checks is the CONTROL FLOW of the benchmark -- every --config, the resident and host-buffer steps (autoTrim between the two
engine calls of the full pipeline), the identity check between them, the roofline bookkeeping and the one JSON line --
which otherwise only runs on the GPU box at round end.

  python tools/bench_on_emulator.py [bench.py options]
prints the JSON line bench.py would print.  Used by tests/test_bench_flow_emu.py."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(argv):
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
            e.last_kernel_ms = lambda: 3.0
            e.last_phase_ms = lambda ph: 3.0 if ph < 0 else 1.0
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
        saved["emit"] = bench.emit_json
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
            bench.emit_json = saved["emit"]
        return out[0] if out else None
    finally:
        E.Engine = saved["Engine"]
        torch.cuda.Stream, torch.cuda.Event, torch.empty = saved["Stream"], saved["Event"], saved["empty"]
        synth.generate_device = saved["gen"]
        for f in ("set_device", "synchronize", "empty_cache", "set_stream"):
            setattr(torch.cuda, f, saved[f])


if __name__ == "__main__":
    av = sys.argv[1:]
    if not any(a == "--pairs" for a in av):
        av = ["--pairs", "3000", "--qc-sample", "1500", "--steps", "2", "--warmup", "1", "--cpu-sample", "2000"] + av
    print(json.dumps(run(av)))
