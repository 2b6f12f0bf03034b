#!/usr/bin/env python
"""bench.py -- the B200 AfterQC hot path on the BASELINE.json configs (metric: M read-pairs/s; achieved HBM GB/s vs peak).

  --config pe150        (default) BASELINE configs[2]: synthetic PE150, 10 M pairs per GPU, reference defaults + `-f 0 -t 0`
                        (qc_sample 200000: prefilter statistics over records 999..200998, postfilter statistics over the good
                        pairs with index < 199999, filter + overlap scan + adapter trim + correction over every pair)
  --config se100        configs[1]: synthetic SE100, 10 M reads, quality + polyX + N filters (no mate, no overlap),
                        `-f 0 -t 0 --qc_sample 0`: every read is stat'd before and, if good, after the filter
  --config pe250_full   configs[3]: synthetic PE250, 2.5 M pairs per GPU (20 M over 8), the FULL pipeline: `--qc_sample 0`,
                        prefilter statistics of every read, autoTrim resolved on the host from them (qualitycontrol.py:359-408),
                        filter with those trims, postfilter statistics of every good pair
  --config pe150_err3   configs[4]: synthetic PE150 with 3 % injected error, 25 M pairs per GPU (200 M over 8), defaults + `-f 0 -t 0`

One "step" = one full pass of that path over the batch:  aqc_stat_reads(prefilter window) [+ autoTrim] + aqc_filter_pairs(all).

  value         whole-job M read-pairs/s (reads/s for se100) with the batch resident in HBM (CUDA events on the launching stream)
  e2e           the same pass through the host-buffer C-ABI entry (pinned host columns -> H2D -> kernels -> D2H of the 32-byte
                records), copies inside the timed region
  roofline      algorithmic bytes of the dominant launch / its event-timed duration vs the measured HBM peak
                (MEASURED_PEAKS.json); `phases` lists every launch group of the step the same way
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm, kind "port") on a bounded sample of the same
                workload with all host threads (rank 0, N = 1 only); cpu_baseline_python = the reference's own Python under
                CPython 3.12, timed in the build container (profiles/r02_reference_python_timing.json; it cannot run on the box)

N > 1 (torchrun): weak scaling, one process per GPU, each rank filters its own contiguous shard (global record indices
rank*n ..), no data-path collective; every step ends with the NCCL all-reduce of the counter blocks (the only exchange the
path has).  Before timing, a 200 k-pair slice is run sharded + all-reduced and compared with one engine over the whole
slice (`shard_parity`).  `--impl reference` times the CPU arm alone (rank 0 only) on the same config, steps and warm-up.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

STAT_LO = 999           # statFile sets the first 999 records aside (qualitycontrol.py:333-341)
E2E_MAX_PAIRS = 10_000_000      # pinned host columns of the e2e leg are capped at this many records per GPU

CONFIGS = {
    "pe150": dict(synth="pe150", paired=True, L=150, units=10_000_000, qc_sample=200_000, autotrim=False, unit="M read-pairs/s",
                  metric="read-pairs/s PE150 (filter + overlap correction + adapter trim + per-cycle QC)",
                  workload="synthetic PE150 10M pairs/GPU, overlap correction + adapter trim, defaults -f 0 -t 0 (BASELINE configs[2])"),
    "se100": dict(synth="se100", paired=False, L=100, units=10_000_000, qc_sample=0, autotrim=False, unit="M reads/s",
                  metric="reads/s SE100 (quality + polyX + N filters + per-cycle QC of every read)",
                  workload="synthetic SE100 10M reads/GPU, quality+polyX filter only, -f 0 -t 0 --qc_sample 0 (BASELINE configs[1])"),
    "pe250_full": dict(synth="pe250", paired=True, L=250, units=2_500_000, qc_sample=0, autotrim=True, unit="M read-pairs/s",
                       metric="read-pairs/s PE250 (full pipeline: prefilter QC of every read, autoTrim, filter + overlap correction, postfilter QC)",
                       workload="synthetic PE250 2.5M pairs/GPU (20M over 8), full pipeline --qc_sample 0 + autoTrim (BASELINE configs[3])"),
    "pe150_err3": dict(synth="pe150_err3", paired=True, L=150, units=25_000_000, qc_sample=200_000, autotrim=False, unit="M read-pairs/s",
                       metric="read-pairs/s PE150 3% error (filter + overlap correction stress + adapter trim + per-cycle QC)",
                       workload="synthetic PE150 3% injected error 25M pairs/GPU (200M over 8), defaults -f 0 -t 0 (BASELINE configs[4])"),
}


def algo_bytes_per_unit(cfg):
    """every base / quality byte once, one result record, the offsets: 4L + 32 + 8 per pair, 2L + 16 + 4 per SE read"""
    return 4 * cfg["L"] + 40 if cfg["paired"] else 2 * cfg["L"] + 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="pe150", choices=list(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="records per GPU (default = the config's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="records in the CPU sample (0 = auto)")
    ap.add_argument("--qc-sample", type=int, default=-1, help="--qc_sample of the workload (profiling runs on fewer pairs scale it to keep the mix)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--filter-kernel", default="default", choices=["default", "warp"],
                    help="default = the engine as shipped (lane_kernel + list mode + stat_kernel); warp = pair_kernel with the fused stat_read "
                         "everywhere (aqc_params.filter_kernel = stat_kernel = 1), the round-1 path, for comparison")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.cfg = dict(CONFIGS[a.config])
    a.n = a.pairs or a.cfg["units"]
    a.qs = a.cfg["qc_sample"] if a.qc_sample < 0 else a.qc_sample
    return a


def workload_config(args, world):
    cfg = args.cfg
    # the same dictionary in the GPU arm and in --impl reference (the driver compares them)
    return {"workload": cfg["workload"], "config": args.config, "units_per_gpu": args.n, "read_len": cfg["L"], "qc_sample": args.qs,
            "autotrim": cfg["autotrim"], "filter_kernel": args.filter_kernel,
            "l2": "inputs (GB per GPU) exceed the 126 MB L2; no explicit flush",
            "parallelism": ("read-sharded x%d, NCCL all-reduce of the counter blocks per step" % world) if world > 1 else "1 GPU"}


# ------------------------------------------------------------------------------------------------
# host-side autoTrim of the full pipeline (qualitycontrol.py:359-408 on the prefilter counters; preprocesser.py:260-280)
# ------------------------------------------------------------------------------------------------
def resolve_autotrim(backend, paired, qs, kmer=8):
    from afterqc_b200 import _abi
    from afterqc_b200.qc import QualityControl
    q = QualityControl(qs, kmer).load(backend.qc(_abi.QC_R1_PRE), None)
    q.calcReadLen(); q.calcPercents(); q.calcQualities()
    return q.autoTrim()           # trim_pair_same (default true) copies R1's values to R2


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (test infrastructure) timed on host cores -- the only place bench.py runs oracle/
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample_size(args, threads):
    if args.cpu_sample:
        return args.cpu_sample
    per_thread = {"pe150": 100_000, "se100": 60_000, "pe250_full": 12_000, "pe150_err3": 100_000}[args.config]
    return min(args.n, per_thread * threads, 2_000_000)


def cpu_arm(args, batch, threads, steps=1, warmup=0):
    """Oracle throughput on the PackedBatch `batch` (a prefix of the workload), split over `threads` host threads (each thread
    runs the whole path on its slice).  The sample keeps the workload's mix: a finite QC window is scaled with the sample."""
    from afterqc_b200 import _abi
    from oracle import oracle as orc_mod
    orc_mod.build()
    cfg = args.cfg
    sample = batch.n
    qs = args.qs if args.qs <= 0 else max(1000, sample * args.qs // cfg["units"])
    params = _abi.Params.defaults(qc_sample=qs, paired=1 if cfg["paired"] else 0)
    per = (sample + threads - 1) // threads
    parts = [batch.slice(i * per, min(sample, (i + 1) * per)) for i in range(threads) if i * per < sample]     # global record indices kept

    def work(part, out, k):
        o = orc_mod.Oracle(_abi.Params.defaults(qc_sample=qs, paired=1 if cfg["paired"] else 0))
        lo = min(STAT_LO, max(0, sample - 1))
        hi = (1 << 62) if qs <= 0 else lo + qs
        o.stat_reads(part, _abi.QC_R1_PRE, _abi.QC_R2_PRE if cfg["paired"] else -1, stat_lo=lo, stat_hi=hi)
        if cfg["autotrim"]:
            f, t = resolve_autotrim(o, cfg["paired"], qs)
            p2 = _abi.Params.defaults(qc_sample=qs, paired=1 if cfg["paired"] else 0, trim_front=f, trim_tail=t, trim_front2=f, trim_tail2=t)
            o.set_params(p2)
        res = o.filter_pairs(part)
        out[k] = int((res["cls"] == 0).sum())
        o.close()

    del params
    times = []
    for it in range(warmup + steps):
        out = [0] * len(parts)
        ths = [threading.Thread(target=work, args=(p, out, k)) for k, p in enumerate(parts)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return {"units_per_s": sample * len(times) / total, "seconds": total, "threads": len(parts),
            "sample": sample, "qc_sample_scaled": qs, "ms_per_step": 1e3 * total / len(times)}


def host_sample(args, units, threads):
    """CPU-generated prefix of the workload (reference arm: no GPU work at all)."""
    import torch
    from afterqc_b200 import synth
    try:
        torch.set_num_threads(max(1, min(threads, 32)))     # torchrun exports OMP_NUM_THREADS=1
    except Exception:
        pass
    return synth.generate(args.cfg["synth"], units)


def cpu_sample_text(args, r):
    return ("%d records of the same %s workload per step (%s), %.1f s of wall time on %d threads; C oracle oracle/aqc_oracle.c, one "
            "context per host thread" % (r["sample"], args.config,
                                         "QC window scaled to %d reads, the same mix" % r["qc_sample_scaled"] if r["qc_sample_scaled"] > 0 else "--qc_sample 0: every read stat'd",
                                         r["seconds"], r["threads"]))


def python_reference_timing(config):
    p = os.path.join(ROOT, "profiles", "r02_reference_python_timing.json")
    try:
        with open(p) as f:
            return json.load(f).get(config)
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample = cpu_sample_size(args, threads)
    r = cpu_arm(args, host_sample(args, sample, threads), threads, steps=args.steps, warmup=args.warmup)
    ups = r["units_per_s"] / 1e6
    cfg = args.cfg
    line = {
        "impl": "reference",
        "metric": cfg["metric"], "value": ups, "unit": cfg["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, 1 if args.gpus <= 1 else args.gpus),
        "cpu_baseline": {"value": ups, "unit": cfg["unit"], "cores": r["threads"], "kind": "port",
                         "sample": cpu_sample_text(args, r) + "; the reference itself is Python 2 and cannot run on the GPU box"},
        "cpu_baseline_python": python_reference_timing(args.config),
        "e2e": {"value": ups, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 50 ms in the background (started before the warm-up: the tool needs a few hundred ms to come up);
    the samples whose timestamps fall into the timed region are the ones reported"""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = "/tmp/aqc_clocks_%d_%d.csv" % (os.getpid(), gpu_index)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def samples(self, t0, t1):
        """[(sm, sm_max, reasons)] of the samples stamped within [t0, t1] (epoch seconds)"""
        out = []
        try:
            with open(self.path, errors="replace") as f:
                lines = f.readlines()
        except Exception:
            return out
        for line in lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                stamp, frac = p[0].split(".") if "." in p[0] else (p[0], "0")
                ts = time.mktime(time.strptime(stamp, "%Y/%m/%d %H:%M:%S")) + float("0." + frac)
                sm, smax = float(p[1]), float(p[2])
            except Exception:
                continue
            if ts < t0 - 0.03 or ts > t1 + 0.03:
                continue
            reasons = [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9])
                       if v.lower().startswith("active")]
            out.append((sm, smax, reasons))
        return out

    def stop(self, t0, t1, note=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        try:
            time.sleep(0.08)
            got = self.samples(t0, t1)
        except Exception:
            got = []
        try:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.f.close()
            os.unlink(self.path)
        except Exception:
            pass
        if not got:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        out = {"sm_mhz": float(np.median([g[0] for g in got])), "sm_max_mhz": float(max(g[1] for g in got)),
               "reasons": sorted({r for g in got for r in g[2]}), "samples": len(got)}
        if note:
            out["window"] = note
        return out


def clocks_of_timed_region(sampler, t0, t1, ms_total, steps, world, device, step, barrier):
    """The clock samples stamped inside the timed region [t0, t1].  When that region was shorter than the tool's period on ANY
    rank, every rank repeats the same step untimed for ~0.4 s -- the same number of steps everywhere, because a step holds a
    collective at N > 1 -- and the samples of that stretch are reported instead, with a note."""
    import torch
    have = torch.tensor([float(len(sampler.samples(t0, t1))), -float(ms_total)], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(have, op=dist.ReduceOp.MIN)
    if have[0].item() >= 1:
        return sampler.stop(t0, t1)
    ms_step_max = max(1e-3, -float(have[1].item()) / max(1, steps))
    reps = int(min(2000, max(1, math.ceil(400.0 / ms_step_max))))
    for _ in range(reps):
        step()
    barrier()
    return sampler.stop(t1, time.time(),
                        "no sample fell into the %.0f ms timed region (on some rank); taken over %d untimed repeats of the same step right after it"
                        % ((t1 - t0) * 1e3, reps))


class TorchBatch:
    """HBM-resident batch backed by torch tensors (plumbing only: memory + RNG)."""

    def __init__(self, t, first_index, n, max_len, paired):
        import torch
        self.t = t
        self.n = n
        self.first_index = first_index
        self.max_len = max_len
        self.paired = paired
        self.results = torch.empty(max(1, n) * 32, dtype=torch.uint8, device=t["seq1"].device)

    def struct(self, lo=0, hi=None):
        """aqc_batch over records [lo, hi); lo must be a multiple of 4 (16-byte aligned offsets for the bulk copies)."""
        from afterqc_b200 import _abi
        hi = self.n if hi is None else hi
        assert lo % 4 == 0
        b = _abi.Batch()
        b.first_index = self.first_index + lo
        b.n = hi - lo
        b.flags = self.max_len
        b.seq1 = self.t["seq1"].data_ptr(); b.qual1 = self.t["qual1"].data_ptr(); b.off1 = self.t["off1"].data_ptr() + 4 * lo
        if self.paired:
            b.seq2 = self.t["seq2"].data_ptr(); b.qual2 = self.t["qual2"].data_ptr(); b.off2 = self.t["off2"].data_ptr() + 4 * lo
        else:
            b.seq2 = None; b.qual2 = None; b.off2 = None
        return b


def make_device_workload(cfg, device, n, seed, first_index):
    import torch
    from afterqc_b200 import synth
    t = synth.generate_device(cfg["synth"], n, device=device, seed=seed)
    for k in ("off1", "off2"):
        if k in t:
            last = t[k][-1:].to(torch.int32)
            t[k] = torch.cat([t[k].to(torch.int32), last.expand(8)])   # slack entries for the 16-byte granular copies
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return TorchBatch(t, first_index, n, cfg["L"], cfg["paired"])


def cuda_tensor_view(ptr, n, device):
    """torch tensor aliasing `n` 64-bit elements at device address `ptr` (for in-place NCCL reductions)."""
    import torch

    class _Shim:
        pass
    s = _Shim()
    s.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(s, device=device)


class Runner:
    """one engine + its workload: the resident step, the host-buffer step, the reductions of the N > 1 runs"""

    def __init__(self, args, wb, device, local_rank, stream, world, dist):
        from afterqc_b200 import _abi
        from afterqc_b200.engine import Engine
        self.abi, self.args, self.wb, self.device, self.world, self.dist, self.stream = _abi, args, wb, device, world, dist, stream
        cfg = args.cfg
        self.cfg = cfg
        self.paired = cfg["paired"]
        kw = dict(qc_sample=args.qs, paired=1 if self.paired else 0)
        if args.filter_kernel == "warp":
            kw.update(filter_kernel=_abi.KERNEL_WARP, stat_kernel=_abi.STAT_WARP)
        self.base_kw = kw
        self.eng = Engine(_abi.Params.defaults(**kw), device=local_rank)
        self.eng.set_stream(stream.cuda_stream)
        self.L = self.eng._L
        self.qc2 = _abi.QC_R2_PRE if self.paired else -1
        # records of this shard inside the prefilter window [999, 999 + qc_sample) -- all from 999 on with --qc_sample 0
        n, fi = wb.n, wb.first_index
        self.w_lo_g = STAT_LO
        self.w_hi_g = (STAT_LO + args.qs) if args.qs > 0 else (1 << 62)
        self.s_lo = max(self.w_lo_g, fi) - fi
        self.s_hi = min(self.w_hi_g, fi + n) - fi
        self.has_window = self.s_hi > self.s_lo
        self.s_lo_al = (max(self.s_lo, 0) // 4) * 4
        self.trims = None
        self.phase_ms = None
        self.sum_views, self.min_views = [], []
        if world > 1:
            import torch  # noqa: F401
            p, cnt = self.eng.device_ptr(0)
            self.sum_views.append(cuda_tensor_view(p, cnt, device))
            for slot in range(4):
                for what in (1, 2, 3, 4, 5, 6):
                    p, cnt = self.eng.device_ptr(what, slot)
                    self.sum_views.append(cuda_tensor_view(p, cnt, device))
                p, cnt = self.eng.device_ptr(7, slot)
                self.min_views.append(cuda_tensor_view(p, cnt, device))

    def allreduce_counters(self):
        """the path's only exchange: reduce (copies of) the counter blocks over NVLink -- one SUM, one MIN"""
        import torch
        for op, views in ((self.dist.ReduceOp.SUM, self.sum_views), (self.dist.ReduceOp.MIN, self.min_views)):
            c = torch.cat(views)
            self.dist.all_reduce(c, op=op)
        return c

    def _autotrim(self):
        """full pipeline: trims from the prefilter statistics (device -> host fetch of one QC slot, host float code, new params)"""
        f, t = resolve_autotrim(self.eng, self.paired, self.args.qs)
        self.trims = (f, t)
        self.eng.set_params(self.abi.Params.defaults(trim_front=f, trim_tail=t, trim_front2=f, trim_tail2=t, **self.base_kw))

    def step_resident(self, record_phases=False):
        eng, L, wb, abi = self.eng, self.L, self.wb, self.abi
        ph = [0.0, 0.0, 0.0, 0.0]          # prefilter statistics | filter kernel | list mode | postfilter statistics
        if self.has_window:
            b = wb.struct(self.s_lo_al, self.s_hi)
            eng._check(L.aqc_stat_reads(eng._h, C.byref(b), abi.MEM_DEVICE, abi.QC_R1_PRE, self.qc2, self.w_lo_g, self.w_hi_g, 0))
            if record_phases:
                ph[0] = eng.last_phase_ms(-1)
        if self.cfg["autotrim"]:
            self._autotrim()
        b = wb.struct()
        eng._check(L.aqc_filter_pairs(eng._h, C.byref(b), abi.MEM_DEVICE, wb.results.data_ptr()))
        if record_phases:
            ph[1], ph[2], ph[3] = eng.last_phase_ms(0), eng.last_phase_ms(1), eng.last_phase_ms(2)
            self.phase_ms = ph
        if self.world > 1:
            self.allreduce_counters()

    def close(self):
        try:
            self.eng.close()
        except Exception:       # noqa: BLE001
            pass


def shard_parity_check(args, device, local_rank, stream, world, rank, dist):
    """N > 1: a 200 k-record slice (the same on every rank) is run sharded + all-reduced, and rank 0 compares counters, per-cycle
    arrays and dense k-mer tables with ONE engine over the whole slice (the exact merge of the side tables is covered by
    tests/test_multigpu_gloo.py)."""
    import torch
    from afterqc_b200 import _abi
    n = 200_000
    cfg = args.cfg
    full = make_device_workload(cfg, device, n, seed=777, first_index=0)
    lo, hi = (n * rank // world) // 4 * 4, (n if rank == world - 1 else (n * (rank + 1) // world) // 4 * 4)
    a2 = argparse.Namespace(**vars(args)); a2.qs = 50_000 if args.qs > 0 else 0
    part = TorchBatch(full.t, 0, n, cfg["L"], cfg["paired"])
    r = Runner(a2, part, device, local_rank, stream, world, dist)
    eng = r.eng

    def run(e, lo_, hi_):
        s_lo, s_hi = max(STAT_LO, lo_), min(hi_, (STAT_LO + a2.qs) if a2.qs > 0 else hi_)
        if s_hi > s_lo:
            b = part.struct(s_lo // 4 * 4, s_hi)
            e._check(e._L.aqc_stat_reads(e._h, C.byref(b), _abi.MEM_DEVICE, _abi.QC_R1_PRE, r.qc2, STAT_LO, (STAT_LO + a2.qs) if a2.qs > 0 else (1 << 62), 0))
        b = part.struct(lo_, hi_)
        res = torch.empty(max(1, hi_ - lo_) * 32, dtype=torch.uint8, device=device)
        e._check(e._L.aqc_filter_pairs(e._h, C.byref(b), _abi.MEM_DEVICE, res.data_ptr()))
        e.sync()
    run(eng, lo, hi)
    # reduce in place so that rank 0's fetches return the merged blocks
    for v in r.sum_views:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    bias = -(1 << 63)       # stamps are unsigned and "never" is all ones: flip the sign bit so that the signed MIN orders them
    for v in r.min_views:
        v.bitwise_xor_(bias)
        dist.all_reduce(v, op=dist.ReduceOp.MIN)
        v.bitwise_xor_(bias)
    torch.cuda.synchronize()
    ok = True
    why = "identical"
    if rank == 0:
        from afterqc_b200.engine import Engine
        one = Engine(_abi.Params.defaults(**r.base_kw), device=local_rank)
        one.set_stream(stream.cuda_stream)
        run(one, 0, n)
        if not np.array_equal(one.counters(), eng.counters()):
            ok, why = False, "counters differ"
        slots = (0, 1, 2, 3) if cfg["paired"] else (0, 2)
        for s in slots:
            if not ok:
                break
            a, b = one.qc(s), eng.qc(s)
            for f in a.dtype.names:
                if not np.array_equal(a[f], b[f]):
                    ok, why = False, "QC slot %d field %s differs" % (s, f)
            ka, kb = one.kmers(s), eng.kmers(s)
            if ok and not (np.array_equal(ka[0], kb[0]) and np.array_equal(ka[1], kb[1])):
                ok, why = False, "dense k-mer table of slot %d differs" % s
        one.close()
    r.close()
    del full, part
    torch.cuda.empty_cache()
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item()), why


def run_ours(args):
    import torch
    import torch.distributed as dist
    from afterqc_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL logs its version there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = None
    if world > 1:       # page-locked columns and the threads that fill them: local to the GPU's PCIe root
        from afterqc_b200.multigpu import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local_rank)
    cfg = args.cfg
    n, QS = args.n, args.qs
    stream = torch.cuda.Stream(device)          # the launching stream of every kernel below (torch events time it)
    torch.cuda.set_stream(stream)
    first_index = rank * n

    shard_parity = None
    if world > 1:
        try:
            shard_parity = shard_parity_check(args, device, local_rank, stream, world, rank, dist)
        except Exception as e:      # noqa: BLE001
            shard_parity = (False, "check raised %r" % (e,))

    wb = make_device_workload(cfg, device, n, seed=20260927 + rank, first_index=first_index)
    R = Runner(args, wb, device, local_rank, stream, world, dist)
    eng, L = R.eng, R.L

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident (kernel-side) number ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    warm = max(3, args.warmup)
    for _ in range(warm):
        R.step_resident()
    barrier()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        R.step_resident()
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = eng.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = clocks_of_timed_region(sampler, t_wall0, t_wall1, ms_total, args.steps, world, device, R.step_resident, barrier)
    # durations of the step's launch groups (CUDA events recorded by the engine on the launching stream), three more steps
    phases = []
    for _ in range(3):
        R.step_resident(record_phases=True)
        phases.append(R.phase_ms)
    phase_ms = [float(np.mean([p[i] for p in phases])) for i in range(4)]
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---------------- end-to-end through the host-buffer C-ABI ----------------
    e2e = None
    if not args.no_e2e:
        ne = min(n, E2E_MAX_PAIRS)
        cols = ("seq1", "qual1", "off1") + (("seq2", "qual2", "off2") if cfg["paired"] else ())
        ends = {"seq1": "off1", "qual1": "off1", "seq2": "off2", "qual2": "off2"}
        host = {}
        for k in cols:
            src = wb.t[k][:ne + 9] if k.startswith("off") else wb.t[k][:(int(wb.t[ends[k]][ne].item()) & 0xFFFFFFFF) + 16]
            h = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            h.copy_(src)
            host[k] = h
        res_host = torch.empty(ne * 32, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        xflags = [0]
        e_lo, e_hi = R.s_lo, min(R.s_hi, ne)
        e_window = e_hi > e_lo

        def hstruct(lo, hi):
            b = _abi.Batch()
            b.first_index = first_index + lo
            b.n = hi - lo
            b.flags = xflags[0] | cfg["L"]                  # bits 0-15: the longest read (saves the engine a pass over the offsets)
            b.seq1 = host["seq1"].data_ptr(); b.qual1 = host["qual1"].data_ptr(); b.off1 = host["off1"].data_ptr() + 4 * lo
            if cfg["paired"]:
                b.seq2 = host["seq2"].data_ptr(); b.qual2 = host["qual2"].data_ptr(); b.off2 = host["off2"].data_ptr() + 4 * lo
            else:
                b.seq2 = None; b.qual2 = None; b.off2 = None
            return b

        off1 = host["off1"].numpy().view(np.uint32).astype(np.int64)      # uint32 offsets held in int32 tensors (columns up to 4 GiB)
        off2 = host["off2"].numpy().view(np.uint32).astype(np.int64) if cfg["paired"] else None
        nm = 2 if cfg["paired"] else 1

        def col_bytes(lo, hi):
            return 2 * int(off1[hi] - off1[lo]) + (2 * int(off2[hi] - off2[lo]) if cfg["paired"] else 0)
        h2d_all = col_bytes(0, ne) + 4 * nm * (ne + 1)
        if e_window:
            h2d_all += col_bytes(e_lo, e_hi) + 4 * nm * (e_hi - e_lo + 1)
        d2h = 32 * ne

        def step_e2e():
            if e_window:
                b = hstruct(e_lo, e_hi)
                eng._check(L.aqc_stat_reads(eng._h, C.byref(b), _abi.MEM_HOST, _abi.QC_R1_PRE, R.qc2, R.w_lo_g, R.w_hi_g, 0))
            if cfg["autotrim"]:
                R._autotrim()
            b = hstruct(0, ne)
            eng._check(L.aqc_filter_pairs(eng._h, C.byref(b), _abi.MEM_HOST, res_host.data_ptr()))
            if world > 1:
                R.allreduce_counters()

        e_steps = max(1, min(args.steps, 5))

        def time_e2e():
            for _ in range(2):
                step_e2e()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(e_steps):
                step_e2e()
            a1.record(stream)
            barrier()
            tt = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            try:        # same records, same parameters: the host-buffer path must return the records of the resident path
                torch.cuda.synchronize()
                same = bool(torch.equal(res_host.to(device, non_blocking=False), wb.results[:ne * 32]))
            except Exception:
                same = None
            return float(tt.item()) / e_steps, same

        variants = {}
        modes = [("copy_all_columns", 0)] + ([("qual2_in_place", _abi.BATCH_QUAL2_IN_PLACE)] if cfg["paired"] and args.filter_kernel == "default" else [])
        best = None
        for name, fl in modes:
            xflags[0] = fl
            res_host.zero_()
            try:
                ms_v, same_v = time_e2e()
            except Exception as ex:      # noqa: BLE001
                variants[name] = {"error": repr(ex)[:200]}
                continue
            hb = h2d_all
            if fl & _abi.BATCH_QUAL2_IN_PLACE:
                # counted from the tensors that are copied: the qual2 column stays in pinned memory; the kernels pull one 32-byte
                # sector per byte pair of the correction walk and the mate-2 qualities of the sampled pairs over PCIe
                n_edits = int(wb.results[:ne * 32].view(-1, 32)[:, 1].sum().item())
                sampled = ne if QS <= 0 else min(ne, max(0, QS - 1 - first_index))
                hb += -int(off2[ne] - off2[0]) + 32 * n_edits + sampled * cfg["L"]
            variants[name] = {"value": world * ne / (ms_v * 1e-3) / 1e6, "ms_per_step": ms_v, "results_match_resident": same_v, "h2d_bytes_per_step": hb}
            if same_v and (best is None or variants[name]["value"] > variants[best]["value"]):
                best = name
        xflags[0] = 0
        if best is not None:
            v = variants[best]
            e2e = {"value": v["value"], "unit": cfg["unit"], "h2d_bytes_per_step": v["h2d_bytes_per_step"], "d2h_bytes_per_step": d2h,
                   "ms_per_step": v["ms_per_step"], "steps": e_steps, "results_match_resident": v["results_match_resident"], "mode": best,
                   "records_per_gpu": ne, "per_gpu_h2d_GBps": v["h2d_bytes_per_step"] / (v["ms_per_step"] * 1e-3) / 1e9,
                   "variants": variants,
                   "note": "aqc_stat_reads + aqc_filter_pairs with AQC_MEM_HOST on pinned host columns; chunked H2D / kernels / D2H pipeline inside; "
                           "qual2_in_place = AQC_BATCH_QUAL2_IN_PLACE, the mate-2 quality column is not copied (the filter never reads it "
                           "outside the correction walk and the sampled statistics)"}
        else:
            e2e = {"value": None, "unit": cfg["unit"], "variants": variants}
        del host, res_host

    # ---------------- roofline: every launch group of the step, the dominant one on top ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak = float(json.load(f)["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    def span(t):        # uint32 offsets held in an int32 tensor
        return (int(t[n].item()) & 0xFFFFFFFF) - (int(t[0].item()) & 0xFFFFFFFF)
    b1 = span(wb.t["off1"])
    b2 = span(wb.t["off2"]) if cfg["paired"] else 0
    nm = 2 if cfg["paired"] else 1
    lane = args.filter_kernel == "default"
    n_pre = max(0, R.s_hi - R.s_lo) if R.has_window else 0
    n_post = n if QS <= 0 else min(n, max(0, QS - 1 - first_index))
    per_rec = (b1 + b2) / max(1, n)          # bases of one record (both mates)
    groups = [
        ("prefilter statistics: " + ("aqc::stamp_bits_kernel + aqc::stat_kernel (one warp per read, histograms in shared memory)" if lane
                                     else "aqc::pair_kernel<MODE_STAT> (stat_read)"), phase_ms[0], 2 * per_rec * n_pre, n_pre, "stat_pre"),
        ("filter: " + ("aqc::lane_kernel (one lane per pair, no statistics)" if lane else "aqc::pair_kernel<MODE_FILTER> (one warp per pair, fused stat_read)"),
         phase_ms[1], 2 * (b1 + b2) + 32 * n + 4 * nm * n, n, "filter"),
        ("list mode: aqc::pair_kernel<MODE_LIST> over the pairs lane_kernel handed over (bytes outside A,C,G,T,N)", phase_ms[2], 0, 0, "list"),
        ("postfilter statistics: aqc::stamp_bits_kernel + aqc::stat_kernel<POST> over the sampled good pairs' records", phase_ms[3],
         (2 * per_rec + 32) * n_post, n_post, "stat_post"),
    ]
    traffic_tab = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r02_dram_bytes.json")) as f:
            traffic_tab = json.load(f)
    except Exception:
        traffic_tab = {}
    ph_out = []
    for name, ms, ab, units, key in groups:
        if ms <= 0:
            continue
        tr = traffic_tab.get(args.config, {}).get(key) if lane else None
        ph_out.append({"launches": name, "ms": ms, "algorithmic_bytes": ab, "achieved_GBps": ab / (ms * 1e-3) / 1e9,
                       "frac": ab / (ms * 1e-3) / 1e9 / peak, "units": units,
                       "traffic": (tr["dram_bytes_per_unit"] * units) if tr else None})
    dom = max(ph_out, key=lambda p: p["ms"]) if ph_out else None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "achieved": dom["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic"],
                    "kernel": "%s, %d records/launch" % (dom["launches"], dom["units"]), "kernel_ms": dom["ms"],
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "peak_source": peak_src,
                    "whole_step": {"algorithmic_bytes": algo_bytes_per_unit(cfg) * n, "ms": ms_per_step if world == 1 else None,
                                   "achieved_GBps": algo_bytes_per_unit(cfg) * n / (ms_per_step * 1e-3) / 1e9 if world == 1 else None},
                    "phases": ph_out,
                    "note": "integer-issue bound filter kernel, shared-memory-atomic bound statistics kernel: DESIGN.md section 3"}

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            threads = host_threads()
            sample = cpu_sample_size(args, threads)
            from afterqc_b200.batch import PackedBatch, SLACK

            def hostcol(name, end):
                a = np.zeros(end + SLACK, dtype=np.uint8)
                a[:end] = wb.t[name][:end].cpu().numpy()
                return a
            o1 = wb.t["off1"][:sample + 1].cpu().numpy().astype(np.uint32)
            if cfg["paired"]:
                o2 = wb.t["off2"][:sample + 1].cpu().numpy().astype(np.uint32)
                hb = PackedBatch(hostcol("seq1", int(o1[-1])), hostcol("qual1", int(o1[-1])), o1,
                                 hostcol("seq2", int(o2[-1])), hostcol("qual2", int(o2[-1])), o2)
            else:
                hb = PackedBatch(hostcol("seq1", int(o1[-1])), hostcol("qual1", int(o1[-1])), o1)
            r = cpu_arm(args, hb, threads)
            cpu = {"value": r["units_per_s"] / 1e6, "unit": cfg["unit"], "cores": r["threads"], "kind": "port", "sample": cpu_sample_text(args, r)}
        confd = workload_config(args, world)
        line = {
            "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": confd, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_python": python_reference_timing(args.config),
            "autotrim_resolved": list(R.trims) if R.trims else None, "input_GB_per_gpu": 2 * (b1 + b2) / 1e9,
        }
        if numa is not None:
            line["numa_rank0"] = numa
        if shard_parity is not None:
            line["shard_parity"] = shard_parity[0]
            line["shard_parity_note"] = shard_parity[1]
    R.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit_json(line)


_REAL_STDOUT = None


def claim_stdout():
    """Route everything any library prints to fd 1 (NCCL's version banner, ...) to stderr and keep the real stdout for the
    ONE JSON line rank 0 emits."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit_json(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
