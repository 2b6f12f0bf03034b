#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 AfterQC hot path (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[2] = synthetic PE150, 10 M read pairs per GPU,
reference defaults + `-f 0 -t 0` (so qc_sample = 200000: prefilter statistics over records
999..200998 of both mates, postfilter statistics over the good pairs with index < 200000, and
filter + overlap scan + adapter trim + correction over every pair).  One "step" = one full pass
of that path over the batch:  aqc_stat_reads(window) + aqc_filter_pairs(all pairs).

  value      whole-job M read-pairs/s with the batch resident in HBM (CUDA events on the launching stream)
  e2e        the same pass through the host-buffer C-ABI entry (pinned host columns -> H2D -> kernels -> D2H of
             the 32-byte records), copies inside the timed region
  roofline   algorithmic bytes of the dominant kernel launch (pair_kernel, filter mode) / its event-timed
             duration vs the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm, kind "port") on a bounded sample
             of the same workload with all host threads

N > 1 (torchrun): weak scaling, one process per GPU, each rank filters its own contiguous shard of
10 M pairs (global record indices rank*n ..), no data-path collective; every step ends with the NCCL
all-reduce of the counter blocks (the only exchange the path has).
`--impl reference` times the CPU arm alone (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PAIRS_PER_GPU = 10_000_000
READ_LEN = 150
QC_SAMPLE = 200_000
STAT_LO = 999
ALGO_BYTES_PER_PAIR = 4 * READ_LEN + 32 + 8          # 2 mates x (bases + quals) + result record + 2 offsets
WORKLOAD = "synthetic PE150 10M pairs/GPU, overlap correction + adapter trim, defaults -f 0 -t 0 (BASELINE configs[2])"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU (default = BASELINE config)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU sample (0 = auto)")
    ap.add_argument("--qc-sample", type=int, default=QC_SAMPLE, help="--qc_sample of the workload (profiling runs on fewer pairs scale it to keep the 2%% mix)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--filter-kernel", default="auto", choices=["auto", "warp", "lane", "lane2"],
                    help="filter kernel: warp = pair_kernel (one warp per pair), lane = lane_kernel (one lane per pair); auto = lane "
                         "only if it first proves bit-identical to warp on this GPU (child process with a timeout, then the full batch)")
    ap.add_argument("--stat-kernel", default="auto", choices=["auto", "warp", "lane", "lane_post"],
                    help="statRead kernel (aqc_params.stat_kernel): warp = stat_read (one warp per read), lane = stat_tile / "
                         "stat_lane_kernel (one lane per read), lane_post = the same with the sampled statistics in their own launch; "
                         "auto = one of the lane forms only after the same two-stage identity check, and only if faster")
    ap.add_argument("--no-pack", action="store_true", help="do not try the packed base transport (AQC_BATCH_PACK_BASES) in the e2e measurement")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (test infrastructure) timed on host cores -- the only place bench.py runs oracle/
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


CPU_SAMPLE_CAP = 2_000_000


def cpu_sample_size(pairs, threads, requested):
    return requested or min(pairs, 100_000 * threads, CPU_SAMPLE_CAP)


def cpu_arm(batch, threads, steps=1, warmup=0):
    """Oracle throughput on the PackedBatch `batch` (a prefix of the workload), split over `threads` host threads.
    The sample keeps the workload's mix: the QC window is scaled to the same 2 % of the pairs."""
    from afterqc_b200 import _abi
    from oracle import oracle as orc_mod
    orc_mod.build()
    sample_pairs = batch.n
    qs = max(1000, sample_pairs * QC_SAMPLE // PAIRS_PER_GPU)
    params = _abi.Params.defaults(qc_sample=qs)
    per = (sample_pairs + threads - 1) // threads
    parts = [batch.slice(i * per, min(sample_pairs, (i + 1) * per)) for i in range(threads) if i * per < sample_pairs]

    def work(part, out, k):
        o = orc_mod.Oracle(params)
        lo = min(STAT_LO, max(0, part.n - 1))
        o.stat_reads(part, _abi.QC_R1_PRE, _abi.QC_R2_PRE, stat_lo=lo, stat_hi=lo + qs)
        res = o.filter_pairs(part)
        out[k] = int((res["cls"] == 0).sum())
        o.close()

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
    return {"pairs_per_s": sample_pairs * len(times) / total, "seconds": total, "threads": len(parts),
            "sample_pairs": sample_pairs, "qc_sample_scaled": qs, "ms_per_step": 1e3 * total / len(times)}


def host_sample(pairs, threads):
    """CPU-generated prefix of the workload (reference arm: no GPU work at all)."""
    import torch
    from afterqc_b200 import synth
    try:
        torch.set_num_threads(max(1, min(threads, 32)))     # torchrun exports OMP_NUM_THREADS=1
    except Exception:
        pass
    return synth.generate("pe150", pairs)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample = cpu_sample_size(args.pairs, threads, args.cpu_sample)
    r = cpu_arm(host_sample(sample, threads), threads, steps=args.steps, warmup=min(args.warmup, 1))
    mps = r["pairs_per_s"] / 1e6
    line = {
        "impl": "reference",
        "metric": "read-pairs/s PE150 (filter + overlap correction + adapter trim + per-cycle QC)",
        "value": mps, "unit": "M read-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_pairs_per_step": sample},
        "cpu_baseline": {"value": mps, "unit": "M read-pairs/s", "cores": r["threads"], "kind": "port",
                         "sample": "%d pairs of the PE150 workload per step, QC window scaled to %d reads (same 2%% mix); "
                                   "C oracle (oracle/aqc_oracle.c), one context per host thread; the reference itself is "
                                   "Python 2 and cannot run on the GPU box" % (sample, r["qc_sample_scaled"])},
        "e2e": {"value": mps, "unit": "M read-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        with open(self.path) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); smax.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


class TorchBatch:
    """HBM-resident batch backed by torch tensors (plumbing only: memory + RNG)."""

    def __init__(self, t, first_index, n, max_len):
        import torch
        self.t = t
        self.n = n
        self.first_index = first_index
        self.max_len = max_len
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
        b.seq2 = self.t["seq2"].data_ptr(); b.qual2 = self.t["qual2"].data_ptr(); b.off2 = self.t["off2"].data_ptr() + 4 * lo
        return b


def make_device_workload(device, n, seed, first_index):
    import torch
    from afterqc_b200 import synth
    t = synth.generate_device("pe150", n, device=device, seed=seed)
    for k in ("off1", "off2"):
        last = t[k][-1:].to(torch.int32)
        t[k] = torch.cat([t[k].to(torch.int32), last.expand(8)])   # slack entries for the 16-byte granular copies
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return TorchBatch(t, first_index, n, READ_LEN)


def cuda_tensor_view(ptr, n, torch_dtype, device):
    """torch tensor aliasing `n` 64-bit elements at device address `ptr` (for in-place NCCL reductions)."""
    import torch

    class _Shim:
        pass
    s = _Shim()
    s.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(s, device=device)


def lane_child_check(local_rank, pairs, timeout_s=300, candidates="lane"):
    """tests/lane_gpu_check.py `full` in a child process with a timeout: pair_kernel and every candidate of the comma-separated
    list on `pairs` pairs of the bench workload, every output compared.  The candidates were committed without having run on
    hardware, so a hang, a crash or a mismatch must not take the benchmark down: anything but a clean 'identical' keeps the
    measured kernels.  Returns {candidate: verdict}; a verdict printed before a later crash or time-out of the child stands."""
    names = [c for c in candidates.split(",") if c]
    env = dict(os.environ)
    ids = [x for x in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip() != ""]
    env["CUDA_VISIBLE_DEVICES"] = (ids[local_rank] if local_rank < len(ids) else ids[0]) if ids else str(local_rank)
    cmd = [sys.executable, os.path.join(ROOT, "tests", "lane_gpu_check.py"), "full", str(pairs), ",".join(names)]
    t0 = time.time()
    note, rc = None, 0
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        stdout, stderr, rc = r.stdout or "", r.stderr or "", r.returncode
    except subprocess.TimeoutExpired as e:
        stdout, stderr, rc = e.stdout or "", e.stderr or "", None
        note = "child timed out after %d s" % timeout_s
    except Exception as e:      # noqa: BLE001
        return {c: {"ok": False, "why": "child could not run: %r" % (e,)} for c in names}
    if isinstance(stdout, bytes):
        stdout = stdout.decode("utf-8", "replace")
    if isinstance(stderr, bytes):
        stderr = stderr.decode("utf-8", "replace")
    if rc not in (0, None):
        note = "child exit %d: %s" % (rc, (stderr or stdout)[-300:].replace("\n", " | "))
    out = {c: {"ok": False} for c in names}
    for ln in stdout.splitlines():
        ln = ln.strip()
        if ln.startswith("{"):
            try:
                j = json.loads(ln)
            except ValueError:
                continue
            c = j.get("candidate")
            if c in out:
                out[c] = dict(j, ok=bool(j.get("identical")))
    seconds = round(time.time() - t0, 1)
    for c in names:
        out[c]["seconds_child"] = seconds
        if note:
            out[c]["after_verdict" if out[c]["ok"] else "why"] = note
            if rc is None:
                out[c].pop("in_place_ok", None)
        if not out[c]["ok"]:
            out[c].setdefault("why", "no verdict in the child's output")
    return out


def lane_full_size_check(wb, n, qs, local_rank, stream, cand_kernel, cand_stat=0, window=None):
    """Both kernels once over the resident full-size batch in this process: records, scalar counters, histograms, error
    matrix, per-cycle statistics and k-mer tables must be identical (with a candidate statistics kernel: also the prefilter
    slots after aqc_stat_reads over `window` = (first record, end record, global lo, global hi))."""
    import torch
    from afterqc_b200 import _abi
    from afterqc_b200.engine import Engine
    ref = None
    slots = (_abi.QC_R1_POST, _abi.QC_R2_POST) + ((_abi.QC_R1_PRE, _abi.QC_R2_PRE) if (cand_stat and window) else ())
    for k, sk in ((_abi.KERNEL_WARP, _abi.STAT_DEFAULT), (cand_kernel, cand_stat)):
        e = Engine(_abi.Params.defaults(qc_sample=qs, filter_kernel=k, stat_kernel=sk), device=local_rank)
        e.set_stream(stream.cuda_stream)
        res = torch.empty(max(1, n) * 32, dtype=torch.uint8, device=wb.results.device)
        if cand_stat and window:
            b = wb.struct(window[0], window[1])
            e._check(e._L.aqc_stat_reads(e._h, C.byref(b), _abi.MEM_DEVICE, _abi.QC_R1_PRE, _abi.QC_R2_PRE, window[2], window[3], 0))
        b = wb.struct()
        e._check(e._L.aqc_filter_pairs(e._h, C.byref(b), _abi.MEM_DEVICE, res.data_ptr()))
        e.sync()
        got = (res, e.counters(), [e.qc(s) for s in slots], [e.kmers(s) for s in slots])
        e.close()
        if ref is None:
            ref = got
            continue
        if not bool(torch.equal(ref[0], got[0])):
            return False, "records differ"
        if not np.array_equal(ref[1], got[1]):
            return False, "counters differ"
        for a, c in zip(ref[2], got[2]):
            for f in a.dtype.names:
                if not np.array_equal(a[f], c[f]):
                    return False, "QC field %s differs" % f
        for a, c in zip(ref[3], got[3]):
            for x, y in zip(a, c):
                if not np.array_equal(x, y):
                    return False, "k-mer tables differ"
    return True, "identical"


def run_ours(args):
    import torch
    import torch.distributed as dist
    from afterqc_b200 import _abi
    from afterqc_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL logs its version there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    n = args.pairs
    QS = args.qc_sample
    stream = torch.cuda.Stream(device)          # the launching stream of every kernel below (torch events time it)
    torch.cuda.set_stream(stream)
    first_index = rank * n

    # ---------------- which filter kernel ----------------
    # warp = pair_kernel (measured since the first GPU session); lane = lane_kernel; lane2 = lane2_kernel (never on hardware when
    # committed).  "auto": each candidate must prove identical to pair_kernel in a child process (a hang or crash there costs a
    # timeout, not the benchmark), the fastest identical one is then compared once more on the full-size batch in this process.
    KID = {"warp": _abi.KERNEL_WARP, "lane": _abi.KERNEL_LANE, "lane2": _abi.KERNEL_LANE2}
    selection = {"requested": args.filter_kernel}
    chosen = args.filter_kernel if args.filter_kernel != "auto" else "warp"
    # aqc_params.stat_kernel: 0 = stat_read; 2 / 3 = statRead with one lane per read (aqc_stat2.cuh), inside the lane-per-pair filter
    # kernel / in a launch of its own after it
    stat2 = {"lane": _abi.STAT_LANE, "lane_post": _abi.STAT_LANE_POST}.get(args.stat_kernel, 0)
    child = None
    if args.filter_kernel == "auto":
        best_ms = None
        verdicts = lane_child_check(local_rank, min(n, 2_000_000), timeout_s=300, candidates="lane,lane2")
        for cand in ("lane", "lane2"):
            c = verdicts[cand]
            selection["child_check_" + cand] = c
            if not c.get("ok"):
                continue
            ms, wms = c.get("lane_ms", 0), c.get("warp_ms", 0)
            if ms > 0 and ms < wms and (best_ms is None or ms < best_ms):
                best_ms, chosen, child = ms, cand, c
    if args.stat_kernel == "auto":
        # the statistics kernel: the chosen filter kernel with stat_kernel = 2 must be identical in a child process and
        # faster on filter + prefilter statistics together (written without GPU access, emulator-verified when committed)
        best_total = None
        # pair_kernel keeps its fused statistics: levels 2 and 3 are the same launches there
        levels = (("_st2", _abi.STAT_LANE),) if chosen == "warp" else (("_st3", _abi.STAT_LANE_POST), ("_st2", _abi.STAT_LANE))
        verdicts = lane_child_check(local_rank, min(n, 2_000_000), timeout_s=240, candidates=",".join(chosen + sfx for sfx, _ in levels))
        for suffix, level in levels:
            c = verdicts[chosen + suffix]
            selection["child_check_" + chosen + suffix] = c
            if not c.get("ok"):
                continue
            # the chosen filter kernel with stat_read: timed by its own child check, or the warp kernel of this one
            base_ms = child.get("lane_ms", 0) if child else (c.get("warp_ms", 0) if chosen == "warp" else None)
            new_ms = c.get("lane_ms", 0) + c.get("stat_ms", 0)
            if base_ms is None:         # explicit --filter-kernel: only the prefilter launches are comparable
                faster = 0 < c.get("stat_ms", 0) < c.get("stat_warp_ms", 0)
            else:
                faster = 0 < new_ms < base_ms + c.get("stat_warp_ms", 0)
            if faster and (best_total is None or new_ms < best_total):
                best_total, stat2 = new_ms, level
    wb = make_device_workload(device, n, seed=20260927 + rank, first_index=first_index)
    # records of this shard inside the prefilter window [999, 999 + qc_sample) (global indices)
    w_lo_g, w_hi_g = STAT_LO, STAT_LO + QS
    s_lo = max(w_lo_g, first_index) - first_index
    s_hi = min(w_hi_g, first_index + n) - first_index
    has_window = s_hi > s_lo
    s_lo_al = (s_lo // 4) * 4
    if (chosen != "warp" or stat2) and (args.filter_kernel == "auto" or args.stat_kernel == "auto"):
        attempts = [(chosen, stat2)] + ([(chosen, 0)] if (stat2 and chosen != "warp") else [])
        chosen, stat2 = "warp", 0
        for cand, st in attempts:       # a failing statistics kernel must not cost the filter kernel its place
            try:
                ok, why = lane_full_size_check(wb, n, QS, local_rank, stream, KID[cand], st,
                                               (s_lo_al, s_hi, w_lo_g, w_hi_g) if has_window else None)
            except Exception as e:      # noqa: BLE001
                ok, why = False, "full-size check raised %r" % (e,)
            selection["full_size_check_%s%s" % (cand, "_st%d" % st if st else "")] = why
            if ok:
                chosen, stat2 = cand, st
                break
    if world > 1:       # every rank runs the same kernel: the most conservative choice any rank made
        flag = torch.tensor([KID[chosen]], dtype=torch.int32, device=device)
        flags = [torch.zeros_like(flag) for _ in range(world)]
        dist.all_gather(flags, flag)
        ids = [int(f.item()) for f in flags]
        if len(set(ids)) == 1:
            agreed = ids[0]
        elif _abi.KERNEL_WARP in ids:
            agreed = _abi.KERNEL_WARP
        else:
            agreed = _abi.KERNEL_LANE
        chosen = {v: k for k, v in KID.items()}[agreed]
        sflag = torch.tensor([stat2, -stat2], dtype=torch.int32, device=device)
        dist.all_reduce(sflag, op=dist.ReduceOp.MIN)
        stat2 = stat2 if int(sflag[0].item()) == -int(sflag[1].item()) else 0      # every rank the same level, else stat_read
    use_lane = chosen != "warp"
    selection["used"] = chosen
    selection["stat_kernel_requested"] = args.stat_kernel
    selection["stat_kernel_used"] = {0: "warp (stat_read)", 2: "lane (stat_tile in the filter kernel / stat_lane_kernel)",
                                     3: "lane_post (stat_lane_kernel, also for the sampled pairs of the filter launch)"}[stat2]
    kernel_label = {"warp": "aqc::pair_kernel (MODE_FILTER, one warp per pair)",
                    "lane": "aqc::lane_kernel (one lane per pair) + aqc::pair_kernel list mode",
                    "lane2": "aqc::lane2_kernel (one lane per pair, 2-column stage, dynamic tiles) + aqc::pair_kernel list mode"}[chosen]
    if use_lane and stat2 == _abi.STAT_LANE_POST:
        kernel_label += " + aqc::stat_lane_kernel<POST> (sampled statistics)"

    if world > 1:       # the packed base transport's host threads: the ranks of one box share its cores
        os.environ.setdefault("AQC_PACK_THREADS", str(max(2, (os.cpu_count() or 8) // (2 * world))))
    params = _abi.Params.defaults(qc_sample=QS, filter_kernel=KID[chosen], stat_kernel=stat2)
    eng = Engine(params, device=local_rank)
    eng.set_stream(stream.cuda_stream)
    L = eng._L

    reduce_views = []
    if world > 1:
        p, cnt = eng.device_ptr(0)
        reduce_views.append(("sum", cuda_tensor_view(p, cnt, torch.int64, device)))
        for slot in range(4):
            for what in (1, 2, 3, 4, 5, 6):
                p, cnt = eng.device_ptr(what, slot)
                reduce_views.append(("sum", cuda_tensor_view(p, cnt, torch.int64, device)))
            p, cnt = eng.device_ptr(7, slot)
            reduce_views.append(("min", cuda_tensor_view(p, cnt, torch.int64, device)))

    sum_views = [v for op, v in reduce_views if op == "sum"]
    min_views = [v for op, v in reduce_views if op == "min"]

    def step_resident():
        if has_window:
            b = wb.struct(s_lo_al, s_hi)
            eng._check(L.aqc_stat_reads(eng._h, C.byref(b), _abi.MEM_DEVICE, _abi.QC_R1_PRE, _abi.QC_R2_PRE, w_lo_g, w_hi_g, 0))
        b = wb.struct()
        eng._check(L.aqc_filter_pairs(eng._h, C.byref(b), _abi.MEM_DEVICE, wb.results.data_ptr()))
        if world > 1:   # the path's only exchange: reduce (copies of) the counter blocks over NVLink -- one SUM, one MIN
            for op, views in (("sum", sum_views), ("min", min_views)):
                c = torch.cat(views)
                dist.all_reduce(c, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MIN)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident (kernel-side) number ----------------
    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    filter_ms = []
    for _ in range(args.steps):
        step_resident()
        filter_ms.append(None)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    # duration of the dominant kernel (filter-mode pair_kernel): events on its launching stream, one more launch
    kms = []
    for _ in range(3):
        b = wb.struct()
        eng._check(L.aqc_filter_pairs(eng._h, C.byref(b), _abi.MEM_DEVICE, wb.results.data_ptr()))
        kms.append(eng.last_kernel_ms())
    kernel_ms = float(np.mean(kms))
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---------------- end-to-end through the host-buffer C-ABI ----------------
    e2e = None
    if not args.no_e2e:
        host = {}
        for k in ("seq1", "qual1", "seq2", "qual2", "off1", "off2"):
            h = torch.empty(wb.t[k].shape, dtype=wb.t[k].dtype, pin_memory=True)
            h.copy_(wb.t[k])
            host[k] = h
        res_host = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()

        xflags = [0]            # transport flags of the calls: AQC_BATCH_QUAL2_IN_PLACE (lane kernels only) | AQC_BATCH_PACK_BASES

        def hstruct(lo, hi):
            b = _abi.Batch()
            b.first_index = first_index + lo
            b.n = hi - lo
            b.flags = xflags[0] | READ_LEN                  # bits 0-15: the longest read (saves the engine a pass over the offsets)
            b.seq1 = host["seq1"].data_ptr(); b.qual1 = host["qual1"].data_ptr(); b.off1 = host["off1"].data_ptr() + 4 * lo
            b.seq2 = host["seq2"].data_ptr(); b.qual2 = host["qual2"].data_ptr(); b.off2 = host["off2"].data_ptr() + 4 * lo
            return b

        off1 = host["off1"].numpy(); off2 = host["off2"].numpy()
        h2d = 2 * int(off1[n] - off1[0]) + 2 * int(off2[n] - off2[0]) + 8 * (n + 1)
        if has_window:
            h2d += 2 * int(off1[s_hi] - off1[s_lo]) + 2 * int(off2[s_hi] - off2[s_lo]) + 8 * (s_hi - s_lo + 1)
        d2h = 32 * n

        def step_e2e():
            if has_window:
                b = hstruct(s_lo, s_hi)
                eng._check(L.aqc_stat_reads(eng._h, C.byref(b), _abi.MEM_HOST, _abi.QC_R1_PRE, _abi.QC_R2_PRE, w_lo_g, w_hi_g, 0))
            b = hstruct(0, n)
            eng._check(L.aqc_filter_pairs(eng._h, C.byref(b), _abi.MEM_HOST, res_host.data_ptr()))

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
            t = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            try:        # same pairs, same parameters: the host-buffer path must return the records of the resident path
                torch.cuda.synchronize()
                same = bool(torch.equal(res_host.to(device, non_blocking=False), wb.results[:n * 32]))
            except Exception:
                same = None
            return float(t.item()) / e_steps, same

        ms_e2e, same = time_e2e()
        e2e = {"value": world * n / (ms_e2e * 1e-3) / 1e6, "unit": "M read-pairs/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": e_steps, "results_match_resident": same,
               "note": "aqc_stat_reads + aqc_filter_pairs with AQC_MEM_HOST on pinned host columns; chunked H2D/kernels/D2H pipeline inside"}

    # ---------------- roofline of the dominant kernel ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak = float(json.load(f)["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    off1 = wb.t["off1"]; off2 = wb.t["off2"]
    col_bytes = 2 * int(off1[n].item() - off1[0].item()) + 2 * int(off2[n].item() - off2[0].item())
    algo_bytes = col_bytes + 32 * n + 8 * n
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_filter_kernel_dram_bytes.json")     # ncu capture of pair_kernel
    if os.path.exists(tp) and not use_lane:
        try:
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_pair"] * n     # ncu capture was taken on a smaller launch; per-pair traffic x pairs/launch
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "%s, %d pairs/launch" % (kernel_label, n), "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                "note": "integer ALU-bound kernel: see DESIGN.md (instruction budget per pair vs HBM budget)"}

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu:
            threads = host_threads()
            sample = cpu_sample_size(n, threads, args.cpu_sample)
            # the first `sample` pairs of the resident batch, copied back to the host
            from afterqc_b200.batch import PackedBatch, SLACK
            o1 = wb.t["off1"][:sample + 1].cpu().numpy().astype(np.uint32); o2 = wb.t["off2"][:sample + 1].cpu().numpy().astype(np.uint32)

            def hostcol(name, end):
                a = np.zeros(end + SLACK, dtype=np.uint8)
                a[:end] = wb.t[name][:end].cpu().numpy()
                return a
            hb = PackedBatch(hostcol("seq1", int(o1[-1])), hostcol("qual1", int(o1[-1])), o1,
                             hostcol("seq2", int(o2[-1])), hostcol("qual2", int(o2[-1])), o2)
            r = cpu_arm(hb, threads)
            cpu = {"value": r["pairs_per_s"] / 1e6, "unit": "M read-pairs/s", "cores": r["threads"], "kind": "port",
                   "sample": "%d pairs of the same PE150 workload (QC window scaled to %d reads, the same 2%% mix), %.1f s of wall time on %d threads; "
                             "C oracle oracle/aqc_oracle.c" % (sample, r["qc_sample_scaled"], r["seconds"], r["threads"])}
        line = {
            "metric": "read-pairs/s PE150 (filter + overlap correction + adapter trim + per-cycle QC)",
            "value": value, "unit": "M read-pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": n, "read_len": READ_LEN, "qc_sample": QS,
                       "filter_kernel": selection,
                       "l2": "inputs (%.1f GB/GPU) exceed the 126 MB L2; no explicit flush" % (col_bytes / 1e9),
                       "parallelism": "read-sharded x%d, NCCL all-reduce of the counter blocks per step" % world if world > 1 else "1 GPU"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
    # ---------------- other transport modes of the same host-buffer calls (identical results, fewer PCIe bytes) ----------------
    # Run LAST: everything the JSON line needs exists by now, so an experimental mode that fails costs only itself.
    #   qual2_in_place (lane kernels, only if the child check saw it work on this GPU): mate-2 qualities stay in the pinned host
    #     column; the kernel fetches the bytes of the correction walk / sampled statRead over PCIe
    #   pack_bases / pack_quals: host threads pack the base columns to 2 bits per base (the quality columns to 6 bits per byte),
    #     unpack_bases_kernel / unpack_quals_kernel restore the bytes in HBM
    if e2e is not None:
        try_in_place = use_lane and (args.filter_kernel in ("lane", "lane2") or bool((child or {}).get("in_place_ok")))
        if world > 1:
            flag = torch.tensor([1 if try_in_place else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            try_in_place = bool(flag.item())
        modes = []
        if try_in_place:
            modes.append(("qual2_in_place", _abi.BATCH_QUAL2_IN_PLACE))
        if not args.no_pack:
            PB, PQ = _abi.BATCH_PACK_BASES, _abi.BATCH_PACK_QUALS
            modes += [("pack_bases", PB), ("pack_bases+pack_quals", PB | PQ)]
            if try_in_place:
                modes += [("pack_bases+qual2_in_place", PB | _abi.BATCH_QUAL2_IN_PLACE),
                          ("pack_bases+pack_quals+qual2_in_place", PB | PQ | _abi.BATCH_QUAL2_IN_PLACE)]
        variants = {"copy_all_columns": {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"], "h2d_bytes_per_step": h2d}}
        best = ("copy_all_columns", e2e["value"])
        try:
            q2_bytes = int(off2[n] - off2[0])
            base_bytes = int(off1[n] - off1[0]) + int(off2[n] - off2[0])
            n_edits = int(wb.results[:n * 32].view(-1, 32)[:, 1].sum().item())
            pulled = 32 * n_edits + min(n, QS) * READ_LEN      # one 32-byte sector per visited mismatch + the stat'd reads
            n_exc, q_exc = 0, {}
            for k in ("seq1", "seq2", "qual1", "qual2"):        # bytes that travel in the exception lists (5 bytes each)
                c = wb.t[k][:int(wb.t["off1" if k.endswith("1") else "off2"][n].item())]
                if k.startswith("seq"):
                    n_exc += int(((c != 65) & (c != 67) & (c != 71) & (c != 84)).sum().item())
                else:
                    q_exc[k] = int(((c < 33) | (c > 96)).sum().item())
            q1_bytes = int(off1[n] - off1[0])
            for name, fl in modes:
                xflags[0] = fl
                res_host.zero_()
                ms_v, same_v = time_e2e()
                xflags[0] = 0
                hb = h2d
                if fl & _abi.BATCH_QUAL2_IN_PLACE:
                    hb += pulled - q2_bytes
                if fl & _abi.BATCH_PACK_BASES:
                    hb += -base_bytes + (base_bytes + 3) // 4 + 5 * n_exc
                if fl & _abi.BATCH_PACK_QUALS:
                    hb += -(q1_bytes // 4) + 5 * q_exc["qual1"]
                    if not (fl & _abi.BATCH_QUAL2_IN_PLACE):
                        hb += -(q2_bytes // 4) + 5 * q_exc["qual2"]
                variants[name] = {"value": world * n / (ms_v * 1e-3) / 1e6, "ms_per_step": ms_v, "results_match_resident": same_v,
                                  "h2d_bytes_per_step": hb}
                if same_v and variants[name]["value"] > best[1]:
                    best = (name, variants[name]["value"])
        except Exception as ex:      # noqa: BLE001
            xflags[0] = 0
            variants["error"] = repr(ex)[:300]
        e2e["variants"] = variants
        e2e["mode"] = best[0]
        if best[0] != "copy_all_columns":
            v = variants[best[0]]
            e2e.update({"value": v["value"], "ms_per_step": v["ms_per_step"], "h2d_bytes_per_step": v["h2d_bytes_per_step"],
                        "results_match_resident": v["results_match_resident"]})
            e2e["note"] += ("; mode %s: AQC_BATCH_QUAL2_IN_PLACE = the qual2 column is not copied, the kernel reads what the correction walk and "
                            "the sampled statRead need from the pinned column over PCIe (estimate in h2d_bytes_per_step); AQC_BATCH_PACK_BASES = "
                            "host threads pack the bases to 2 bits before the copy (+ 5 bytes per byte that is not A,C,G,T); AQC_BATCH_PACK_QUALS = ... and "
                            "the qualities to 6 bits" % best[0])
        if line is not None:
            line["e2e"] = e2e
        del host, res_host
    try:
        eng.close()
    except Exception:       # noqa: BLE001
        pass
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
