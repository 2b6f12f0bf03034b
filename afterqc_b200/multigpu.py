"""Read-sharded multi-GPU execution.

Pairs are independent, so rank r filters the contiguous record range [n*r/W, n*(r+1)/W) with its own
engine and NO data-path collective.  The only exchange is the reduction of the accumulator blocks
(scalar counters, histograms, error matrix, per-cycle QC counters, dense k-mer tables) -- SUM, and MIN for
the k-mer first-seen stamps -- done once per phase over torch.distributed (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Rank 0 then writes the JSON; good/bad outputs are written per rank and
concatenated in rank order, which reproduces the single-process files byte for byte.
"""
import os
import shutil
import tempfile

import numpy as np

from . import _abi


class DistBackend:
    """Wraps a local backend (Engine or, in tests, the oracle); fetches return GLOBAL (all-reduced) values."""

    def __init__(self, local, group=None, device=None):
        import torch.distributed as dist
        self.local = local
        self.dist = dist
        self.group = group
        self.device = device          # torch device for the collective buffers ("cuda:N" for NCCL, None/cpu for gloo)
        self.params = local.params

    # ---- pass-through of the per-shard work ----
    def set_params(self, p):
        self.params = p
        self.local.set_params(p)

    def stat_reads(self, *a, **k):
        return self.local.stat_reads(*a, **k)

    def filter_pairs(self, *a, **k):
        return self.local.filter_pairs(*a, **k)

    def reset_filter_counters(self):
        self.local.reset_filter_counters()

    def close(self):
        self.local.close()

    # ---- reductions ----
    def _allreduce(self, arr, op):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy())
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MIN, group=self.group)
        return t.cpu().numpy()

    def counters(self):
        return self._allreduce(self.local.counters(), "sum")

    def qc(self, slot):
        rec = self.local.qc(slot)
        flat = np.frombuffer(rec.tobytes(), dtype=np.int64)
        out = self._allreduce(flat, "sum")
        return np.frombuffer(out.tobytes(), dtype=_abi.QC_DTYPE)[0]

    def kmers(self, slot):
        cnt, first, skeys, scnt, sfirst = self.local.kmers(slot)
        cnt = self._allreduce(cnt, "sum").view(np.uint64)
        # stamps are unsigned; MIN over int64 views is order-preserving only below 2^63, NEVER (all ones) is -1:
        # shift into the signed range first
        bias = np.uint64(1) << np.uint64(63)
        f = self._allreduce((first ^ bias), "min").view(np.uint64) ^ bias
        # side tables (k-mers with a byte outside A,C,G,T): gathered and merged by key
        world = self.dist.get_world_size(group=self.group)
        gathered = [None] * world
        if hasattr(self.local, "kmer_side_raw"):
            # the engine's unresolved table: counts add, direct sightings and foreign-byte seedings take the minimum, and the
            # insertion rule is applied to the merged table -- exact for every byte (quirk Q12)
            self.dist.all_gather_object(gathered, self.local.kmer_side_raw(slot), group=self.group)
            keys = np.concatenate([g[0] for g in gathered])
            if not len(keys):
                return cnt, f, skeys, scnt, sfirst
            never = np.iinfo(np.uint64).max
            uk, inv = np.unique(keys, return_inverse=True)
            c2 = np.zeros(len(uk), dtype=np.uint64); np.add.at(c2, inv, np.concatenate([g[1] for g in gathered]))
            d2 = np.full(len(uk), never, dtype=np.uint64); np.minimum.at(d2, inv, np.concatenate([g[2] for g in gathered]))
            s2 = np.full(len(uk), never, dtype=np.uint64); np.minimum.at(s2, inv, np.concatenate([g[3] for g in gathered]))
            stamp = resolve_side(uk, d2, s2, self.params.qc_kmer)
            keep = stamp != never
            return cnt, f, uk[keep], c2[keep], stamp[keep]
        # a backend that only exposes resolved stamps (the oracle in the CPU tests): minimum of the resolved stamps, exact for
        # k-mers over util.COMP's alphabet
        self.dist.all_gather_object(gathered, (skeys, scnt, sfirst), group=self.group)
        keys = np.concatenate([g[0] for g in gathered])
        if len(keys):
            cs = np.concatenate([g[1] for g in gathered]); fs = np.concatenate([g[2] for g in gathered])
            uk, inv = np.unique(keys, return_inverse=True)
            c2 = np.zeros(len(uk), dtype=np.uint64); np.add.at(c2, inv, cs)
            f2 = np.full(len(uk), np.iinfo(np.uint64).max, dtype=np.uint64); np.minimum.at(f2, inv, fs)
            return cnt, f, uk, c2, f2
        return cnt, f, skeys, scnt, sfirst


_COMP = {65: 84, 84: 65, 67: 71, 71: 67, 97: 116, 116: 97, 99: 103, 103: 99, 78: 78, 10: 10}      # util.py:27 as bytes


def resolve_side(keys, direct, seed, k):
    """Insertion stamps of the side-table k-mers from their first direct sighting and their first seeding by a k-mer that
    holds a byte outside util.COMP -- the rule of aqc_get_kmer_side (csrc/aqc_engine.cu; derivation above stat_read in
    csrc/aqc_device.cuh), here on a table merged over shards.  keys: sorted, unique, k raw bytes big-endian."""
    never = int(np.iinfo(np.uint64).max)
    pos = {int(x): i for i, x in enumerate(keys)}
    out = np.full(len(keys), never, dtype=np.uint64)
    for i, key in enumerate(keys):
        key = int(key)
        bs = [(key >> (8 * (k - 1 - j))) & 0xFF for j in range(k)]
        foreign = any(b not in _COMP for b in bs)
        di, si = int(direct[i]), int(seed[i])
        if foreign:                                   # no pre-image under reverseComplement: present from its first sighting
            p = di
        else:
            p = min(di, si)
            rkey = 0
            for j, b in enumerate(bs):
                rkey |= _COMP[b] << (8 * j)
            if rkey != key:
                jx = pos.get(rkey)
                if jx is not None and int(direct[jx]) < di and int(direct[jx]) < int(seed[jx]):
                    p = min(p, int(direct[jx]) | 1)   # the partner was sighted first and newly inserted: it seeded this one
        out[i] = p
    return out


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the page-locked columns it allocates
    and the threads that fill them are local to that GPU's PCIe root.  Returns a short description; a box with one node (or
    without the sysfs entries) is left alone."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown (single node)"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, allowed)
        return "numa node %d, %d cpus" % (node, len(allowed))
    except Exception as e:       # noqa: BLE001
        return "not bound (%s)" % type(e).__name__


def shard_range(n, rank, world):
    return (n * rank) // world, (n * (rank + 1)) // world


def concat_outputs(paths, dest):
    """Concatenate per-rank output files (plain or gzip members) in rank order."""
    with open(dest, "wb") as out:
        for p in paths:
            with open(p, "rb") as f:
                shutil.copyfileobj(f, out)


def _resolve(spec):
    mod, attr = spec.split(":")
    import importlib
    return getattr(importlib.import_module(mod), attr)


def _worker(rank, world, options, port, backend_name, result_q, local_factory=None):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import datetime
    tmo = datetime.timedelta(seconds=int(os.environ.get("AQC_DIST_TIMEOUT", "1800")))
    if backend_name == "nccl":
        torch.cuda.set_device(rank)
        bind_to_gpu_numa_node(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank), timeout=tmo)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=tmo)
    from .pipeline import seqFilter, default_backend

    def factory(params):
        if backend_name == "nccl":
            from .engine import Engine
            return DistBackend(Engine(params, device=rank), device="cuda:%d" % rank)
        if local_factory is not None:        # tests: "module:callable" building the per-rank backend
            return DistBackend(_resolve(local_factory)(params))
        return DistBackend(default_backend(params))
    sf = seqFilter(options, backend_factory=factory, shard=(rank, world))
    sf.run()
    dist.barrier()
    dist.destroy_process_group()
    if result_q is not None and rank == 0:
        result_q.put(True)


def free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_sharded(options, gpus, port=None, timeout_s=None):
    """after.py --gpus N: one process per GPU of this box.  The rendezvous port is probed (directory mode may run several
    sharded jobs one after another); when one shard process dies the others are terminated instead of waiting in a
    collective for ever."""
    import time
    import torch.multiprocessing as mp
    port = port or free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, gpus, options, port, "nccl", None)) for r in range(gpus)]
    for p in procs:
        p.start()
    t0 = time.time()
    failed = False
    while any(p.is_alive() for p in procs):
        if any((not p.is_alive()) and p.exitcode != 0 for p in procs) or (timeout_s and time.time() - t0 > timeout_s):
            failed = True
            break
        time.sleep(0.05)
    if failed:
        for p in procs:
            if p.is_alive():
                p.terminate()
    for p in procs:
        p.join()
    if failed or any(p.exitcode != 0 for p in procs):
        raise RuntimeError("a shard process failed")
