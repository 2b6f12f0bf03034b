"""Seeded synthetic FASTQ batches for the BASELINE.json configs.

torch is used only as a vectorised RNG (CPU for tests, CUDA for the full-size bench so that
10 M pairs are generated in seconds); the generated columns are plain uint8/uint32 arrays.

Model (config 3, PE150): fragment length ~ N(260, 70) clipped to [40, 600]; R1 = fragment
head, R2 = reverse complement of the fragment tail; when the fragment is shorter than the
read the remainder is random "adapter"; per-base substitution error `err` whose quality is
Q2..Q14 with probability 0.8 (else Q30..Q40); normal bases Q30..Q40; N rate `n_rate` (Q2);
`polyx_rate` of the pairs get a >= 35 long homopolymer tail with <= 2 interruptions;
`lowq_rate` of the reads get > 60 low-quality positions; `nrich_rate` get 6..20 N.
"""
import numpy as np
import torch

from .batch import PackedBatch, SLACK

_BASES = torch.tensor([ord(c) for c in "ACGT"], dtype=torch.uint8)

CONFIGS = {
    # name: dict(L, paired, frag_mean, frag_sd, frag_min, frag_max, err, seed)
    "se100": dict(L=100, paired=False, frag_mean=0, frag_sd=0, frag_min=0, frag_max=0, err=0.0, seed=20260926,
                  lowq_rate=0.01, polyx_rate=0.005, nrich_rate=0.01, n_rate=0.0),
    "pe150": dict(L=150, paired=True, frag_mean=260, frag_sd=70, frag_min=40, frag_max=600, err=0.01, seed=20260927,
                  lowq_rate=0.0, polyx_rate=0.002, nrich_rate=0.0, n_rate=0.001),
    "pe250": dict(L=250, paired=True, frag_mean=420, frag_sd=120, frag_min=60, frag_max=900, err=0.01, seed=20260928,
                  lowq_rate=0.0, polyx_rate=0.002, nrich_rate=0.0, n_rate=0.001),
    "pe150_err3": dict(L=150, paired=True, frag_mean=260, frag_sd=70, frag_min=40, frag_max=600, err=0.03, seed=20260929,
                       lowq_rate=0.0, polyx_rate=0.002, nrich_rate=0.0, n_rate=0.001),
}


def _gen_chunk(n, cfg, gen, device):
    L = cfg["L"]
    paired = cfg["paired"]
    ar = torch.arange(L, device=device)

    def rnd(*shape):
        return torch.rand(*shape, generator=gen, device=device)

    def rint(lo, hi, *shape):
        return torch.randint(lo, hi, shape, generator=gen, device=device)

    def quals_and_errors(codes):
        """apply substitution errors / N / quality model to a code matrix [n, L] (0..3) -> (bases, quals)"""
        q = rint(30, 41, n, L)
        if cfg["err"] > 0:
            e = rnd(n, L) < cfg["err"]
            codes = torch.where(e, (codes + rint(1, 4, n, L)) % 4, codes)
            lowq = e & (rnd(n, L) < 0.8)
            q = torch.where(lowq, rint(2, 15, n, L), q)
        bases = _BASES.to(device)[codes]
        if cfg["n_rate"] > 0:
            isn = rnd(n, L) < cfg["n_rate"]
            bases = torch.where(isn, torch.full_like(bases, ord("N")), bases)
            q = torch.where(isn, torch.full_like(q, 2), q)
        return bases, q

    def decorate(bases, q):
        """read-level artefacts: low-quality reads, polyX tails, N-rich reads"""
        if cfg["lowq_rate"] > 0:
            sel = rnd(n) < cfg["lowq_rate"]
            npos = rint(61, max(62, L - 5), n)
            score = rnd(n, L)
            kth = torch.sort(score, dim=1).values.gather(1, (npos.clamp(max=L - 1)).unsqueeze(1))
            m = sel.unsqueeze(1) & (score < kth)
            q = torch.where(m, rint(2, 15, n, L), q)
        if cfg["polyx_rate"] > 0:
            sel = rnd(n) < cfg["polyx_rate"]
            run = rint(35, min(L, 80) + 1, n)
            kind = rnd(n)
            pb = _BASES.to(device)[rint(0, 4, n)]
            pb = torch.where(kind < 0.25, torch.full_like(pb, ord("G")), pb)
            pb = torch.where((kind >= 0.25) & (kind < 0.4), torch.full_like(pb, ord("N")), pb)
            tail = ar.unsqueeze(0) >= (L - run).unsqueeze(1)
            # <= 2 interruptions inside the tail
            keep = torch.ones(n, L, dtype=torch.bool, device=device)
            for _ in range(2):
                pos = (L - run + (rnd(n) * run.float()).long().clamp(max=L - 1)).clamp(max=L - 1)
                hit = rnd(n) < 0.5
                keep[torch.arange(n, device=device)[hit], pos[hit]] = False
            m = sel.unsqueeze(1) & tail & keep
            bases = torch.where(m, pb.unsqueeze(1).expand(n, L), bases)
        if cfg["nrich_rate"] > 0:
            sel = rnd(n) < cfg["nrich_rate"]
            cnt = rint(6, 21, n)
            score = rnd(n, L)
            kth = torch.sort(score, dim=1).values.gather(1, cnt.unsqueeze(1))
            m = sel.unsqueeze(1) & (score < kth)
            bases = torch.where(m, torch.full_like(bases, ord("N")), bases)
            q = torch.where(m, torch.full_like(q, 2), q)
        return bases, (q + 33).to(torch.uint8)

    if not paired:
        b1, q1 = quals_and_errors(rint(0, 4, n, L))
        b1, q1 = decorate(b1, q1)
        return b1, q1, None, None

    fmax = cfg["frag_max"]
    flen = (torch.randn(n, generator=gen, device=device) * cfg["frag_sd"] + cfg["frag_mean"]).round().long()
    flen = flen.clamp(cfg["frag_min"], fmax)
    frag = rint(0, 4, n, fmax)
    adapter1 = rint(0, 4, n, L)
    adapter2 = rint(0, 4, n, L)
    inside = ar.unsqueeze(0) < flen.unsqueeze(1)
    c1 = torch.where(inside, frag[:, :L] if fmax >= L else torch.nn.functional.pad(frag, (0, L - fmax)), adapter1)
    idx = (flen.unsqueeze(1) - 1 - ar.unsqueeze(0)).clamp(min=0)
    c2 = torch.where(inside, 3 - frag.gather(1, idx), adapter2)   # codes ACGT: complement = 3 - code
    b1, q1 = quals_and_errors(c1)
    b2, q2 = quals_and_errors(c2)
    b1, q1 = decorate(b1, q1)
    # polyX / lowq on mate 2 only via its own draw (kept simple: decorate again with fresh draws)
    b2, q2 = decorate(b2, q2)
    return b1, q1, b2, q2


def generate_device(config, n, device="cpu", seed=None, chunk=1 << 20, len_jitter=0):
    """Return dict of torch tensors on `device`: seq1, qual1, off1 (int64 -> uint32 view by caller) [, seq2, qual2, off2].

    len_jitter > 0 shortens each read by U[0, len_jitter] bases (variable-length SoA)."""
    cfg = dict(CONFIGS[config]) if isinstance(config, str) else dict(config)
    gen = torch.Generator(device=device)
    gen.manual_seed(cfg["seed"] if seed is None else seed)
    L = cfg["L"]
    parts = {k: [] for k in ("b1", "q1", "b2", "q2", "l1", "l2")}
    done = 0
    while done < n:
        m = min(chunk, n - done)
        b1, q1, b2, q2 = _gen_chunk(m, cfg, gen, device)
        if len_jitter > 0:
            l1 = L - torch.randint(0, len_jitter + 1, (m,), generator=gen, device=device)
            l2 = L - torch.randint(0, len_jitter + 1, (m,), generator=gen, device=device)
        else:
            l1 = torch.full((m,), L, dtype=torch.long, device=device)
            l2 = l1
        ar = torch.arange(L, device=device).unsqueeze(0)
        k1 = ar < l1.unsqueeze(1)
        parts["b1"].append(b1[k1]); parts["q1"].append(q1[k1]); parts["l1"].append(l1)
        if b2 is not None:
            k2 = ar < l2.unsqueeze(1)
            parts["b2"].append(b2[k2]); parts["q2"].append(q2[k2]); parts["l2"].append(l2)
        done += m
    out = {}
    pad = torch.zeros(SLACK, dtype=torch.uint8, device=device)

    def col(xs):
        return torch.cat(xs + [pad])

    def offs(ls):
        l = torch.cat(ls)
        o = torch.zeros(l.numel() + 1, dtype=torch.long, device=device)
        torch.cumsum(l, 0, out=o[1:])
        return o

    out["seq1"] = col(parts["b1"]); out["qual1"] = col(parts["q1"]); out["off1"] = offs(parts["l1"])
    if cfg["paired"]:
        out["seq2"] = col(parts["b2"]); out["qual2"] = col(parts["q2"]); out["off2"] = offs(parts["l2"])
    return out


def generate(config, n, seed=None, first_index=0, len_jitter=0):
    """Host PackedBatch (CPU RNG; deterministic for a given torch version)."""
    t = generate_device(config, n, device="cpu", seed=seed, len_jitter=len_jitter)
    def np8(x):
        return x.numpy().copy()
    def npo(x):
        return x.numpy().astype(np.uint32)
    if "seq2" in t:
        return PackedBatch(np8(t["seq1"]), np8(t["qual1"]), npo(t["off1"]), np8(t["seq2"]), np8(t["qual2"]), npo(t["off2"]),
                           first_index=first_index)
    return PackedBatch(np8(t["seq1"]), np8(t["qual1"]), npo(t["off1"]), first_index=first_index)


def write_fastq(batch, path1, path2=None, name_prefix="SYN:1:FC:1:1101"):
    """Write a PackedBatch as plain FASTQ (Illumina-style names)."""
    import gzip

    def op(p):
        return gzip.open(p, "wb", compresslevel=1) if p.endswith(".gz") else open(p, "wb")

    def dump(path, mate, seq, qual, off):
        with op(path) as f:
            out = []
            for i in range(batch.n):
                g = batch.first_index + i
                a, b = int(off[i]), int(off[i + 1])
                out.append(b"@%s:%d:%d %d:N:0:A\n" % (name_prefix.encode(), g, g, mate))
                out.append(seq[a:b].tobytes()); out.append(b"\n+\n"); out.append(qual[a:b].tobytes()); out.append(b"\n")
                if len(out) > 40000:
                    f.write(b"".join(out)); out = []
            f.write(b"".join(out))

    dump(path1, 1, batch.seq1, batch.qual1, batch.off1)
    if path2 is not None and batch.paired:
        dump(path2, 2, batch.seq2, batch.qual2, batch.off2)
