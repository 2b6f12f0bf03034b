"""Synthetic barcoded (UMI) reads: barcode (design length +-1) + verify sequence + insert; templates about one read long
so that cleanBarcodeTail (barcodeprocesser.py:48-74) finds read-through tails; a few short reads and names without a colon."""
import random, os
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
def rc(s): return "".join(COMP[c] for c in reversed(s))
def rnd(rng, n, al="ACGT"): return "".join(rng.choice(al) for _ in range(n))
def mutate(rng, s, p):
    return "".join(rng.choice("ACGTN") if rng.random() < p else c for c in s)
def make(n, seed, L=100, verify="CAGTA", blen=12, paired=True, colon=True):
    rng = random.Random(seed)
    r1s, r2s = [], []
    for i in range(n):
        b1 = rnd(rng, blen + rng.choice([0, 0, 0, 0, -1, 1]))
        b2 = rnd(rng, blen + rng.choice([0, 0, 0, 0, -1, 1]))
        v1 = mutate(rng, verify, 0.08); v2 = mutate(rng, verify, 0.08)
        kind = rng.random()
        if kind < 0.5:
            ins = rnd(rng, rng.randint(150, 300))
        elif kind < 0.9:
            ins = rnd(rng, rng.randint(L - 2 * (blen + 5) - 12, L - 2 * (blen + 5) + 12))     # template about one read long
        else:
            ins = rnd(rng, rng.randint(5, 40))
        frag = b1 + v1 + ins + rc(v2) + rc(b2)
        s1 = mutate(rng, (frag + rnd(rng, L))[:L], 0.004)
        s2 = mutate(rng, (rc(frag) + rnd(rng, L))[:L], 0.004)
        if rng.random() < 0.03: s1 = s1[:rng.randint(5, 25)]
        if rng.random() < 0.03: s2 = s2[:rng.randint(5, 25)]
        q1 = "".join(rng.choice("#5AFII") for _ in s1); q2 = "".join(rng.choice("#5AFII") for _ in s2)
        nm = ("@INST:1:FC:1:%d:%d:%d" % (i % 7, i, i)) if (colon or i % 5) else ("@nocolon%d" % i)
        r1s.append((nm + " 1:N:0", s1, q1)); r2s.append((nm + " 2:N:0", s2, q2))
    return r1s, (r2s if paired else None)
def write(d, subs, r1s, r2s, stem="x_barcode"):
    for sub in subs:
        os.makedirs(os.path.join(d, sub), exist_ok=True)
        for m, recs in ((1, r1s), (2, r2s)):
            if recs is None: continue
            with open(os.path.join(d, sub, "%s_R%d.fq" % (stem, m)), "w") as f:
                for nm, s, q in recs:
                    f.write("%s\n%s\n+\n%s\n" % (nm, s, q))
