#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump per CUDA source line of
this repo.  SASS that the compiler attributes to CUDA header intrinsics (ballot, shfl, atomics) is folded into
the nearest preceding repo line by address.
usage: python tools/ncu_lines.py dump.csv <kernel_index> [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
rows = list(csv.reader(open(path, newline="")))
kern = -1
seen_files = set()
fname = None
hdr = None
cur_line = None
cur_text = ""
sass = []   # (addr, inst, samples, file, line, text, sasstext)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        if fname in seen_files or kern < 0:
            kern += 1
            seen_files = set()
        seen_files.add(fname)
        continue
    if r[0] in ("Function Name",):
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
        continue
    if kern != want or hdr is None:
        continue
    if r[0].strip().isdigit():
        cur_line = int(r[0]); cur_text = r[1]
        continue
    if r[0] == "" and len(r) > ie and r[2].startswith("0x"):
        try:
            sass.append((int(r[2], 16), int(r[ie] or 0), int(r[sm] or 0), fname, cur_line, cur_text, r[3].strip()))
        except ValueError:
            pass
sass.sort()
agg = defaultdict(lambda: [0, 0, ""])
last = ("?", 0, "")
mine = lambda f: f.startswith("aqc_")
tot = tots = 0
for addr, inst, smp, f, l, text, st in sass:
    if mine(f):
        last = (f, l, text)
    key = last
    a = agg[(key[0], key[1])]
    a[0] += inst; a[1] += smp; a[2] = key[2]
    tot += inst; tots += smp
print("kernel %d: %d SASS rows, %d warp-instructions, %d samples" % (want, len(sass), tot, tots))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * a[0] / max(tot, 1), 100.0 * a[1] / max(tots, 1), f, l, a[2].strip()[:100]))
