#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
warp-stall samples, executed warp instructions.  usage: srcagg.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
cur = None
for r in rows:
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] not in ("", "-"):          # a source line summary row
        try:
            ln = int(r[0])
        except ValueError:
            continue
        cur = (ln, r[1].strip())
        try:
            smp, inst = int(r[4]), int(r[7])
        except ValueError:
            continue
        a = agg.setdefault(cur, [0, 0])
        a[0] += smp
        a[1] += inst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
for (ln, src), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  L%-4d %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, ln, src[:110]))
