#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line.
usage: ncu_src_summary.py <csv> [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if len(r) > 8 and r[0] == "Line No" and "Instructions Executed" in r]
h = rows[hi[0]]
ie, it, isamp = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
agg = collections.OrderedDict()
for r in rows[hi[0] + 1: hi[1] if len(hi) > 1 else None]:
    if len(r) <= it:
        continue
    try:
        inst, thr, smp = float(r[ie]), float(r[it]), float(r[isamp] or 0)
    except ValueError:
        continue
    key = (r[0], r[1].strip())
    a = agg.setdefault(key, [0.0, 0.0, 0.0])
    a[0] += inst; a[1] += thr; a[2] += smp
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"total warp instructions {tot:.0f}, samples {tots:.0f}")
for (ln, src), (inst, thr, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{inst / tot * 100:5.1f}% inst {smp / max(tots, 1) * 100:5.1f}% smp  thr/inst {thr / max(inst, 1):5.1f}  L{ln}: {src[:100]}")
