#!/usr/bin/env python
"""Per-CUDA-source-line summary of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--launch-skip k --launch-count 1]`:
share of warp instructions, share of stall samples, top stall reasons.   usage: ncu_src_summary.py <csv> [top_n]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if len(r) > 8 and r[0] == "Line No" and "Instructions Executed" in r]
h = rows[hi[0]]
ie, it, isamp = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
stall_cols = {n: i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n}
src, agg = {}, collections.OrderedDict()
for r in rows[hi[0] + 1: hi[1] if len(hi) > 1 else None]:
    if len(r) <= isamp:
        continue
    if not r[2]:            # pure source row (no SASS address): remember the text only
        src[r[0]] = r[1].strip()
        continue
    try:
        inst, thr, smp = float(r[ie] or 0), float(r[it] or 0), float(r[isamp] or 0)
    except ValueError:
        continue
    a = agg.setdefault(r[0], [0.0, 0.0, 0.0, collections.Counter()])
    a[0] += inst; a[1] += thr; a[2] += smp
    for n, i in stall_cols.items():
        try:
            a[3][n[6:]] += float(r[i] or 0)
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"warp instructions {tot:.0f}, stall samples {tots:.0f}")
for ln, (inst, thr, smp, st) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    reasons = ", ".join(f"{k}:{v / max(smp, 1) * 100:.0f}%" for k, v in st.most_common(3))
    print(f"{smp / max(tots, 1) * 100:5.1f}% smp {inst / max(tot, 1) * 100:5.1f}% inst thr/inst {thr / max(inst, 1):4.1f} L{ln}: {src.get(ln, '')[:64]} | {reasons}")
