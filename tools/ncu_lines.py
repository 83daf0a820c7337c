#!/usr/bin/env python
"""Per-source-line digest of an ncu report: python tools/ncu_lines.py <rep> <launch-skip> [top]"""
import csv, subprocess, sys, io
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if len(r) > 8 and r[0] == "Line No" and "Instructions Executed" in r]
h = rows[hi[0]]
ie, isamp, it = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
stall = {n: i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n}
def f(x):
    try: return float(x)
    except ValueError: return 0.0
src = []
for r in rows[hi[0] + 1:hi[1] if len(hi) > 1 else None]:
    if len(r) > isamp and r[2] == '-' and r[0]:
        st = sorted(((f(r[i]), n[6:]) for n, i in stall.items()), reverse=True)[:2]
        src.append((r[0], r[1].strip()[:70], f(r[ie]), f(r[it]), f(r[isamp]), st))
tot = sum(s[2] for s in src); ts = sum(s[4] for s in src)
print("total warp inst %.0f, samples %.0f" % (tot, ts))
for s in sorted(src, key=lambda s: -s[4])[:top]:
    print("%5.1f%% smp %5.1f%% inst thr %4.1f L%s: %s | %s" % (s[4] / ts * 100, s[2] / tot * 100, s[3] / max(s[2], 1), s[0], s[1],
          ", ".join("%s:%.0f%%" % (n, v / max(s[4], 1) * 100) for v, n in s[5])))
