#!/usr/bin/env python
"""Top CUDA source lines of an .ncu-rep by executed warp instructions. usage: tools_ncu_lines.py rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(k for k, r in enumerate(rows) if "Source" in r and len(r) > 5)
hh = rows[hi]
ii, ti, si, st, av = hh.index("Instructions Executed"), hh.index("Thread Instructions Executed"), 1, hh.index("Warp Stall Sampling (All Samples)"), hh.index("Avg. Threads Executed")
agg = {}
for r in rows[hi + 1:]:
    try:
        key = (r[0], r[si])
        a = agg.setdefault(key, [0.0, 0.0, 0.0, 0])
        a[0] += float(r[ii] or 0); a[1] += float(r[ti] or 0); a[2] += float(r[st] or 0); a[3] += 1
    except Exception:
        pass
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"total warp instr {tot:.0f}, stall samples {tots:.0f}")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{a[0]/tot*100:5.1f}% inst  {a[2]/max(tots,1)*100:5.1f}% stall  lanes {a[1]/max(a[0],1):5.1f}  sass {a[3]:3d}  L{ln}: {src.strip()[:90]}")
