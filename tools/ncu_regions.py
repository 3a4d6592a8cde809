#!/usr/bin/env python
"""Aggregate an .ncu-rep source page by code region of k_force_tile (instructions + stall samples)."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(k for k, r in enumerate(rows) if "Source" in r and len(r) > 5)
hh = rows[hi]
ii, st = hh.index("Instructions Executed"), hh.index("Warp Stall Sampling (All Samples)")
files = {}
for f in ("cellular_raza_b200/csrc/crb_force_tile.cuh", "cellular_raza_b200/csrc/crb_kernels.cuh"):
    files[f] = open(f).read().splitlines()
tile = files["cellular_raza_b200/csrc/crb_force_tile.cuh"]
def find(pat):
    return next(i + 1 for i, l in enumerate(tile) if pat in l)
marks = [("prologue", find("k_force_tile(const KParams")), ("window_search", find("window size: largest kc")),
         ("row_ranges", find("---- row ranges")), ("staging", find("---- stage the candidates")),
         ("target_setup", find("---- targets:")), ("drain_ctl", find("auto drain = [&]()")),
         ("filter", find("for (int r = 0; r < kTileRows; ++r) {")), ("epilogue", find("if (valid) {"))]
def region(ln, text):
    t = text.strip()
    in_tile = 0 < ln <= len(tile) and tile[ln - 1].strip() == t
    k = files["cellular_raza_b200/csrc/crb_kernels.cuh"]
    in_k = 0 < ln <= len(k) and k[ln - 1].strip() == t
    if in_tile and not in_k:
        if ln < marks[0][1]:
            return "eval(tile_eval)" if ln > 60 else "other"
        r = "prologue"
        for name, start in marks:
            if ln >= start: r = name
        return r
    if in_k:
        if ln < 250: return "eval(coef/rsqrt)"
        return "epilogue(integrate)"
    if "fadd2" in t or "ffma2" in t or "fmul2" in t: return "filter"
    if "reduce_max" in t or "vote_any" in t: return "filter"
    return "other:" + t[:30]
agg = {}
for r in rows[hi + 1:]:
    try:
        ln = int(r[0]); key = region(ln, r[1])
        a = agg.setdefault(key, [0.0, 0.0]); a[0] += float(r[ii] or 0); a[1] += float(r[st] or 0)
    except Exception:
        pass
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp instr {ti:.0f}  stall samples {ts:.0f}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[0] / ti > 0.002 or a[1] / ts > 0.002:
        print(f"  {k:28s} inst {a[0]/ti*100:5.1f}%  time(samples) {a[1]/ts*100:5.1f}%")
