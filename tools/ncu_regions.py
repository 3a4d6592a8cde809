#!/usr/bin/env python
"""Aggregate an .ncu-rep source page by code region of the tile force kernels (instructions + stall
samples).  usage: ncu_regions.py rep [tile|tile2]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 else "tile2"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(k for k, r in enumerate(rows) if "Source" in r and len(r) > 5)
hh = rows[hi]
ii, st = hh.index("Instructions Executed"), hh.index("Warp Stall Sampling (All Samples)")
D = "cellular_raza_b200/csrc/"
files = {f: open(D + f).read().splitlines() for f in ("crb_force_tile.cuh", "crb_force_tile2.cuh", "crb_force_tile3.cuh", "crb_kernels.cuh")}
mainname = {"tile2": "crb_force_tile2.cuh", "tile3": "crb_force_tile3.cuh"}.get(which, "crb_force_tile.cuh")
main = files[mainname]
def find(pat):
    return next(i + 1 for i, l in enumerate(main) if pat in l)
if which == "tile3":
    marks = [("prologue", find("const KParams P, const TileLaunch T)")), ("window_search", find("window size: largest kc")),
             ("row_ranges", find("---- row ranges")), ("staging", find("---- stage the candidates")),
             ("target_setup", find("---- targets:")), ("filter_setup(A frags)", find("---- filter phase")),
             ("filter_rows", find("while (r < 9 && !full)")), ("filter_mma", find("for (int s = 0; s < 4; ++s)")),
             ("filter_reduce_push", find("// OR over the quad")), ("drain_ctl", find("---- drain (FP64, exact)")),
             ("epilogue", find("if (valid) {"))]
else:
    marks = [("prologue", find("const KParams P, const TileLaunch T)")), ("window_search", find("window size: largest kc")),
         ("row_ranges", find("---- row ranges")), ("staging", find("---- stage the candidates")),
         ("target_setup", find("---- targets:")), ("drain_ctl", find("auto drain = [&]()")),
         ("filter", find("for (int r = 0; r < 9; ++r)" if which == "tile2" else "for (int r = 0; r < kTileRows; ++r) {")), ("epilogue", find("if (valid) {"))]
def at(f, ln, t):
    L = files[f]
    return 0 < ln <= len(L) and L[ln - 1].strip() == t
def region(ln, text):
    t = text.strip()
    mainf = mainname
    if at(mainf, ln, t) and not at("crb_kernels.cuh", ln, t) and not (which in ("tile2", "tile3") and at("crb_force_tile.cuh", ln, t) and 60 < ln < 135):
        if ln < marks[0][1]:
            if which == "tile3": return "filter_mma"
            return "eval(tile_eval)" if (which != "tile2" and ln > 60) else "other"
        r = "prologue"
        for name, start in marks:
            if ln >= start: r = name
        return r
    if which in ("tile2", "tile3") and at("crb_force_tile.cuh", ln, t) and ln < 200:
        return "eval(tile_eval)"
    if at("crb_kernels.cuh", ln, t):
        return "eval(coef/rsqrt)" if ln < 260 else "epilogue(integrate)"
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
