#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + hottest source lines. usage: tools_ncu_summary.py rep [nlines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h, u = r[0], r[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_fmaheavy.avg.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled", "launch__grid_size", "launch__block_size",
        "smsp__warps_eligible.avg.per_cycle_active", "lts__t_bytes.sum ", "sm__cycles_elapsed.max"]
for row in r[2:]:
    print("== kernel:", row[h.index("Kernel Name")][:80])
    for a, b, c in zip(h, row, u):
        if any((a + " ").startswith(w) or a.startswith(w) for w in want):
            if "stalled" in a and "_per_issue_active" not in a: continue
            print(f"  {a} = {b} {c}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
try:
    rows = list(csv.reader(io.StringIO(src)))
    hi = None
    for k, row in enumerate(rows):
        if "Source" in row and any("Instructions Executed" in x for x in row): hi = k; break
    if hi is not None:
        hh = rows[hi]
        si = hh.index("Source")
        cols = {n: hh.index(n) for n in hh if n in ("# Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Samples)", "Instructions Executed", "Address", "#")}
        sk = [n for n in cols if "Stall" in n][0]
        data = []
        for row in rows[hi + 1:]:
            try: data.append((float(row[cols[sk]] or 0), float(row[cols["Instructions Executed"]] or 0), row[si]))
            except Exception: pass
        tot = sum(d[0] for d in data) or 1
        print(f"-- hottest lines by stall samples (total {tot:.0f})")
        for d in sorted(data, reverse=True)[:nl]:
            print(f"  {d[0] / tot * 100:5.1f}%  inst={d[1]:12.0f}  {d[2][:110]}")
except Exception as e:
    print("source page parse failed:", e)
