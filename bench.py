#!/usr/bin/env python
"""bench.py — agent-steps/s of the mechanics/interaction hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of chili on host cores

A "step" is one full pass of the hot path (pair forces -> AB3 update -> boundary -> cell-list
rebuild/migration) over the whole population.  Workload: 3-D cell_sorting (NewtonDamped3D + species
potential) at the density of the shipped example (rho ~ 7.1 agents per 18^3 voxel), `--cells-per-gpu`
agents per GPU (weak scaling, BASELINE.json configs[4] = 8 M -> 64 M at 1/2/4/8 GPUs).
Prints ONE JSON line (see the task contract / DESIGN.md §6 for the meaning of every key).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "agent_steps_per_s"
UNIT = "agent-steps/s"
FLOP_PER_PAIR = 50.0  # species potential counted as written, per ordered candidate pair
# k_fill + k_reorder per agent (uniform parameters): dynamic state read+write 2*(pos 24 + vel 24 +
# increments 48 + hlen 1) + id 16 + species 2 + fill (newkey 4, rank 4, tmp_idx 4) + reorder
# (tmp_idx 4, newkey 4, sortkey 8, key_out 4) = 244 B  
REORDER_BYTES_PER_AGENT = 2 * (24 + 24 + 48 + 1) + 16 + 2 + 12 + 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells-per-gpu", type=int, default=None,
                    help="agents per GPU (default 8 M; bacterial_rods 2 M = configs[3]; as-shipped 1,600)")
    ap.add_argument("--density", type=float, default=7.1, help="agents per voxel")
    ap.add_argument("--n-threads-ref", type=int, default=1,
                    help="chili n_threads whose pair weights are reproduced (CRB_WEIGHT_CHILI)")
    ap.add_argument("--cpu-sample-cells", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-seconds", type=float, default=12.0,
                    help="cpu_baseline: CPU work of the bounded sample (steps chosen from one probe step)")
    ap.add_argument("--workload", default="cell_sorting", choices=["cell_sorting", "brownian_morse", "bacterial_rods", "cell_sorting_as_shipped"],
                    help="cell_sorting = the headline (BASELINE.json configs[1]/[4]); brownian_morse = configs[2] "
                         "(N > 1: ONE cube of cells-per-gpu x N agents cut into slabs); bacterial_rods = configs[3] (N > 1: slabs, weak scaling); "
                         "cell_sorting_as_shipped = configs[0] (1,600 cells, latency-bound, 1 GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="N = 1 headline run: skip the short legs of BASELINE.json configs[0..3] (key extra_configs)")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: slab exchange through crb_slab_step (peer memory over NVLink, no host sync) or the "
                         "six-phase protocol over ncclSend/ncclRecv")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: halo exchange before the force pass instead of behind its interior planes")
    args = ap.parse_args()
    if args.cells_per_gpu is None:
        args.cells_per_gpu = {"bacterial_rods": 2_000_000, "cell_sorting_as_shipped": 1600}.get(args.workload, 8_000_000)
    return args


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.tmp = None

    def start(self):
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in Path(self.tmp.name).read_text().strip().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            pass
    return {}


def ncu_traffic(cells: int, kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel, from the committed ncu
    capture at this population (profiles/traffic.json; null when that size was not captured)."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(p.read_text())[str(cells)][kernel]
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps: int, warmup: int, sample_cells: int, target_s: float | None = None):
    """Times the CPU restatement of chili (oracle/) on the host cores: one emulated chili subdomain
    per core, `sample_cells` agents at the workload's density.  The Rust reference itself cannot be
    built in this image, so kind = "port".  With `target_s` the number of timed steps is chosen from
    one probe step so that the sample is about that many seconds of CPU work (bounded sample)."""
    from cellular_raza_b200 import workloads

    cores = host_cores()
    if args.workload == "bacterial_rods":
        from oracle.rods_oracle import RodsOracle

        side = max(int(round((sample_cells / 5) ** 0.5)), 1)
        w = workloads.bacterial_rods(side, side, rods_per_tile=5, n_threads=cores, spring_length_range=(4.0, 6.0))
        o = RodsOracle(w.config(), n_threads=cores)
        o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
        shape = f"{w.n} rods ({side}x{side} tiles, same density)"
    else:
        from oracle.chili_oracle import ChiliOracle

        if args.workload == "brownian_morse":
            w = workloads.brownian_morse_scaled(sample_cells, 4.0, n_threads=cores, noise_mode=0)
        elif args.workload == "cell_sorting_as_shipped":
            cores = min(cores, 4)  # main.rs:22 ships n_threads = 4
            w = workloads.cell_sorting(n_threads=cores)
        else:
            w = workloads.cell_sorting_scaled(sample_cells, args.density, n_threads=cores)
        o = ChiliOracle(w.config(), n_threads=cores)
        o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
        shape = f"{w.n} cells ({w.cfg_kwargs['n_voxels'][0]}^3 voxels, same density)"
    if warmup:
        o.step(warmup)
    if target_s is not None:
        t0 = time.perf_counter()
        o.step(1)
        probe = max(time.perf_counter() - t0, 1e-6)
        steps = int(min(max(round(target_s / probe), 2), 20000))
    n_before = w.n if not hasattr(o, "n_agents") else o.n_agents
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    n_after = w.n if not hasattr(o, "n_agents") else o.n_agents
    val = 0.5 * (n_before + n_after) * steps / dt
    sample = (f"{shape} x {steps} steps, {cores} threads = {cores} emulated chili subdomains, {dt:.1f} s")
    o.close()
    return val, cores, sample, dt / steps * 1e3


def reference_sample_cells(args) -> int:
    """cells the reference arm (--impl reference) steps: a bounded sample of the configured population at the same
    density (about 1 s per step on 16 host threads at 1 M cells), so that --steps K --warmup W ends within minutes"""
    total = args.steps + args.warmup
    sample_cells = args.cpu_sample_cells if total <= 30 else max(args.cpu_sample_cells // 4, 50_000)
    if args.workload == "bacterial_rods":
        sample_cells = min(sample_cells, 200_000)
    if args.workload == "cell_sorting_as_shipped":
        sample_cells = 1600
    return int(min(sample_cells, args.cells_per_gpu))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_cells = reference_sample_cells(args)
    val, cores, sample, ms = cpu_reference_run(args, args.steps, args.warmup, sample_cells)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # what this arm actually stepped: a bounded sample of the configured population at the same density
        # (agent-steps/s of this code is size-independent to a few per cent between 1 M and 8 M cells)
        "ran": {"cells": sample_cells, "steps_timed": args.steps, "warmup": args.warmup,
                "of_configured_cells_per_gpu": args.cells_per_gpu},
    }
    print(json.dumps(line), flush=True)


def flop_per_pair(args) -> float:
    # species potential 50 flop, Morse 30 flop (+ 1 exp, not counted) per ordered candidate pair
    return 30.0 if args.workload == "brownian_morse" else FLOP_PER_PAIR


def make_workload(args, n: int):
    from cellular_raza_b200 import _abi, workloads

    if args.workload == "brownian_morse":
        return workloads.brownian_morse_scaled(n, 4.0, n_threads=args.n_threads_ref, noise_mode=_abi.CRB_NOISE_PHILOX)
    if args.workload == "cell_sorting_as_shipped":
        return workloads.cell_sorting(n_threads=max(args.n_threads_ref, 4))  # main.rs:22 ships n_threads = 4
    return workloads.cell_sorting_scaled(n, args.density, n_threads=args.n_threads_ref)


def workload_config(args) -> dict:
    cfg = _workload_config(args)
    cfg["reference_arm_sample_cells"] = reference_sample_cells(args)
    return cfg


def _workload_config(args) -> dict:
    n_total = args.cells_per_gpu * args.gpus
    if args.workload == "brownian_morse":
        return {
            "workload": "brownian3d_morse (BASELINE.json configs[2])", "cells_per_gpu": args.cells_per_gpu,
            "cells_total": n_total, "agents_per_voxel": 4.0, "mechanics": "Brownian3D{D 0.3, kT 0.3}",
            "interaction": "MorsePotential{radius 3, stiffness 0.5, cutoff 12, strength 0.01}",
            "solver": "AB3 (order 2)", "noise": "counter-based Philox normal per (seed, id, step)",
            "pair_weights": f"chili-as-written(n_threads={args.n_threads_ref})", "dt": 0.01,
            "l2_policy": "working set larger than L2", "decomposition": f"slab{args.gpus}" if args.gpus > 1 else "single",
        }
    if args.workload == "cell_sorting_as_shipped":
        return {
            "workload": "cell_sorting as shipped (BASELINE.json configs[0]: 1,600 cells, 6^3 voxels; latency-bound)",
            "cells_per_gpu": 1600, "cells_total": 1600, "agents_per_voxel": 7.4, "mechanics": "NewtonDamped3D",
            "interaction": "cell_sorting species potential", "solver": "AB3 (order 2)",
            "pair_weights": f"chili-as-written(n_threads={max(args.n_threads_ref, 4)})", "dt": 0.2,
            "l2_policy": "working set fits L2 by construction (1,600 agents): launch-latency bound, reported not tuned",
            "stepping": "crb_step(K) replays CUDA graphs of two steps each; stage times from a second, event-instrumented pass",
            "decomposition": "single",
        }
    if args.workload == "bacterial_rods":
        return {"workload": "bacterial_rods (BASELINE.json configs[3])", "rods_initial": args.cells_per_gpu, "vertices": 5,
                "aspects": "Mechanics, Interaction, Cycle, NeighborSensing", "dt_min": 0.1}
    return {
        "workload": "cell_sorting_3d_weak_scaling (BASELINE.json configs[4]; configs[1] with --cells-per-gpu 1000000)",
        "cells_per_gpu": args.cells_per_gpu, "cells_total": n_total, "agents_per_voxel": args.density,
        "mechanics": "NewtonDamped3D", "interaction": "cell_sorting species potential", "solver": "AB3 (order 2)",
        "pair_weights": f"chili-as-written(n_threads={args.n_threads_ref})", "dt": 0.2,
        "l2_policy": "working set larger than L2 (3 state copies x >= 100 B/agent x cells_per_gpu)",
        "decomposition": f"slab{args.gpus}" if args.gpus > 1 else "single",
    }


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus != 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with WORLD_SIZE={args.gpus} (got {world})")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if args.gpus > 1:
        if args.workload == "cell_sorting_as_shipped":
            raise SystemExit("--workload cell_sorting_as_shipped is a single-GPU line (1,600 cells do not fill one GPU): "
                             "run it with --gpus 1")
        from cellular_raza_b200 import multi_gpu_bench

        if args.workload == "bacterial_rods":
            return multi_gpu_bench.run_rods(args, rank, world, local_rank, sys.modules[__name__])

        return multi_gpu_bench.run(args, rank, world, local_rank, sys.modules[__name__])

    dev_id = local_rank
    torch.cuda.set_device(dev_id)
    line = measure_point(args, dev_id)
    if args.workload == "cell_sorting" and not args.no_extra_configs:
        # the other BASELINE.json configurations as short legs of the same run, so that the driver's record holds
        # a measured number for every one of them (value, roofline, e2e each; no CPU baseline; < 60 s together)
        line["extra_configs"] = extra_config_legs(args, dev_id)
    print(json.dumps(line), flush=True)


EXTRA_LEGS = (
    # key, workload, cells, steps, warmup, e2e steps
    ("configs[0] cell_sorting as shipped (1,600 cells)", "cell_sorting_as_shipped", 1600, 2000, 50, 20),
    ("configs[1] cell_sorting 1 M cells", "cell_sorting", 1_000_000, 20, 5, 12),
    ("configs[2] Brownian3D + Morse 8 M cells (1 GPU)", "brownian_morse", 8_000_000, 8, 3, 8),
    ("configs[3] bacterial_rods 2 M rods", "bacterial_rods", 2_000_000, 8, 3, 8),
)


def extra_config_legs(args, dev_id: int) -> dict:
    import copy

    out = {}
    for key, workload, cells, steps, warmup, e2e_steps in EXTRA_LEGS:
        a = copy.copy(args)
        a.workload, a.cells_per_gpu, a.steps, a.warmup = workload, cells, steps, warmup
        a.density = 7.1
        a.no_cpu_baseline = True
        a.e2e_steps = e2e_steps
        t0 = time.perf_counter()
        try:
            leg = measure_point(a, dev_id)
            keep = ("value", "unit", "ms_per_step", "steps", "warmup", "config", "roofline", "roofline_hbm", "e2e",
                    "gpu_launches", "stage_ms_per_step")
            out[key] = {k: leg[k] for k in keep if k in leg}
        except Exception as e:  # a failing leg must not take the headline down
            out[key] = {"error": f"{type(e).__name__}: {e}"}
        out[key]["leg_wall_s"] = round(time.perf_counter() - t0, 2)
    return out


def pcie_copy_floor_ms(h2d_bytes: int, d2h_bytes: int, steps: int, barrier=None) -> float:
    """What the host link alone allows for the e2e loop: per step ``h2d_bytes`` pinned host -> device and ``d2h_bytes``
    device -> pinned host on two streams (full duplex), nothing else, averaged over ``steps``; with ``barrier`` every
    rank of the node copies at the same time (the ranks share the host's memory and PCIe root complexes).  Reported
    next to e2e so that the plumbing and the kernels can be told apart."""
    import torch

    hin = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8).pin_memory()
    hout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()
    din = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device="cuda")
    dout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def one():
        with torch.cuda.stream(s_in):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s_out):
            hout.copy_(dout, non_blocking=True)

    one()
    torch.cuda.synchronize()
    if barrier is not None:
        barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


def measure_point(args, dev_id: int) -> dict:
    """one workload on one GPU -> the JSON line as a dict"""
    import torch

    from cellular_raza_b200 import workloads
    from cellular_raza_b200.device import DeviceSimulation, measure_fp64_peak

    n = args.cells_per_gpu
    if args.workload == "bacterial_rods":
        return run_rods(args, dev_id)
    w = make_workload(args, n)
    n = w.n
    sim = DeviceSimulation(w.config(device=dev_id))
    sim.set_uniform_params(w.mech_params[0], w.inter_params[0])
    sim.set_agents(w.pos, w.vel, None, None, w.species)
    stream = torch.cuda.ExternalStream(sim.stream, device=dev_id)
    pairs = sim.count_candidate_pairs()

    sim.step(max(args.warmup, 3))
    fp64_peak, _ = measure_fp64_peak(dev_id)

    # ---- timed region: K steps, device-resident ---------------------------------------------------
    sampler = ClockSampler(dev_id)
    sampler.start()
    time.sleep(0.3)
    # the per-stage CUDA events live inside the timed region, except for the launch-bound as-shipped case:
    # six extra event records per 60-us step would be a large part of what is measured there (and the library
    # replays CUDA graphs only while its timers are off), so its stage times come from a second pass
    split_timers = args.workload == "cell_sorting_as_shipped"
    sim.enable_timers(not split_timers)
    sim.reset_timers()
    l0 = sim.kernel_launches
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = sim.kernel_launches - l0
    ms_stage_total = ms_total
    if split_timers:
        sim.enable_timers(True)
        sim.reset_timers()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        sim.step(args.steps)
        e3.record(stream)
        torch.cuda.synchronize()
        ms_stage_total = e2.elapsed_time(e3)
    force_ms, force_launches = sim.stage_time_ms(0)
    scan_ms, _ = sim.stage_time_ms(1)
    reorder_ms, reorder_launches = sim.stage_time_ms(2)
    sim.enable_timers(False)
    clocks = sampler.stop()
    pairs_after = sim.count_candidate_pairs()
    value = n * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (force + integrate): FP64 pipe --------------------------
    pairs_per_launch = 0.5 * (pairs + pairs_after)
    force_avg_ms = force_ms / max(force_launches, 1)
    achieved_tf = pairs_per_launch * flop_per_pair(args) / (force_avg_ms * 1e-3) / 1e12
    peaks = measured_peaks()
    roofline = {
        "kernel": sim.force_kernel_name, "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None,
        "peak_source": "DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
        "traffic": ncu_traffic(n, "force"), "share_of_step": force_ms / ms_stage_total,
        "algorithmic_flop_per_pair": flop_per_pair(args), "pairs_per_launch": pairs_per_launch,
        "pair_forces_per_s": pairs_per_launch / (force_avg_ms * 1e-3),
        # `achieved` is delivered reference-equivalent work (SURVEY 8d counts the potential for EVERY candidate of the
        # 27 voxels; the FP32 filter skips the ~85 % beyond the cutoff, whose contribution is exactly zero), not
        # executed DFMAs: what limits the kernel is instruction issue
        "limiter": "instruction issue (ncu, 8 M cells: issue-active 57.5 %, executed FP64 pipe 20 %, "
                   "profiles/r2_ncu_k_force_tile2_v12_8M.txt)",
    }
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    reorder_avg_ms = reorder_ms / max(args.steps, 1)
    reorder_gbs = n * REORDER_BYTES_PER_AGENT / (reorder_avg_ms * 1e-3) / 1e9
    roofline_hbm = {
        "kernel": "k_fill+k_reorder (cell-list rebuild / migration scatter)", "bound": "hbm",
        "achieved": reorder_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": reorder_gbs / hbm_peak,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
        "traffic": ncu_traffic(n, "reorder"), "share_of_step": reorder_ms / ms_stage_total, "scan_share_of_step": scan_ms / ms_stage_total,
        "algorithmic_bytes_per_agent": REORDER_BYTES_PER_AGENT,
    }

    # ---- e2e: host buffers in, host buffers out, every step --------------------------------------
    # Two handles alternate so that the upload of population k + 1 (crb_set_agents_async), the step of population k
    # (crb_step_async) and the download of population k - 1 (crb_snapshot_async) overlap: PCIe is full duplex and
    # the copies of one handle run beside the kernels of the other.  Every iteration still moves its complete input
    # from pinned host memory to the device and its complete result back, inside the timed region.
    e2e = None
    if not args.no_e2e:
        import ctypes as C

        lib = sim._lib
        sim_b = DeviceSimulation(w.config(device=dev_id))
        sim_b.set_uniform_params(w.mech_params[0], w.inter_params[0])
        handles = [sim._h, sim_b._h]
        # two buffer sets per handle: the upload of a handle's next population is queued BEFORE the download of its
        # previous one is awaited (the download reads the packed snapshot buffer, not the state), so the result
        # that feeds back as an input is the one of four iterations ago
        ins = [[None, None], [None, None]]
        outs = [[None, None], [None, None]]
        for k in range(2):
            for j in range(2):
                ins[k][j] = [torch.from_numpy(w.pos.copy()).pin_memory(), torch.from_numpy(w.vel.copy()).pin_memory()]
                outs[k][j] = [torch.empty(w.pos.shape, dtype=torch.float64).pin_memory(),
                              torch.empty(w.vel.shape, dtype=torch.float64).pin_memory()]
        hs = torch.from_numpy(w.species).pin_memory() if w.species is not None else None
        hs_ptr = hs.data_ptr() if hs is not None else None
        e2e_steps = getattr(args, "e2e_steps", None) or args.steps
        pending = [None, None]  # buffer set of the snapshot in flight per handle
        nn = C.c_uint64()

        def chk(rc, h):
            if rc:
                raise RuntimeError((lib.crb_last_error(h) or b"").decode())

        def retire(k):
            h, j = handles[k], pending[k]
            chk(lib.crb_snapshot_wait(h, C.byref(nn)), h)
            pending[k] = None
            ins[k][j], outs[k][j] = outs[k][j], ins[k][j]  # this result is a later input of the handle

        t0 = None
        warm = 8
        for it in range(warm + e2e_steps):
            k, j = it % 2, (it // 2) % 2
            if it == warm:
                for q in range(2):
                    if pending[q] is not None:
                        retire(q)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            h = handles[k]
            chk(lib.crb_set_agents_async(h, n, ins[k][j][0].data_ptr(), ins[k][j][1].data_ptr(), None, None, hs_ptr, None), h)
            chk(lib.crb_step_async(h, 1), h)
            if pending[k] is not None:
                retire(k)  # the download of this handle's previous population (started two iterations ago)
            chk(lib.crb_snapshot_async(h, n, outs[k][j][0].data_ptr(), outs[k][j][1].data_ptr(), None, None, None, None, None), h)
            pending[k] = j
        for q in range(2):
            if pending[q] is not None:
                retire(q)
        for h in handles:
            chk(lib.crb_sync(h), h)  # latched device errors of the asynchronous calls surface here
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        sim_b.close()
        # bytes that cross the host link per step: positions always; velocities unless the mechanics has none
        # (Brownian3D: the library does not read them); one species byte where the potential uses it
        h2d = int(n * (24 + (0 if args.workload == "brownian_morse" else 24) + (1 if hs is not None else 0)))
        floor_ms = pcie_copy_floor_ms(h2d, int(n * 48), max(e2e_steps, 5))
        e2e = {
            "value": n * e2e_steps / dt_e2e, "unit": UNIT,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(n * 48),
            "ms_per_step": dt_e2e / e2e_steps * 1e3,
            "pcie_copy_floor_ms_per_step": floor_ms,  # the same bytes both ways, no kernels (measured in this run)
            "what": "per step: crb_set_agents_async(pinned host pos, vel, species) + crb_step_async(1) + "
                    "crb_snapshot_async(pinned host pos, vel) + crb_snapshot_wait; two handles alternate so that "
                    "upload, step and download of consecutive populations overlap",
        }

    # ---- CPU baseline (oracle, bounded sample) ----------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_run(args, 1, 0, min(args.cpu_sample_cells, n), target_s=args.cpu_sample_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args), "roofline": roofline, "roofline_hbm": roofline_hbm,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "stage_ms_per_step": {"force_integrate": force_ms / args.steps, "scan": scan_ms / args.steps,
                              "reorder": reorder_ms / args.steps},
        "published_reference": "chili on Threadripper 3960X: 2.02e4 (1 thread, 1.3 M cells), best 1.01e5 (43 threads, 20 k cells); BASELINE.md",
    }
    sim.close()
    return line


# per ordered candidate rod pair, 20 vertex-segment evaluations x (Morse 30 flop + nearest
# point ~25 flop) + 25 neighbour-sensing distance checks x 9 flop (the exp is not counted)
ROD_FLOP_PER_PAIR = 20 * (30.0 + 25.0) + 25 * 9.0


def run_rods(args, dev_id: int):
    """BASELINE.json configs[3]: bacterial_rods (RodMechanics + RodInteraction, growth / division => the rod
    count changes inside the timed region; agent-steps = sum over steps of the current count, here
    steps x mean(count before, count after))."""
    import torch

    from cellular_raza_b200 import workloads
    from cellular_raza_b200.device import measure_fp64_peak
    from cellular_raza_b200.rods import RodSimulation

    n = args.cells_per_gpu
    per_tile = 5
    side = max(int(round((n / per_tile) ** 0.5)), 1)
    # spring lengths up to the division threshold (6 um): ~10 % of the rods divide inside the timed region
    w = workloads.bacterial_rods(side, side, rods_per_tile=per_tile, n_threads=args.n_threads_ref,
                                 spring_length_range=(4.0, 6.0))
    sim = RodSimulation(w.config(device=dev_id))
    sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    stream = torch.cuda.ExternalStream(sim.stream, device=dev_id)
    sim.step(max(args.warmup, 3))
    fp64_peak, _ = measure_fp64_peak(dev_id)
    pairs0 = sim.count_candidate_pairs()
    sampler = ClockSampler(dev_id)
    sampler.start()
    time.sleep(0.3)
    sim.enable_timers(True)
    sim.reset_timers()
    l0, n0 = sim.kernel_launches, sim.n_agents
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    n1 = sim.n_agents
    launches = int(sim.kernel_launches - l0)
    force_ms, force_launches = sim.stage_time_ms(0)
    scan_ms, _ = sim.stage_time_ms(1)
    reorder_ms, _ = sim.stage_time_ms(2)
    divide_ms, _ = sim.stage_time_ms(4)
    sim.enable_timers(False)
    clocks = sampler.stop()
    pairs1 = sim.count_candidate_pairs()
    value = 0.5 * (n0 + n1) * args.steps / (ms_total * 1e-3)
    pairs_per_launch = 0.5 * (pairs0 + pairs1)
    # stage 0 of a rod step = k_rod_force (pair forces + neighbour counts -> force_out) followed by k_rod_update
    # (per-rod spring/curvature increment, Adams-Bashforth, growth): two launches, timed together per step
    force_avg_ms = force_ms / max(args.steps, 1)
    achieved_tf = pairs_per_launch * ROD_FLOP_PER_PAIR / (force_avg_ms * 1e-3) / 1e12
    roofline = {
        "kernel": "k_rod_force<5 vertices> (one thread per rod) + k_rod_update", "bound": "fp64", "achieved": achieved_tf,
        "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None,
        "peak_source": "DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
        "traffic": ncu_traffic(n, "rod_force"), "share_of_step": force_ms / ms_total,
        "algorithmic_flop_per_pair": ROD_FLOP_PER_PAIR, "pairs_per_launch": pairs_per_launch,
        "pair_forces_per_s": pairs_per_launch / (force_avg_ms * 1e-3),
        "note": "the shipped rod density is 0.31 rods per voxel (~2 candidate rods per rod): the kernel is "
                "dominated by the per-rod spring/curvature mechanics and by divergence, not by pair flops",
    }

    # ---- e2e: host buffers in, host buffers out, every step (the rod count is pinned by re-uploading the
    # same n rods every step).  Two handles alternate; a rod step always synchronises (the division count is read
    # back), so the upload of population k + 1 is queued on the OTHER handle before the step of population k is
    # called: the host link carries upload k + 1 and download k - 1 while the device steps population k. ----------
    e2e = None
    if not args.no_e2e:
        import ctypes

        nn = w.n
        V = 5
        hp = torch.from_numpy(np.ascontiguousarray(w.pos)).pin_memory()
        hv = torch.from_numpy(np.ascontiguousarray(w.vel)).pin_memory()
        hm = torch.from_numpy(np.ascontiguousarray(w.mech_params)).pin_memory()
        hi = torch.from_numpy(np.ascontiguousarray(w.inter_params)).pin_memory()
        cap_out = int(nn * 1.2) + 16
        lib = sim._lib
        sim2 = RodSimulation(w.config(device=dev_id))
        hs = [sim._h, sim2._h]
        outs = [(torch.empty((cap_out, V, 3), dtype=torch.float64).pin_memory(),
                 torch.empty((cap_out, V, 3), dtype=torch.float64).pin_memory()) for _ in range(2)]
        e2e_steps = getattr(args, "e2e_steps", None) or args.steps
        warm = 2
        total = warm + e2e_steps

        def chk(rc, h):
            if rc:
                raise RuntimeError((lib.crb_last_error(h) or b"").decode())

        def upload(k):
            h = hs[k % 2]
            chk(lib.crb_set_agents_async(h, nn, hp.data_ptr(), hv.data_ptr(), hm.data_ptr(), hi.data_ptr(), None, None), h)

        upload(0)
        pending = [False, False]
        n_out = ctypes.c_uint64()
        for k in range(total):
            if k == warm:
                for j in range(2):
                    if pending[j]:
                        chk(lib.crb_snapshot_wait(hs[j], ctypes.byref(n_out)), hs[j])
                        pending[j] = False
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            j = k % 2
            h = hs[j]
            if k + 1 < total:
                upload(k + 1)  # (its previous download reads the packed snapshot buffer, not the state)
            chk(lib.crb_step(h, 1), h)
            if pending[1 - j]:
                chk(lib.crb_snapshot_wait(hs[1 - j], ctypes.byref(n_out)), hs[1 - j])
                pending[1 - j] = False
            chk(lib.crb_snapshot_async(h, cap_out, outs[j][0].data_ptr(), outs[j][1].data_ptr(), None, None, None, None, None), h)
            pending[j] = True
        for j in range(2):
            if pending[j]:
                chk(lib.crb_snapshot_wait(hs[j], ctypes.byref(n_out)), hs[j])
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        sim2.close()
        h2d = int(nn * (2 * V * 24 + 8 * 8 + 4 * 8))
        d2h = int(int(n_out.value) * 2 * V * 24)
        floor_ms = pcie_copy_floor_ms(h2d, d2h, max(e2e_steps, 5))
        e2e = {
            "value": nn * e2e_steps / dt_e2e, "unit": UNIT,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": dt_e2e / e2e_steps * 1e3,
            "pcie_copy_floor_ms_per_step": floor_ms,
            "what": "per step: crb_set_agents_async(pinned host pos, vel, mech, inter rows; rows checked on the device) + "
                    "crb_step(1) + crb_snapshot_async(pinned host pos, vel) + crb_snapshot_wait; two handles alternate, "
                    "the next upload is queued before the (synchronising) rod step",
        }

    cpu = None
    if not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_run(args, 1, 0, min(args.cpu_sample_cells, 200_000, n),
                                                target_s=args.cpu_sample_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "bacterial_rods (BASELINE.json configs[3])", "rods_initial": w.n, "rods_before": n0,
                   "rods_after": n1, "vertices": 5, "tiles": [side, side],
                   "aspects": "Mechanics, Interaction, Cycle, NeighborSensing", "dt_min": 0.1,
                   "l2_policy": "working set larger than L2" if w.n >= 500_000 else "working set fits L2 (small run)"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks,
        "stage_ms_per_step": {"force_integrate": force_ms / args.steps, "divide": divide_ms / args.steps,
                              "scan": scan_ms / args.steps, "reorder": reorder_ms / args.steps},
        "note": "not the headline workload: k_rod_force is a per-rod kernel bound by the FP64 pipe (69 % under ncu; DESIGN.md §9)",
    }
    sim.close()
    return line


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
