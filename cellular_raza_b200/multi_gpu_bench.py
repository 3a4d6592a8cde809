"""bench.py --gpus N (N > 1): weak scaling of the mechanics/interaction step over slabs, one process
per GPU under torchrun (nccl over NVLink).  See bench.py for the meaning of the JSON keys."""
from __future__ import annotations

import json
import os
import time

import numpy as np


def run(args, rank: int, world: int, local_rank: int, helpers):
    import torch
    import torch.distributed as dist

    from . import workloads
    from .device import measure_fp64_peak
    from .slabs import SlabSimulation

    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # the halo transfer must get SM slots while the interior force pass saturates the device:
    # nccl kernels on a high-priority stream (the interior pass runs on the lowest priority)
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
    n = args.cells_per_gpu
    if args.workload == "brownian_morse":
        # configs[2]: ONE cube of n * world agents cut into slabs of whole planes
        w = workloads.brownian_morse_slab(n * world, rank, world, n_threads=args.n_threads_ref)
    else:
        w = workloads.cell_sorting_slab(n, rank, world, args.density, n_threads=args.n_threads_ref)
    n_rank = w.n
    # the fixed-size halo / migration buffers are sized from the capacity: it must be the same on every rank
    nmax = torch.tensor([n_rank], device="cuda", dtype=torch.int64)
    dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
    cfg = w.config(capacity=int(int(nmax.item()) * 1.1) + 4096, device=local_rank)
    transport = getattr(args, "transport", "p2p")
    sim = SlabSimulation(cfg, device=local_rank, overlap=not args.no_overlap, transport=transport)
    sim.set_uniform_params(w.mech_params[0], w.inter_params[0])
    sim.set_agents(w.pos, w.vel, None, None, w.species, w.extra["ids"])
    slab = sim.slab
    stream = slab.torch_stream

    # ---- protocol check on the hardware the numbers come from: the same warm-up steps through crb_slab_step
    # (peer-memory exchange, device-side counts) and through the six-phase protocol over ncclSend/ncclRecv must
    # leave every rank with the same bits; the global agent count is conserved ---------------------------------
    import hashlib

    def state_hash():
        pos, vel, ids, vox = slab.get_agents()
        o = np.argsort(ids, kind="stable")
        h = hashlib.blake2b(digest_size=16)
        for a in (ids[o], vox[o], pos[o], vel[o]):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.digest(), len(ids)

    n_check = max(args.warmup, 3)
    protocol_check = None
    if transport == "p2p":
        sim.step(n_check)
        sim.sync()
        h_p2p, n_p2p = state_hash()
        sim.set_agents(w.pos, w.vel, None, None, w.species, w.extra["ids"])
        slab.set_iteration(0)  # the Philox generator of configs[2] is keyed by (seed, id, iteration)
        sim.transport = "nccl"
        sim.step(n_check)
        sim.sync()
        h_nccl, n_nccl = state_hash()
        sim.transport = "p2p"
        same = torch.tensor([int(h_p2p == h_nccl), n_p2p, n_nccl], device="cuda", dtype=torch.int64)
        allr = [torch.zeros_like(same) for _ in range(world)]
        dist.all_gather(allr, same)
        protocol_check = {
            "p2p_equals_nccl_bitwise_on_every_rank": bool(all(int(t[0]) == 1 for t in allr)),
            "steps": n_check, "agents_total_p2p": int(sum(int(t[1]) for t in allr)),
            "agents_total_nccl": int(sum(int(t[2]) for t in allr)), "agents_expected": int(n * world),
        }
        if not protocol_check["p2p_equals_nccl_bitwise_on_every_rank"] or protocol_check["agents_total_p2p"] != n * world:
            raise RuntimeError(f"slab protocol check failed: {protocol_check}")
        sim.set_agents(w.pos, w.vel, None, None, w.species, w.extra["ids"])
        slab.set_iteration(0)

    sim.halo_exchange()
    pairs = slab.count_candidate_pairs()
    sim.step(max(args.warmup, 3))
    sim.sync()
    fp64_peak, _ = measure_fp64_peak(local_rank)

    sampler = helpers.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    slab.enable_timers(True)
    slab.reset_timers()
    l0 = slab.kernel_launches
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ms_local = e0.elapsed_time(e1)
    t = torch.tensor([ms_local], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    launches = slab.kernel_launches - l0
    force_ms, force_launches = slab.stage_time_ms(0)
    scan_ms, _ = slab.stage_time_ms(1)
    reorder_ms, _ = slab.stage_time_ms(2)
    halo_ms, _ = slab.stage_time_ms(3)
    mig_ms, _ = slab.stage_time_ms(4)
    slab.enable_timers(False)
    clocks = sampler.stop() if rank == 0 else None
    n_now = torch.tensor([float(slab.n_agents)], device="cuda", dtype=torch.float64)
    dist.all_reduce(n_now)
    sim.halo_exchange()
    pairs_after = slab.count_candidate_pairs()
    pr = torch.tensor([0.5 * (pairs + pairs_after)], device="cuda", dtype=torch.float64)
    fm = torch.tensor([force_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(pr)  # total pairs per step over all ranks
    dist.all_reduce(fm, op=dist.ReduceOp.MAX)
    value = n * world * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in / out every step (every rank its own slab) ---------------------------
    # per rank: crb_set_agents_async (pinned host -> device) + the slab step + crb_snapshot_async (device -> pinned
    # host); the download of step k overlaps the upload and the step of k + 1 (the snapshot reads its own packed
    # buffer).  The pinned buffers are allocated after the process has been bound to the CPUs next to its GPU
    # (NVML affinity: first-touch puts the pages on that NUMA node).
    e2e = None
    if not args.no_e2e:
        import ctypes as C

        numa = "not bound"
        try:
            import pynvml

            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            ncpu = os.cpu_count() or 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, (ncpu + 63) // 64)
            cpus = [c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa = f"bound to {len(cpus)} CPUs next to GPU {local_rank}"
        except Exception as e:  # affinity is an optimisation, never a failure
            numa = f"not bound ({type(e).__name__})"
        lib, h = slab._lib, slab._h
        ins = [[torch.from_numpy(w.pos.copy()).pin_memory(), torch.from_numpy(w.vel.copy()).pin_memory()] for _ in range(2)]
        cap_out = int(cfg.capacity)
        outs = [[torch.empty((cap_out, 3), dtype=torch.float64).pin_memory(),
                 torch.empty((cap_out, 3), dtype=torch.float64).pin_memory()] for _ in range(2)]
        hs = torch.from_numpy(w.species).pin_memory() if w.species is not None else None
        hs_ptr = hs.data_ptr() if hs is not None else None
        hi = torch.from_numpy(w.extra["ids"]).pin_memory()
        e2e_steps = max(args.steps // 2, 3)
        nn = C.c_uint64()
        pending = False

        def chk(rc):
            if rc:
                raise RuntimeError((lib.crb_last_error(h) or b"").decode())

        for k in range(3 + e2e_steps):
            j = k % 2
            if k == 3:
                if pending:
                    chk(lib.crb_snapshot_wait(h, C.byref(nn)))
                    pending = False
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
            chk(lib.crb_set_agents_async(h, n_rank, ins[j][0].data_ptr(), ins[j][1].data_ptr(), None, None, hs_ptr, hi.data_ptr()))
            sim.step(1)  # halo + migration with the neighbours; synchronises this rank at its end
            if pending:
                chk(lib.crb_snapshot_wait(h, C.byref(nn)))
            chk(lib.crb_snapshot_async(h, cap_out, outs[j][0].data_ptr(), outs[j][1].data_ptr(), None, None, None, None, None))
            pending = True
        if pending:
            chk(lib.crb_snapshot_wait(h, C.byref(nn)))
        torch.cuda.synchronize()
        dist.barrier()
        dt_e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt_e2e, op=dist.ReduceOp.MAX)
        dt_e2e = float(dt_e2e.item())
        # the host link alone, every rank copying at the same time (same bytes per rank and step, no kernels)
        floor_t = torch.tensor([helpers.pcie_copy_floor_ms(int(n_rank * (24 + 24 + 1 + 8)), int(n_rank * 48),
                                                           max(e2e_steps, 5), barrier=dist.barrier)],
                               device="cuda", dtype=torch.float64)
        dist.all_reduce(floor_t, op=dist.ReduceOp.MAX)
        e2e = {
            "value": n * world * e2e_steps / dt_e2e, "unit": helpers.UNIT,
            "pcie_copy_floor_ms_per_step": float(floor_t.item()),
            "h2d_bytes_per_step": int(world * n * (24 + 24 + 1 + 8)), "d2h_bytes_per_step": int(world * n * 48),
            "ms_per_step": dt_e2e / e2e_steps * 1e3, "host_memory": numa,
            "what": "per rank and step: crb_set_agents_async(pinned host pos, vel, species, ids) + slab step (halo + migration "
                    "with the neighbours) + crb_snapshot_async(pinned host pos, vel); the download of a step overlaps the "
                    "upload and the step of the next one",
        }

    if rank == 0:
        peaks = helpers.measured_peaks()
        pairs_per_launch = float(pr.item()) / world  # per GPU
        # the overlapped step splits the force pass into interior + boundary launches: time per STEP
        force_avg_ms = float(fm.item()) / max(args.steps, 1)
        achieved_tf = pairs_per_launch * helpers.flop_per_pair(args) / (force_avg_ms * 1e-3) / 1e12
        roofline = {
            "kernel": slab.force_kernel_name + " (per GPU, slowest rank)", "bound": "fp64", "achieved": achieved_tf,
            "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None,
            "peak_source": "DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
            "traffic": None, "share_of_step": float(fm.item()) / ms_total,
            "algorithmic_flop_per_pair": helpers.flop_per_pair(args), "pairs_per_launch": pairs_per_launch,
            "pair_forces_per_s": pairs_per_launch * world / (force_avg_ms * 1e-3),
        }
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        reorder_avg_ms = reorder_ms / max(args.steps, 1)
        reorder_gbs = n_rank * helpers.REORDER_BYTES_PER_AGENT / (reorder_avg_ms * 1e-3) / 1e9
        roofline_hbm = {
            "kernel": "k_fill+k_reorder (rank 0)", "bound": "hbm", "achieved": reorder_gbs, "peak": hbm_peak,
            "unit": "GB/s", "frac": reorder_gbs / hbm_peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
            "traffic": None, "share_of_step": reorder_ms / ms_total,
            "algorithmic_bytes_per_agent": helpers.REORDER_BYTES_PER_AGENT,
        }
        cfgd = helpers.workload_config(args)
        cfgd.update({"halo_bytes_per_face": slab.halo_bytes, "migration_bytes_per_face": slab.mig_bytes,
                     "exchange": ("crb_slab_step: export kernels store into the neighbour's arena over NVLink (CUDA IPC), "
                                  "device-side flags and counts, no host synchronisation inside the batch, no force return"
                                  if transport == "p2p" else
                                  "ncclSend/ncclRecv (torch.distributed.batch_isend_irecv), 2 phases per step, no force return"),
                     "halo_overlap": "off" if args.no_overlap else "interior planes on a second stream while the halo travels"})
        line = {
            "metric": helpers.METRIC, "value": value, "unit": helpers.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfgd,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": None, "e2e": e2e,
            "gpu_launches": int(launches) * world, "clocks": clocks,
            "stage_ms_per_step": {"force_integrate": force_ms / args.steps, "scan": scan_ms / args.steps,
                                  "reorder": reorder_ms / args.steps, "halo_pack_unpack": halo_ms / args.steps,
                                  "migration_pack_unpack": mig_ms / args.steps},
            "agents_after": float(n_now.item()),
            "protocol_check": protocol_check,
        }
        print(json.dumps(line), flush=True)
    sim.close()
    dist.destroy_process_group()


def run_rods(args, rank: int, world: int, local_rank: int, helpers):
    """BASELINE.json configs[3] over slabs: every rank owns ``cells-per-gpu`` rods of a domain ``world`` times longer
    along axis 0 (weak scaling); six-phase protocol over ncclSend/ncclRecv (halo = vertex positions of the boundary
    planes, migration = complete rods), divisions inside the timed region.  agent-steps = steps x mean of the global
    rod count before / after."""
    import torch
    import torch.distributed as dist

    from . import workloads
    from .device import measure_fp64_peak
    from .slabs import SlabSimulation

    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
    w = workloads.bacterial_rods_slab(args.cells_per_gpu, rank, world, n_threads=args.n_threads_ref)
    cfg = w.config(capacity=int(w.n * 1.5) + 4096, device=local_rank)
    sim = SlabSimulation(cfg, device=local_rank)
    sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, ids=w.extra["ids"])
    slab = sim.slab
    stream = slab.torch_stream

    def total(x):
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    sim.halo_exchange()
    pairs0 = slab.count_candidate_pairs()
    sim.step(max(args.warmup, 3))
    sim.sync()
    fp64_peak, _ = measure_fp64_peak(local_rank)
    sampler = helpers.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    slab.enable_timers(True)
    slab.reset_timers()
    l0 = slab.kernel_launches
    n_before = total(slab.n_agents)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    launches = slab.kernel_launches - l0
    force_ms, _ = slab.stage_time_ms(0)
    scan_ms, _ = slab.stage_time_ms(1)
    reorder_ms, _ = slab.stage_time_ms(2)
    halo_ms, _ = slab.stage_time_ms(3)
    mig_ms, _ = slab.stage_time_ms(4)
    slab.enable_timers(False)
    clocks = sampler.stop() if rank == 0 else None
    n_after = total(slab.n_agents)
    sim.halo_exchange()
    pairs = total(0.5 * (pairs0 + slab.count_candidate_pairs()))
    fm = torch.tensor([force_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(fm, op=dist.ReduceOp.MAX)
    value = 0.5 * (n_before + n_after) * args.steps / (ms_total * 1e-3)
    if rank == 0:
        force_avg_ms = float(fm.item()) / max(args.steps, 1)
        achieved_tf = pairs / world * helpers.ROD_FLOP_PER_PAIR / (force_avg_ms * 1e-3) / 1e12
        line = {
            "metric": helpers.METRIC, "value": value, "unit": helpers.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "bacterial_rods (BASELINE.json configs[3]) over slabs", "rods_per_gpu_initial": w.n,
                       "rods_before": n_before, "rods_after": n_after, "vertices": 5,
                       "aspects": "Mechanics, Interaction, Cycle, NeighborSensing", "dt_min": 0.1,
                       "decomposition": f"slab{world}", "halo_bytes_per_face": slab.halo_bytes,
                       "migration_bytes_per_face": slab.mig_bytes,
                       "exchange": "six-phase protocol, ncclSend/ncclRecv (torch.distributed.batch_isend_irecv); the rod "
                                   "force pass is not split into interior / boundary planes; two host synchronisations "
                                   "per step (division count, arrival counts)",
                       "l2_policy": "working set larger than L2 (three state copies x ~1 KB per rod)"},
            "roofline": {"kernel": "k_rod_force + k_rod_update (per GPU, slowest rank)", "bound": "fp64",
                         "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None,
                         "peak_source": "DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                         "traffic": None, "share_of_step": float(fm.item()) / ms_total,
                         "algorithmic_flop_per_pair": helpers.ROD_FLOP_PER_PAIR, "pairs_per_launch": pairs / world},
            "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches) * world, "clocks": clocks,
            "stage_ms_per_step": {"force_update": force_ms / args.steps, "scan": scan_ms / args.steps,
                                  "reorder": reorder_ms / args.steps, "halo_pack_unpack": halo_ms / args.steps,
                                  "divide_migration": mig_ms / args.steps},
        }
        print(json.dumps(line), flush=True)
    sim.close()
    dist.destroy_process_group()
