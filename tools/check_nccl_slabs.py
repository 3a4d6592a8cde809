#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_nccl_slabs.py [nccl|p2p] : slabs with one process per GPU — the
six-phase protocol over ncclSend/ncclRecv, or crb_slab_step with the neighbours' arenas mapped through CUDA IPC —
must reproduce the single-handle run on rank 0 bit for bit.  Prints 'NCCL_SLABS_OK' on success."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import torch.distributed as dist

    from cellular_raza_b200 import workloads
    from cellular_raza_b200.device import DeviceSimulation
    from cellular_raza_b200.slabs import SlabSimulation, split_agents_by_slab

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = 24
    transport = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    w = workloads.cell_sorting_scaled(120_000, 7.1, n_threads=3)  # identical on every rank (seeded)
    cfg = w.config(device=local)
    owner = split_agents_by_slab(cfg, world, w.pos)
    m = owner == rank
    ids = np.arange(w.n, dtype=np.uint64)
    sim = SlabSimulation(cfg, device=local, transport=transport)
    sim.set_agents(w.pos[m], w.vel[m], w.mech_params[m], w.inter_params[m], w.species[m], ids[m])
    crossed = 0
    for _ in range(3):
        sim.step(steps // 3)
        crossed += sum(sim.slab.last_migration)
    sim.sync()
    pos, vel, pid, vox = sim.get_agents()
    n_loc = torch.tensor([len(pid)], device="cuda")
    sizes = [torch.zeros_like(n_loc) for _ in range(world)]
    dist.all_gather(sizes, n_loc)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    pack = np.zeros((nmax, 8))
    pack[: len(pid), 0:3] = pos
    pack[: len(pid), 3:6] = vel
    pack[: len(pid), 6] = pid.astype(np.float64)
    pack[: len(pid), 7] = vox.astype(np.float64)
    t = torch.from_numpy(pack).cuda()
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    cr = torch.tensor([float(crossed)], device="cuda")
    dist.all_reduce(cr)
    ok = True
    if rank == 0:
        allp = np.concatenate([o.cpu().numpy()[:s] for o, s in zip(out, sizes)])
        allp = allp[np.argsort(allp[:, 6])]
        one = DeviceSimulation(w.config(device=local))
        one.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species, ids)
        one.step(steps)
        p1, v1, i1, x1 = one.get_agents()
        o = np.argsort(i1)
        ok = (len(allp) == w.n and np.array_equal(allp[:, 0:3], p1[o]) and np.array_equal(allp[:, 3:6], v1[o])
              and np.array_equal(allp[:, 7].astype(np.uint64), x1[o]))
        print(f"world {world} transport {transport}: agents {len(allp)}, crossed slabs (sampled steps) {int(cr.item())}, bitwise equal: {ok}")
        if ok:
            print("NCCL_SLABS_OK")
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def main_rods():
    """the same for bacterial rods (six-phase protocol over ncclSend/ncclRecv): positions, identifiers incl.
    Division(voxel, counter) and voxel assignment of the gathered slabs == single handle on rank 0"""
    import torch
    import torch.distributed as dist

    from cellular_raza_b200 import workloads
    from cellular_raza_b200.rods import RodSimulation
    from cellular_raza_b200.slabs import SlabSimulation, rank_of_position

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = 90
    um = 1e-6
    w = workloads.bacterial_rods(4, 3, rods_per_tile=8, spring_length_range=(5.5, 5.99), capacity_factor=4.0)
    pos = w.pos.copy()
    pos[:, :, 0] += 175 * um  # the rod clusters straddle the planes x = 200, 400, 600 um
    Lx = w.cfg_kwargs["domain_max"][0]
    keep = (pos[:, :, 0].min(axis=1) > um) & (pos[:, :, 0].max(axis=1) < Lx - um)
    pos, mech, inter = pos[keep], w.mech_params[keep], w.inter_params[keep]
    n = len(pos)
    V = pos.shape[1]
    cfg = w.config(device=local)
    mean0 = sum(pos[:, v, 0] for v in range(V)) / float(V)
    owner = rank_of_position(mean0, cfg.domain_min[0], cfg.dx[0], int(cfg.n_voxels[0]), world)
    m = owner == rank
    ids = np.arange(n, dtype=np.uint64)
    sim = SlabSimulation(cfg, device=local)
    sim.set_agents(pos[m], None, mech[m], inter[m], ids=ids[m])
    crossed = 0
    for _ in range(steps):
        sim.step(1)
        crossed += sum(sim.slab.last_migration)
    sim.sync()
    p, v, ident, parent, vox = sim.get_agents()
    k = len(ident)
    n_loc = torch.tensor([k], device="cuda")
    sizes = [torch.zeros_like(n_loc) for _ in range(world)]
    dist.all_gather(sizes, n_loc)
    sizes = [int(x.item()) for x in sizes]
    nmax = max(sizes)
    W = 3 * V
    pack = np.zeros((nmax, 2 * W + 7), dtype=np.float64)
    pack[:k, :W] = p.reshape(k, W)
    pack[:k, W:2 * W] = v.reshape(k, W)
    ints = np.zeros((nmax, 7), dtype=np.uint64)
    ints[:k, 0:3] = ident
    ints[:k, 3:6] = parent
    ints[:k, 6] = vox
    pack[:, 2 * W:] = ints.view(np.float64)  # bit pattern, not value
    t = torch.from_numpy(pack).cuda()
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    cr = torch.tensor([float(crossed)], device="cuda")
    dist.all_reduce(cr)
    ok = True
    if rank == 0:
        allp = np.concatenate([o.cpu().numpy()[:s] for o, s in zip(out, sizes)])  # rank order = device order
        one = RodSimulation(w.config(device=local))
        one.set_agents(pos, None, mech, inter, ids=ids)
        one.step(steps)
        p1, v1, i1, pa1, x1 = one.get_agents()
        gi = np.ascontiguousarray(allp[:, 2 * W:]).view(np.uint64)
        ok = (len(allp) == len(i1) and len(i1) > n and np.array_equal(allp[:, :W], p1.reshape(len(i1), W))
              and np.array_equal(allp[:, W:2 * W], v1.reshape(len(i1), W)) and np.array_equal(gi[:, 0:3], i1)
              and np.array_equal(gi[:, 3:6], pa1) and np.array_equal(gi[:, 6], x1))
        print(f"rods: world {world}: {n} -> {len(allp)} rods, crossed slabs {int(cr.item())}, bitwise equal: {ok}")
        if ok:
            print("NCCL_SLABS_OK")
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rods":
        main_rods()
    else:
        main()
