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


if __name__ == "__main__":
    main()
