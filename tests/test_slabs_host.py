"""Host-side slab logic on CPU: plane partition, ownership, and the neighbour exchange pattern under
torch.distributed with the gloo backend (world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest

from cellular_raza_b200.slabs import neighbor_exchange, rank_of_position, slab_planes


def test_slab_planes_partition():
    for n in range(1, 60):
        for w in range(1, min(n, 9) + 1):
            sl = slab_planes(n, w)
            assert sl[0][0] == 0 and sum(c for _, c in sl) == n
            assert all(sl[k][0] + sl[k][1] == sl[k + 1][0] for k in range(w - 1))
            counts = [c for _, c in sl]
            assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)
    # the C3 example of SURVEY.md 8d: 126 planes over 8 GPUs = 16 x 6 + 15 x 2
    assert [c for _, c in slab_planes(126, 8)] == [16] * 6 + [15] * 2
    with pytest.raises(ValueError):
        slab_planes(3, 4)


def test_rank_of_position():
    x = np.array([0.0, 17.99, 18.0, 53.9, 54.0, 107.9, 108.0, 1e9])
    r = rank_of_position(x, 0.0, 18.0, 6, 3)  # planes [0,1] [2,3] [4,5]
    assert r.tolist() == [0, 0, 0, 1, 1, 2, 2, 2]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for step in range(3):
            outs = [torch.full((16,), 100.0 * rank + 10 * step + side) for side in (0, 1)]
            ins = [torch.full((16,), -1.0) for _ in (0, 1)]
            neighbor_exchange(dist, rank, world, outs, ins)
            if rank > 0:  # my lower in-buffer holds the upper out-buffer of rank-1
                ok &= bool((ins[0] == 100.0 * (rank - 1) + 10 * step + 1).all())
            else:
                ok &= bool((ins[0] == -1.0).all())
            if rank + 1 < world:
                ok &= bool((ins[1] == 100.0 * (rank + 1) + 10 * step + 0).all())
            else:
                ok &= bool((ins[1] == -1.0).all())
        # conservation check used by the slab driver: sum of out counts == sum of in counts
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t)
        ok &= t.item() == world * (world + 1) / 2
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbor_exchange_gloo(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def test_rod_slab_workload_lies_in_its_own_planes():
    """bacterial_rods_slab (bench.py --gpus N --workload bacterial_rods): every rank generates the rods of its own
    slab of the N-times-longer domain, ids are globally unique, and the owner computed from the mean of the vertices
    (SortCells::get_voxel_index_of of CartesianCuboidRods) is the generating rank"""
    from cellular_raza_b200 import workloads
    from cellular_raza_b200.slabs import rank_of_position, slab_planes

    world = 4
    parts = [workloads.bacterial_rods_slab(2000, r, world) for r in range(world)]
    nv = parts[0].cfg_kwargs["n_voxels"]
    assert all((p.cfg_kwargs["n_voxels"] == nv).all() for p in parts)  # the GLOBAL domain on every rank
    ids = np.concatenate([p.extra["ids"] for p in parts])
    assert len(np.unique(ids)) == len(ids)
    planes = slab_planes(int(nv[0]), world)
    dx0 = parts[0].cfg_kwargs["dx"][0]
    for r, p in enumerate(parts):
        V = p.pos.shape[1]
        mean0 = sum(p.pos[:, v, 0] for v in range(V)) / float(V)
        owner = rank_of_position(mean0, 0.0, dx0, int(nv[0]), world)
        assert (owner == r).all()
        first, count = planes[r]
        assert mean0.min() >= first * dx0 and mean0.max() < (first + count) * dx0
    # one rank of one is the plain workload
    a, b = workloads.bacterial_rods_slab(500, 0, 1), workloads.bacterial_rods(10, 10, seed=[5, 0], spring_length_range=(4.0, 6.0))
    assert np.array_equal(a.pos, b.pos)
