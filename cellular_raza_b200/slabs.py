"""Slab decomposition over several GPUs: host-side orchestration of the
``crb_slab_*`` entry points of ``include/crb.h``.

chili cuts the voxel list into ``n_threads`` contiguous chunks, one ``SubDomainBox`` per OS thread,
and moves ``PosInformation`` / ``ForceInformation`` / ``SendCell`` messages through crossbeam channels
separated by a global barrier (cellular_raza-core/src/backend/chili/simulation_flow.rs:350-423,
438-450, 689-707; update_mechanics.rs:153-238, 418-513).  Here the domain is cut into slabs of whole
voxel planes along axis 0, one ``crb_sim`` handle per GPU; per step every rank sends ONE halo buffer
(boundary plane: positions + interaction information) and ONE migration buffer (complete agents that
crossed the face) to each of its at most two neighbours.  There is no force return message: each side
evaluates its own half of a cross-slab pair, exactly as chili evaluates every pair from both sides.

Two transports move the fixed-size device buffers:

* :class:`LocalSlabGroup` — all handles live in this process (same or different devices); buffers are
  moved with device copies.  Used by the tests (several slabs on ONE GPU reproduce the single-handle
  run bit for bit) and usable for one-process multi-GPU runs.
* :class:`SlabSimulation` — one process per GPU (``torch.distributed``, backend ``nccl``): buffers
  travel with ``ncclSend/ncclRecv`` over NVLink (``torch.distributed.batch_isend_irecv``).

torch is used for device buffers, streams and the NCCL plumbing only.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _abi
from .device import DeviceSimulation


def slab_planes(n_planes: int, world: int) -> list[tuple[int, int]]:
    """(first global plane, number of planes) of every rank: ranks below ``n_planes % world`` own one
    plane more.  Same rule as crb_create (csrc/crb_api.cu) and as chili's chunking of the voxel list
    into sizes ceil/floor (cartesian_cuboid_n.rs:19-59)."""
    if world < 1 or world > n_planes:
        raise ValueError("need 1 <= world <= n_planes")
    base, rem = divmod(n_planes, world)
    out = []
    for r in range(world):
        out.append((r * base + min(r, rem), base + (1 if r < rem else 0)))
    return out


def rank_of_position(x0: np.ndarray, domain_min0: float, dx0: float, n_planes: int, world: int) -> np.ndarray:
    """slab that owns agents with axis-0 coordinate ``x0`` (voxel index as get_voxel_index_of_raw,
    cartesian_cuboid_n.rs:320-337)."""
    plane = np.floor((np.asarray(x0, dtype=np.float64) - domain_min0) / dx0).astype(np.int64)
    plane = np.clip(plane, 0, n_planes - 1)
    firsts = np.array([f for f, _ in slab_planes(n_planes, world)], dtype=np.int64)
    return np.searchsorted(firsts, plane, side="right") - 1


def neighbor_exchange(dist, rank: int, world: int, outs, ins, group=None):
    """Posts the at most four point-to-point operations of one exchange phase: ``outs[0]`` goes to
    rank-1 and ``ins[0]`` comes from it, ``outs[1]`` / ``ins[1]`` likewise for rank+1.  Works with any
    ``torch.distributed`` backend (nccl on device buffers; gloo on CPU tensors in the tests).
    Returns the work handles (already waited: for nccl the wait is stream-ordered)."""
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, outs[0], rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, ins[0], rank - 1, group))
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, outs[1], rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, ins[1], rank + 1, group))
    works = dist.batch_isend_irecv(ops) if ops else []
    for w in works:
        w.wait()
    return works


class SlabRank(DeviceSimulation):
    """One slab: a ``crb_sim`` handle created with (rank, world_size) plus its exchange buffers."""

    def __init__(self, cfg: _abi.CrbConfig, rank: int, world: int, device: int = 0,
                 ghost_capacity: int = 0, migration_capacity: int = 0):
        import torch

        cfg.rank = rank
        cfg.world_size = world
        cfg.device = device
        cfg.ghost_capacity = ghost_capacity
        cfg.migration_capacity = migration_capacity
        super().__init__(cfg)
        self.rank, self.world, self.device = rank, world, device
        info = (C.c_uint64 * 8)()
        self._check(self._lib.crb_slab_info(self._h, info))
        self.first_plane, self.n_planes = int(info[0]), int(info[1])
        self.has_lo, self.has_hi = bool(info[2]), bool(info[3])
        self.halo_bytes = int(self._lib.crb_slab_halo_bytes(self._h))
        self.mig_bytes = int(self._lib.crb_slab_migration_bytes(self._h))
        dev = torch.device("cuda", device)
        self.torch_stream = torch.cuda.ExternalStream(self.stream, device=dev)

        def buf(nbytes, present):
            return torch.zeros(nbytes, dtype=torch.uint8, device=dev) if present else None

        # [lo, hi] per direction
        self.halo_out = [buf(self.halo_bytes, self.has_lo), buf(self.halo_bytes, self.has_hi)]
        self.halo_in = [buf(self.halo_bytes, self.has_lo), buf(self.halo_bytes, self.has_hi)]
        self.mig_out = [buf(self.mig_bytes, self.has_lo), buf(self.mig_bytes, self.has_hi)]
        self.mig_in = [buf(self.mig_bytes, self.has_lo), buf(self.mig_bytes, self.has_hi)]
        self.last_migration = (0, 0, 0, 0)  # out lo, out hi, in lo, in hi

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    # -- the six phases of one step ---------------------------------------------------------------
    def halo_export(self):
        self._check(self._lib.crb_slab_halo_export(self._h, self._p(self.halo_out[0]), self._p(self.halo_out[1])))

    def halo_import(self):
        self._check(self._lib.crb_slab_halo_import(self._h, self._p(self.halo_in[0]), self._p(self.halo_in[1])))

    def compute(self, save_prev: bool = True):
        self._check(self._lib.crb_slab_compute(self._h, 1 if save_prev else 0))

    def compute_interior(self, save_prev: bool = True):
        """planes without a ghost neighbour, on the handle's second stream (needs no halo)"""
        self._check(self._lib.crb_slab_compute_interior(self._h, 1 if save_prev else 0))

    def compute_boundary(self):
        """planes next to a ghost plane, after halo_import; joins the second stream"""
        self._check(self._lib.crb_slab_compute_boundary(self._h))

    def migration_export(self):
        n = (C.c_uint64 * 2)()
        self._check(self._lib.crb_slab_migration_export(self._h, self._p(self.mig_out[0]), self._p(self.mig_out[1]), n))
        return int(n[0]), int(n[1])

    def migration_export_async(self):
        """enqueue only: the counts come back with the arrival counts (one host sync per step)"""
        self._check(self._lib.crb_slab_migration_export(self._h, self._p(self.mig_out[0]), self._p(self.mig_out[1]), None))

    def migration_out_counts(self):
        n = (C.c_uint64 * 2)()
        self._check(self._lib.crb_slab_migration_counts(self._h, n))
        return int(n[0]), int(n[1])

    def migration_import(self):
        n = (C.c_uint64 * 2)()
        self._check(self._lib.crb_slab_migration_import(self._h, self._p(self.mig_in[0]), self._p(self.mig_in[1]), n))
        return int(n[0]), int(n[1])

    def finish(self, save_prev: bool = True):
        self._check(self._lib.crb_slab_finish(self._h, 1 if save_prev else 0))

    # -- crb_slab_step: exchange over peer memory, no host synchronisation ------------------------------
    @property
    def arena(self) -> int:
        return int(self._lib.crb_slab_arena(self._h) or 0)

    def ipc_handle(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        self._check(self._lib.crb_slab_ipc_handle(self._h, buf))
        return bytes(buf)

    def connect(self, lo_arena: int | None, hi_arena: int | None):
        """neighbours in THIS process: raw device pointers of their arenas"""
        self._check(self._lib.crb_slab_connect(self._h, C.c_void_p(lo_arena) if lo_arena else None,
                                               C.c_void_p(hi_arena) if hi_arena else None))

    def connect_ipc(self, lo_handle: bytes | None, hi_handle: bytes | None):
        """neighbours in other processes of the node: their CUDA IPC handles"""
        lo = (C.c_ubyte * 64).from_buffer_copy(lo_handle) if lo_handle else None
        hi = (C.c_ubyte * 64).from_buffer_copy(hi_handle) if hi_handle else None
        self._check(self._lib.crb_slab_connect_ipc(self._h, lo, hi))

    def slab_step(self, n_steps: int = 1):
        self._check(self._lib.crb_slab_step(self._h, n_steps))

    def slab_step_enqueue(self, n_steps: int = 1):
        """queue only (crb_slab_step_async); :meth:`sync` completes the batch"""
        self._check(self._lib.crb_slab_step_async(self._h, n_steps))

    def step(self, n_steps: int = 1):  # pragma: no cover - guarded in C as well
        raise RuntimeError("a slab is stepped by its group (LocalSlabGroup.step / SlabSimulation.step)")


def uniform_row(mp, ip):
    """(mech_row, inter_row) when every agent shares one parameter row, else None (also None for an empty
    population or missing arrays)"""
    if mp is None or ip is None or len(mp) == 0:
        return None
    if (mp == mp[0]).all() and (ip == ip[0]).all():
        return np.array(mp[0], dtype=np.float64), np.array(ip[0], dtype=np.float64)
    return None


def split_agents_by_slab(cfg: _abi.CrbConfig, world: int, pos: np.ndarray) -> np.ndarray:
    return rank_of_position(pos[:, 0], cfg.domain_min[0], cfg.dx[0], int(cfg.n_voxels[0]), world)


class LocalSlabGroup:
    """``world`` slabs driven by one process.  ``devices[r]`` is the CUDA device of rank r (all 0 to
    run several slabs on one GPU)."""

    def __init__(self, make_cfg, world: int, devices: Sequence[int] | None = None, overlap: bool = True,
                 transport: str = "copy", **caps):
        """``transport="copy"``: the six-phase protocol, buffers moved by device copies from this thread;
        ``transport="p2p"``: ``crb_slab_step`` — the export kernels store into the neighbours' arenas and the
        ranks synchronise through device-side flags (one step per rank, round robin: include/crb.h)."""
        devices = list(devices) if devices is not None else [0] * world
        self.world = world
        self.overlap = overlap
        self.transport = transport
        self.ranks = [SlabRank(make_cfg(), r, world, devices[r], **caps) for r in range(world)]
        self.cfg = self.ranks[0].cfg
        self.dim = self.ranks[0].dim
        if transport == "p2p":
            for r, rk in enumerate(self.ranks):
                rk.connect(self.ranks[r - 1].arena if r > 0 else None,
                           self.ranks[r + 1].arena if r + 1 < world else None)
        elif transport != "copy":
            raise ValueError("transport must be 'copy' or 'p2p'")

    def close(self):
        for r in self.ranks:
            r.close()

    # -- population -------------------------------------------------------------------------------
    def set_agents(self, pos, vel, mech_params, inter_params, species=None, ids=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = pos.shape[0]
        owner = split_agents_by_slab(self.cfg, self.world, pos)
        ids = np.arange(n, dtype=np.uint64) if ids is None else np.asarray(ids, dtype=np.uint64)

        def sel(a, m, width=None):
            if a is None:
                return None
            a = np.asarray(a)
            if width is not None:
                a = np.broadcast_to(a, (n, width))
            return np.ascontiguousarray(a[m])

        # uniform vs per-agent parameters is decided on the WHOLE population and declared on every slab: a slab that
        # starts empty, or whose own agents happen to share one row, must not differ from its neighbours (the halo /
        # migration buffers of a uniform slab carry no parameter columns; crb_slab_*_import checks the mode word)
        rk0 = self.ranks[0]
        mp = None if mech_params is None else np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, rk0.n_mech))
        ip = None if inter_params is None else np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, rk0.n_inter))
        row = uniform_row(mp, ip)
        for r, rk in enumerate(self.ranks):
            m = owner == r
            if row is not None:
                rk.set_uniform_params(*row)
                rk.set_agents(pos[m], sel(vel, m), None, None, sel(species, m), ids[m])
            elif mp is None or ip is None:
                # rows declared earlier with set_uniform_params on every rank
                rk.set_agents(pos[m], sel(vel, m), None, None, sel(species, m), ids[m])
            else:
                rk.set_uniform_params(None, None)
                rk.set_agents(pos[m], sel(vel, m), np.ascontiguousarray(mp[m]), np.ascontiguousarray(ip[m]),
                              sel(species, m), ids[m])

    def set_uniform_params(self, mech_row=None, inter_row=None):
        for rk in self.ranks:
            rk.set_uniform_params(mech_row, inter_row)

    def set_noise(self, z):
        for rk in self.ranks:
            rk.set_noise(z)

    # -- exchange ---------------------------------------------------------------------------------
    def _exchange(self, out_attr: str, in_attr: str):
        """rank r's 'hi' out-buffer -> rank r+1's 'lo' in-buffer and vice versa (device copies ordered
        after the producer's stream and before the consumer's)."""
        import torch

        for r in range(self.world - 1):
            a, b = self.ranks[r], self.ranks[r + 1]
            for src_rank, src, dst_rank, dst in (
                (a, getattr(a, out_attr)[1], b, getattr(b, in_attr)[0]),
                (b, getattr(b, out_attr)[0], a, getattr(a, in_attr)[1]),
            ):
                with torch.cuda.stream(src_rank.torch_stream):
                    dst.copy_(src, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(src_rank.torch_stream)
                dst_rank.torch_stream.wait_event(ev)

    def halo_exchange(self):
        for rk in self.ranks:
            rk.halo_export()
        self._exchange("halo_out", "halo_in")
        for rk in self.ranks:
            rk.halo_import()

    def step(self, n_steps: int = 1):
        if self.transport == "p2p":
            # one step per rank, round robin: a rank's queue waits (on the device) for exports of its neighbours,
            # which this thread has to queue as well; nothing synchronises until every step of every rank is queued
            for it in range(n_steps):
                for rk in self.ranks:
                    rk.slab_step_enqueue(1)
                if it % 8 == 7:  # keep the launch queues short (a full queue would block this thread mid-round)
                    for rk in self.ranks:
                        rk.sync()
            for rk in self.ranks:
                rk.sync()
            return
        for it in range(n_steps):
            last = it + 1 == n_steps
            if self.overlap:
                # interior planes start at once; the halo travels meanwhile
                for rk in self.ranks:
                    rk.compute_interior(last)
                self.halo_exchange()
                for rk in self.ranks:
                    rk.compute_boundary()
            else:
                self.halo_exchange()
                for rk in self.ranks:
                    rk.compute(last)
            for rk in self.ranks:
                rk.migration_export_async()
            self._exchange("mig_out", "mig_in")
            ins = [rk.migration_import() for rk in self.ranks]
            outs = [rk.migration_out_counts() for rk in self.ranks]
            for rk, o, i in zip(self.ranks, outs, ins):
                rk.last_migration = (*o, *i)
                rk.finish(last)
        for rk in self.ranks:
            rk.sync()

    # -- gathered views (device order inside a slab, slabs in rank order) --------------------------
    @property
    def n_agents(self) -> int:
        return sum(rk.n_agents for rk in self.ranks)

    def get_agents(self):
        parts = [rk.get_agents() for rk in self.ranks]
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(4))

    def get_aux(self):
        parts = [rk.get_aux() for rk in self.ranks]
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(3))

    def debug_forces(self):
        self.halo_exchange()
        return np.concatenate([rk.debug_forces() for rk in self.ranks])

    def get_voxel_counts(self):
        return np.concatenate([rk.get_voxel_counts()[: rk.n_planes * self._plane()] for rk in self.ranks])

    def _plane(self):
        nv = [int(self.cfg.n_voxels[i]) for i in range(self.dim)]
        return int(np.prod(nv[1:]))

    def count_candidate_pairs(self):
        self.halo_exchange()
        return sum(rk.count_candidate_pairs() for rk in self.ranks)

    @property
    def kernel_launches(self):
        return sum(rk.kernel_launches for rk in self.ranks)


class RodSlabRank(SlabRank):
    """One slab of bacterial rods (``mechanics_kind == CRB_MECH_ROD``): the slab protocol of :class:`SlabRank`
    with the rod-shaped setters / getters of :class:`cellular_raza_b200.rods.RodSimulation`."""

    def __init__(self, cfg: _abi.CrbConfig, rank: int, world: int, device: int = 0, **caps):
        if int(cfg.mechanics_kind) != _abi.CRB_MECH_ROD:
            raise ValueError("RodSlabRank needs mechanics_kind == CRB_MECH_ROD")
        super().__init__(cfg, rank, world, device, **caps)
        self.V = int(cfg.rod_n_vertices)


def _rod_methods():
    # the rod-shaped accessors are plain functions of (self._lib, self._h, self.V, self.dim): shared, not copied
    from .rods import RodSimulation

    for name in ("set_agents", "get_agents", "get_params", "get_aux", "set_noise", "debug_forces",
                 "debug_neighbor_counts"):
        setattr(RodSlabRank, name, getattr(RodSimulation, name))
    RodSlabRank.n_divisions = RodSimulation.n_divisions


_rod_methods()


class LocalRodSlabGroup(LocalSlabGroup):
    """``world`` slabs of rods driven by one process through the six-phase protocol (halo: the vertex positions and
    bounding spheres of the boundary planes; migration: complete rods incl. identifiers and AuxStorage; divisions
    happen on the slab that owns the mother's voxel, as ``Voxel::id_counter`` demands).  Gathered views come back
    in rank order = ascending global voxel, chili in-voxel order: the order of a single handle."""

    def __init__(self, make_cfg, world: int, devices: Sequence[int] | None = None, **caps):
        devices = list(devices) if devices is not None else [0] * world
        self.world = world
        self.overlap = False  # (the rod force pass is not split into interior / boundary planes)
        self.transport = "copy"
        self.ranks = [RodSlabRank(make_cfg(), r, world, devices[r], **caps) for r in range(world)]
        self.cfg = self.ranks[0].cfg
        self.dim = self.ranks[0].dim
        self.V = self.ranks[0].V

    def set_agents(self, pos, vel, mech_params, inter_params, ids=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = pos.shape[0]
        pos = pos.reshape(n, self.V, 3)
        # SortCells::get_voxel_index_of of CartesianCuboidRods: mean of the vertices (bacterial_rods.rs:368-372); the
        # sum runs over the vertices in order, like nalgebra's row_sum and like the device
        mean0 = np.zeros(n)
        for v in range(self.V):
            mean0 = mean0 + pos[:, v, 0]
        mean0 = mean0 / float(self.V)
        owner = rank_of_position(mean0, self.cfg.domain_min[0], self.cfg.dx[0], int(self.cfg.n_voxels[0]), self.world)
        ids = np.arange(n, dtype=np.uint64) if ids is None else np.asarray(ids, dtype=np.uint64)
        rk0 = self.ranks[0]
        mp = np.ascontiguousarray(np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, rk0.n_mech)))
        ip = np.ascontiguousarray(np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, rk0.n_inter)))
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64).reshape(n, self.V, 3)
        for r, rk in enumerate(self.ranks):
            m = owner == r
            # every rank declares the uniform rows (spring_length and growth_rate are per rod and come with the rows)
            rk.set_uniform_params(mp[0] if n else None, ip[0] if n else None)
            rk.set_agents(pos[m], None if vel is None else vel[m], mp[m], ip[m], ids=ids[m])

    def get_agents(self):
        parts = [rk.get_agents() for rk in self.ranks]
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(5))

    def get_params(self):
        return np.concatenate([rk.get_params() for rk in self.ranks])

    def debug_neighbor_counts(self):
        self.halo_exchange()
        return np.concatenate([rk.debug_neighbor_counts() for rk in self.ranks])

    @property
    def n_divisions(self) -> int:
        return sum(rk.n_divisions for rk in self.ranks)


class SlabSimulation:
    """One slab per process (``torch.distributed`` must be initialised with the nccl backend); halo and
    migration buffers travel with ncclSend/ncclRecv over NVLink."""

    def __init__(self, cfg: _abi.CrbConfig, device: int | None = None, group=None, overlap: bool = True,
                 transport: str = "nccl", **caps):
        """``transport="nccl"``: the six-phase protocol with ncclSend/ncclRecv between the phases (one host
        synchronisation per step); ``transport="p2p"``: ``crb_slab_step`` — the neighbours' exchange arenas are mapped
        through CUDA IPC and the export kernels store into them over NVLink, no host synchronisation inside a batch."""
        import torch
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.overlap = overlap
        self.transport = transport
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        device = torch.cuda.current_device() if device is None else device
        self.rods = int(cfg.mechanics_kind) == _abi.CRB_MECH_ROD
        if self.rods:
            # rods: six-phase protocol only (division counts are read back every step; the force pass is not split)
            self.overlap, self.transport, transport = False, "nccl", "nccl"
            self.slab = RodSlabRank(cfg, self.rank, self.world, device, **caps)
        else:
            self.slab = SlabRank(cfg, self.rank, self.world, device, **caps)
        # the exchange buffers have a fixed size derived from the capacities: a rank with another size
        # would leave its neighbour waiting in ncclRecv forever, so compare before the first exchange
        mine = torch.tensor([self.slab.halo_bytes, self.slab.mig_bytes], dtype=torch.int64,
                            device=torch.device("cuda", device))
        every = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=group)
        if any(not torch.equal(t, mine) for t in every):
            sizes = [tuple(int(v) for v in t.tolist()) for t in every]
            self.slab.close()
            raise _abi.SimulationError(_abi.CRB_ERR_INVALID_ARG,
                                       f"halo/migration buffer sizes differ between ranks {sizes}: create every "
                                       "slab with the same capacity, ghost_capacity and migration_capacity")
        self.cfg = self.slab.cfg
        self.dim = self.slab.dim
        self._declared = False
        if transport == "p2p":
            handles = [None] * self.world
            dist.all_gather_object(handles, self.slab.ipc_handle(), group=group)
            self.slab.connect_ipc(handles[self.rank - 1] if self.rank > 0 else None,
                                  handles[self.rank + 1] if self.rank + 1 < self.world else None)
            dist.barrier(group=group)  # every rank has reset its flags before anyone steps
        elif transport != "nccl":
            raise ValueError("transport must be 'nccl' or 'p2p'")

    def close(self):
        self.slab.close()

    def set_agents(self, pos, vel, mech_params, inter_params, species=None, ids=None):
        """``pos`` … are THIS rank's agents (axis-0 coordinate inside the slab).  When parameter arrays are given,
        whether all agents of ALL ranks share one row is agreed on with one all_gather: every rank then declares
        that row (or none does) — a rank never decides from its own subset (an empty slab has no row at all)."""
        if self.rods:
            # rods: uniform rows except spring_length / growth_rate, declared on every rank (row 0 of the first rank
            # that holds a rod); the per-rod rows travel with set_agents
            n = len(pos)
            mp = np.ascontiguousarray(np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, self.slab.n_mech)))
            ip = np.ascontiguousarray(np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, self.slab.n_inter)))
            every = [None] * self.world
            self.dist.all_gather_object(every, (mp[0].tolist(), ip[0].tolist()) if n else None, group=self.group)
            rows = [e for e in every if e is not None]
            if not rows:
                raise ValueError("no rank holds a rod")
            self.slab.set_uniform_params(np.array(rows[0][0]), np.array(rows[0][1]))
            self.slab.set_agents(pos, vel, mp, ip, ids=ids)
            return
        if mech_params is not None and inter_params is not None and not self._declared:
            n = len(pos)
            mp = np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, self.slab.n_mech))
            ip = np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, self.slab.n_inter))
            row = uniform_row(mp, ip)
            if n == 0:
                mine = ("empty",)
            elif row is None:
                mine = ("mixed",)
            else:
                mine = ("row", row[0].tolist(), row[1].tolist())
            every = [None] * self.world
            self.dist.all_gather_object(every, mine, group=self.group)
            rows = [e for e in every if e[0] == "row"]
            mixed = any(e[0] == "mixed" for e in every)
            same = bool(rows) and all(r[1:] == rows[0][1:] for r in rows)
            if same and not mixed:
                self.slab.set_uniform_params(np.array(rows[0][1]), np.array(rows[0][2]))
                self.slab.set_agents(pos, vel, None, None, species, ids)
            else:
                self.slab.set_uniform_params(None, None)
                self.slab.set_agents(pos, vel, np.ascontiguousarray(mp), np.ascontiguousarray(ip), species, ids)
            return
        self.slab.set_agents(pos, vel, mech_params, inter_params, species, ids)

    def set_uniform_params(self, mech_row=None, inter_row=None):
        """must be called with the same rows on every rank"""
        self._declared = mech_row is not None or inter_row is not None
        self.slab.set_uniform_params(mech_row, inter_row)

    def _exchange(self, outs, ins):
        import torch

        with torch.cuda.stream(self.slab.torch_stream):
            # stream-ordered for nccl: the waits block the handle's stream, not the host
            neighbor_exchange(self.dist, self.rank, self.world, outs, ins, self.group)

    def halo_exchange(self):
        s = self.slab
        s.halo_export()
        self._exchange(s.halo_out, s.halo_in)
        s.halo_import()

    def step(self, n_steps: int = 1):
        s = self.slab
        if self.transport == "p2p":
            s.slab_step(n_steps)  # the whole batch behind one call of the C ABI
            return
        for it in range(n_steps):
            last = it + 1 == n_steps
            if self.overlap:
                # interior planes on the second stream; pack, ncclSend/ncclRecv and unpack of the halo
                # run meanwhile on the (high-priority) main stream, then the boundary planes
                s.compute_interior(last)
                self.halo_exchange()
                s.compute_boundary()
            else:
                self.halo_exchange()
                s.compute(last)
            # export, ncclSend/ncclRecv of the fixed-size buffers and the count read-back are all enqueued
            # while the force pass still runs: ONE host synchronisation per step (in migration_import)
            s.migration_export_async()
            self._exchange(s.mig_out, s.mig_in)
            inn = s.migration_import()
            out = s.migration_out_counts()
            s.last_migration = (*out, *inn)
            s.finish(last)

    def sync(self):
        self.slab.sync()

    @property
    def n_agents(self):
        return self.slab.n_agents

    def get_agents(self):
        return self.slab.get_agents()
