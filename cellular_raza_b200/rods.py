"""Bacterial rods on the device: object wrapper over the C ABI for ``CRB_MECH_ROD``.

Mirrors the agent of ``cellular_raza-examples/bacterial_rods/src/main.rs`` (RodMechanics +
RodInteraction<MorsePotential>, aspects [Mechanics, Interaction, Cycle, NeighborSensing]): positions and
velocities are ``(n, V, 3)`` arrays (one row per vertex, like the ``nalgebra::MatrixXx3`` of the
reference), the agent count changes when rods divide, identifiers are ``CellIdentifier`` triples.
All arithmetic happens in the CUDA library; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .device import DeviceSimulation, _f64, _ptr

# mech row of CRB_MECH_ROD (include/crb.h)
ROD_MECH_FIELDS = ("diffusion_constant", "spring_tension", "rigidity", "spring_length", "damping",
                   "growth_rate", "spring_length_threshold", "growth_base_rate")
# CellIdentifier kinds (cellular_raza-concepts/src/cell.rs:18-25)
IDENT_DIVISION, IDENT_INITIAL, IDENT_INSERTED = 0, 1, 2
NO_PARENT = np.uint64(0xFFFFFFFFFFFFFFFF)


class RodSimulation(DeviceSimulation):
    """``crb_sim`` handle created with ``mechanics_kind == CRB_MECH_ROD``."""

    def __init__(self, cfg: _abi.CrbConfig):
        if int(cfg.mechanics_kind) != _abi.CRB_MECH_ROD:
            raise ValueError("RodSimulation needs mechanics_kind == CRB_MECH_ROD")
        super().__init__(cfg)
        self.V = int(cfg.rod_n_vertices)

    def set_agents(self, pos, vel, mech_params, inter_params, species=None, ids=None):
        pos = _f64(pos)
        n = pos.shape[0]
        pos = _f64(pos, (n, self.V, self.dim))
        vel = None if vel is None else _f64(vel, (n, self.V, self.dim))
        mp = _f64(np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, self.n_mech)))
        ip = _f64(np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, self.n_inter)))
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        self._check(self._lib.crb_set_agents(self._h, n, _ptr(pos), _ptr(vel), _ptr(mp), _ptr(ip), None, _ptr(ids)))

    def set_agents_async(self, pos, vel, mech_params, inter_params, ids=None):
        """``crb_set_agents_async``: the rows are staged in page-locked buffers of this handle and only queued; the
        parameter rows are checked on the device and a mismatch is reported by the next ``step``."""
        pos = np.asarray(pos, dtype=np.float64)
        n = pos.shape[0]
        bufs = {"pos": self._pinned("a_pos", (n, self.V, self.dim), np.float64),
                "mp": self._pinned("a_mp", (n, self.n_mech), np.float64),
                "ip": self._pinned("a_ip", (n, self.n_inter), np.float64)}
        bufs["pos"][:] = pos.reshape(n, self.V, self.dim)
        bufs["mp"][:] = np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, self.n_mech))
        bufs["ip"][:] = np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, self.n_inter))
        if vel is not None:
            bufs["vel"] = self._pinned("a_vel", (n, self.V, self.dim), np.float64)
            bufs["vel"][:] = np.asarray(vel, dtype=np.float64).reshape(n, self.V, self.dim)
        if ids is not None:
            bufs["ids"] = self._pinned("a_ids", (n,), np.uint64)
            bufs["ids"][:] = ids
        p = lambda k: _ptr(bufs[k]) if k in bufs else None
        self._check(self._lib.crb_set_agents_async(self._h, n, p("pos"), p("vel"), p("mp"), p("ip"), None, p("ids")))

    def snapshot_async(self, aux: bool = False):
        """``crb_snapshot_async`` for rods: positions, velocities, identifier numbers, voxels (the ring buffers of a rod
        are read with :meth:`get_aux`).  Returns a token for :meth:`snapshot_wait`."""
        if aux:
            raise ValueError("rods: the ring buffers are not part of a snapshot (get_aux)")
        n = self.n_agents
        bufs = {"pos": self._pinned("pos", (n, self.V, self.dim), np.float64),
                "vel": self._pinned("vel", (n, self.V, self.dim), np.float64),
                "ids": self._pinned("ids", (n,), np.uint64), "vox": self._pinned("vox", (n,), np.uint64)}
        self._check(self._lib.crb_snapshot_async(self._h, n, _ptr(bufs["pos"]), _ptr(bufs["vel"]), _ptr(bufs["ids"]),
                                                 _ptr(bufs["vox"]), None, None, None))
        return bufs

    def get_agents(self):
        """-> (pos[n,V,3], vel[n,V,3], ident[n,3], parent[n,3], voxel[n]) in device order."""
        n = self.n_agents
        pos = np.empty((n, self.V, self.dim))
        vel = np.empty((n, self.V, self.dim))
        vox = np.empty(n, dtype=np.uint64)
        self._check(self._lib.crb_get_agents(self._h, n, _ptr(pos), _ptr(vel), None, _ptr(vox)))
        ident = np.empty((n, 3), dtype=np.uint64)
        parent = np.empty((n, 3), dtype=np.uint64)
        self._check(self._lib.crb_get_identifiers(self._h, n, _ptr(ident), _ptr(parent)))
        return pos, vel, ident, parent, vox

    def get_params(self):
        n = self.n_agents
        mp = np.empty((n, self.n_mech))
        self._check(self._lib.crb_get_params(self._h, n, _ptr(mp), None, None))
        return mp

    def get_aux(self):
        n = self.n_agents
        hx = np.empty((n, 2, self.V, self.dim))
        hv = np.empty((n, 2, self.V, self.dim))
        hl = np.empty(n, dtype=np.uint32)
        self._check(self._lib.crb_get_aux(self._h, n, _ptr(hx), _ptr(hv), _ptr(hl)))
        return hx, hv, hl

    def set_noise(self, z):
        """z[n_steps, n_ids, 3 V] standard normal numbers, entry ``c * V + v`` of a row feeding coordinate c of vertex
        v (nalgebra's ``from_distribution`` fills the V x 3 matrix column by column, bacterial_rods.rs:230-236).  The
        row of a rod is its ``Initial(k)`` number k, a daughter's row is a fixed hash of ``Division(voxel, counter)``;
        rows are taken modulo ``n_ids``."""
        z = _f64(z)
        if z.ndim != 3 or z.shape[2] != self.dim * self.V:
            raise ValueError("noise must have shape (n_steps, n_ids, dim * V)")
        self._check(self._lib.crb_set_noise(self._h, _ptr(z), z.shape[0], z.shape[1]))

    @property
    def n_divisions(self) -> int:
        """Division events carried out by the last step."""
        return int(self._lib.crb_n_divisions(self._h))

    def debug_forces(self) -> np.ndarray:
        n = self.n_agents
        f = np.empty((n, self.V, self.dim))
        self._check(self._lib.crb_debug_forces(self._h, n, _ptr(f)))
        return f

    def debug_neighbor_counts(self) -> np.ndarray:
        n = self.n_agents
        c = np.empty(n, dtype=np.uint64)
        self._check(self._lib.crb_debug_neighbor_counts(self._h, n, _ptr(c)))
        return c


def ident_key(ident: np.ndarray) -> np.ndarray:
    """Sortable scalar-ish key of identifier triples (for matching device and oracle order)."""
    ident = np.asarray(ident, dtype=np.uint64)
    return np.lexsort((ident[:, 2], ident[:, 1], ident[:, 0]))
