"""ctypes wrapper of ``oracle/librods_oracle.so`` — TEST INFRASTRUCTURE ONLY.

CPU restatement of chili's step for the bacterial-rods agent (RodMechanics + RodInteraction<Morse>,
aspects [Mechanics, Interaction, Cycle, NeighborSensing]); see rods_oracle.cpp for the file:line
citations.  Only ``tests/``, ``bench.py`` (cpu_baseline / ``--impl reference``) and
``__graft_entry__.smoke()`` may import this module; the product package never does.
PARITY STATUS: pinned on the nearest-point and rod-division doctests of the reference; everything
else on this path has no numeric expectation in the reference (parity unpinned).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from cellular_raza_b200._abi import CrbConfig

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "librods_oracle.so"
_lib = None

N_MECH = 8
N_INTER = 4


def build(force: bool = False) -> Path:
    srcs = [_DIR / "rods_oracle.cpp", _DIR / "chili_common.hpp", _DIR.parent / "include" / "crb.h"]
    if force or not _LIB_PATH.exists() or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_DIR), "-B", "librods_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        L = _lib
        vp = C.c_void_p
        L.rods_oracle_create.restype = vp
        L.rods_oracle_create.argtypes = [C.POINTER(CrbConfig)]
        L.rods_oracle_destroy.argtypes = [vp]
        L.rods_oracle_last_error.restype = C.c_char_p
        L.rods_oracle_set_threads.argtypes = [vp, C.c_int]
        L.rods_oracle_set_agents.argtypes = [vp, C.c_uint64] + [vp] * 5
        L.rods_oracle_n_agents.restype = C.c_uint64
        L.rods_oracle_n_agents.argtypes = [vp]
        L.rods_oracle_n_divisions.restype = C.c_uint64
        L.rods_oracle_n_divisions.argtypes = [vp]
        L.rods_oracle_step.argtypes = [vp, C.c_uint64]
        L.rods_oracle_set_noise.argtypes = [vp, vp, C.c_uint64, C.c_uint64]
        L.rods_oracle_forces.argtypes = [vp] * 4
        L.rods_oracle_get_agents.argtypes = [vp] * 6
        L.rods_oracle_get_params.argtypes = [vp, vp]
        L.rods_oracle_get_aux.argtypes = [vp] * 4
        L.rods_oracle_get_voxel_counts.argtypes = [vp, vp]
        L.rods_oracle_count_candidate_pairs.restype = C.c_uint64
        L.rods_oracle_count_candidate_pairs.argtypes = [vp]
        L.rods_oracle_nearest_point.argtypes = [C.c_uint32] + [vp] * 6
        L.rods_oracle_divide.argtypes = [C.c_uint32, C.c_uint32, vp, C.c_double, C.c_double, vp, vp, vp]
        L.rods_oracle_increment.argtypes = [C.c_uint32, C.c_uint32] + [vp] * 6
        L.rods_oracle_pair_force.argtypes = [C.c_uint32, C.c_uint32, vp, vp, vp, C.c_double, vp, vp]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class RodsOracle:
    """Same method surface as ``cellular_raza_b200.device.DeviceSimulation`` for CRB_MECH_ROD."""

    def __init__(self, cfg: CrbConfig, n_threads: int = 1):
        self.L = lib()
        self.cfg = cfg
        self.dim = int(cfg.dim)
        self.V = int(cfg.rod_n_vertices)
        self.h = self.L.rods_oracle_create(C.byref(cfg))
        if not self.h:
            raise OracleError(-16, self.L.rods_oracle_last_error().decode())
        self.L.rods_oracle_set_threads(self.h, n_threads)

    def close(self):
        if self.h:
            self.L.rods_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OracleError(rc, self.L.rods_oracle_last_error().decode())

    def set_agents(self, pos, vel, mech_params, inter_params, species=None, ids=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = pos.shape[0]
        assert pos.shape == (n, self.V, self.dim)
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        mp = np.ascontiguousarray(np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, N_MECH)))
        ip = np.ascontiguousarray(np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, N_INTER)))
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        self._check(self.L.rods_oracle_set_agents(self.h, n, _p(pos), _p(vel), _p(mp), _p(ip), _p(ids)))

    @property
    def n_agents(self):
        return int(self.L.rods_oracle_n_agents(self.h))

    @property
    def n_divisions(self):
        return int(self.L.rods_oracle_n_divisions(self.h))

    def set_noise(self, z):
        """z[step][row][c * V + v] standard normal numbers (column-major fill of the nrows x D matrix, the order
        nalgebra's from_distribution consumes the generator in)"""
        z = np.ascontiguousarray(z, dtype=np.float64)
        self._check(self.L.rods_oracle_set_noise(self.h, _p(z), z.shape[0], z.shape[1]))

    def step(self, n_steps=1):
        self._check(self.L.rods_oracle_step(self.h, n_steps))

    def forces(self):
        """-> (force[n,V,D], abs_sum[n], neighbours[n]) in (voxel, in-voxel) order; state untouched."""
        n = self.n_agents
        f = np.empty((n, self.V, self.dim))
        a = np.empty(n)
        nb = np.empty(n, dtype=np.uint64)
        self._check(self.L.rods_oracle_forces(self.h, _p(f), _p(a), _p(nb)))
        return f, a, nb

    def get_agents(self):
        """-> pos[n,V,D], vel[n,V,D], ident[n,3], parent[n,3], voxel[n]"""
        n = self.n_agents
        pos = np.empty((n, self.V, self.dim))
        vel = np.empty((n, self.V, self.dim))
        ident = np.empty((n, 3), dtype=np.uint64)
        parent = np.empty((n, 3), dtype=np.uint64)
        vox = np.empty(n, dtype=np.uint64)
        self._check(self.L.rods_oracle_get_agents(self.h, _p(pos), _p(vel), _p(ident), _p(parent), _p(vox)))
        return pos, vel, ident, parent, vox

    def get_params(self):
        n = self.n_agents
        mp = np.empty((n, N_MECH))
        self._check(self.L.rods_oracle_get_params(self.h, _p(mp)))
        return mp

    def get_aux(self):
        n = self.n_agents
        hx = np.empty((n, 2, self.V, self.dim))
        hv = np.empty((n, 2, self.V, self.dim))
        hl = np.empty(n, dtype=np.uint32)
        self._check(self.L.rods_oracle_get_aux(self.h, _p(hx), _p(hv), _p(hl)))
        return hx, hv, hl

    def get_voxel_counts(self):
        nv = 1
        for i in range(self.dim):
            nv *= int(self.cfg.n_voxels[i])
        c = np.empty(nv, dtype=np.uint32)
        self._check(self.L.rods_oracle_get_voxel_counts(self.h, _p(c)))
        return c

    def count_candidate_pairs(self):
        return int(self.L.rods_oracle_count_candidate_pairs(self.h))


# ---- known-answer taps -----------------------------------------------------------------------------
def nearest_point(p, a, b):
    p = np.ascontiguousarray(p, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    d = len(p)
    dist, t = C.c_double(), C.c_double()
    out = np.zeros(d)
    lib().rods_oracle_nearest_point(d, _p(p), _p(a), _p(b), C.addressof(dist), _p(out), C.addressof(t))
    return dist.value, out, t.value


def divide(pos, spring_length, radius):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    V, d = pos.shape
    p1, p2 = np.zeros_like(pos), np.zeros_like(pos)
    nl = C.c_double()
    rc = lib().rods_oracle_divide(d, V, _p(pos), spring_length, radius, _p(p1), _p(p2), C.addressof(nl))
    if rc != 0:
        raise OracleError(rc, "DivisionError")
    return p1, p2, nl.value


def increment(pos, vel, mech_row, force=None):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    vel = np.ascontiguousarray(vel, dtype=np.float64)
    V, d = pos.shape
    mr = np.ascontiguousarray(mech_row, dtype=np.float64)
    f = None if force is None else np.ascontiguousarray(force, dtype=np.float64)
    dx, dv = np.zeros_like(pos), np.zeros_like(pos)
    lib().rods_oracle_increment(d, V, _p(pos), _p(vel), _p(mr), _p(f), _p(dx), _p(dv))
    return dx, dv


def pair_force(own, ext, inter_row, ext_radius):
    own = np.ascontiguousarray(own, dtype=np.float64)
    ext = np.ascontiguousarray(ext, dtype=np.float64)
    V, d = own.shape
    ir = np.ascontiguousarray(inter_row, dtype=np.float64)
    fo, fe = np.zeros_like(own), np.zeros_like(own)
    lib().rods_oracle_pair_force(d, V, _p(own), _p(ext), _p(ir), ext_radius, _p(fo), _p(fe))
    return fo, fe
