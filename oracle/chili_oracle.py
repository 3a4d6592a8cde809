"""ctypes wrapper of ``oracle/libchili_oracle.so`` — TEST INFRASTRUCTURE ONLY.

CPU restatement of cellular_raza's ``chili`` mechanics/interaction step (see chili_oracle.cpp).
Only ``tests/``, ``bench.py`` (cpu_baseline / ``--impl reference``) and
``__graft_entry__.smoke()`` may import this module; the product package never does.
PARITY STATUS: pinned on the reference's known-answer tests; pair forces, cross-voxel weights and
the AB wrapper recurrence are parity-unpinned (no numeric expectations exist in the reference).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from cellular_raza_b200._abi import N_INTER_PARAMS, N_MECH_PARAMS, CrbConfig

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "libchili_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    srcs = [_DIR / "chili_oracle.cpp", _DIR / "chili_common.hpp", _DIR.parent / "include" / "crb.h"]
    if force or not _LIB_PATH.exists() or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_DIR), "-B", "libchili_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        L = _lib
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(CrbConfig)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_agents.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 6
        L.oracle_set_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
        L.oracle_n_agents.restype = C.c_uint64
        L.oracle_n_agents.argtypes = [C.c_void_p]
        L.oracle_step.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_forces.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_get_agents.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.oracle_get_aux.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.oracle_n_migrations.restype = C.c_uint64
        L.oracle_n_migrations.argtypes = [C.c_void_p]
        L.oracle_get_migrations.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.oracle_get_voxel_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_count_candidate_pairs.restype = C.c_uint64
        L.oracle_count_candidate_pairs.argtypes = [C.c_void_p]
        L.oracle_decomp_res.argtypes = [C.c_uint64, C.c_uint64] + [C.POINTER(C.c_uint64)] * 3
        L.oracle_domain_from_interaction_range.argtypes = [
            C.c_uint32, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_domain_from_n_voxels.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_subdomain_sizes.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_sub_of.restype = C.c_uint64
        L.oracle_sub_of.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_voxel_index.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_neighbors.restype = C.c_uint64
        L.oracle_neighbors.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        for name, nargs in (("oracle_helper_euler", 4), ("oracle_helper_ab2", 5), ("oracle_helper_ab3", 6),
                            ("oracle_xapy", 3)):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_double] * nargs
        L.oracle_ring.restype = C.c_uint32
        L.oracle_ring.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_pair_force.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint8,
                                        C.c_double, C.c_uint8, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class ChiliOracle:
    """Same method surface as ``cellular_raza_b200.device.DeviceSimulation``."""

    def __init__(self, cfg: CrbConfig, n_threads: int = 1):
        self.L = lib()
        self.cfg = cfg
        self.dim = int(cfg.dim)
        self.n_mech = N_MECH_PARAMS[int(cfg.mechanics_kind)]
        self.n_inter = N_INTER_PARAMS[int(cfg.interaction_kind)]
        self.h = self.L.oracle_create(C.byref(cfg))
        if not self.h:
            raise OracleError(-5, self.L.oracle_last_error().decode())
        self.L.oracle_set_threads(self.h, n_threads)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OracleError(rc, self.L.oracle_last_error().decode())

    def set_agents(self, pos, vel, mech_params, inter_params, species=None, ids=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        mp = np.ascontiguousarray(np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, self.n_mech)))
        ip = np.ascontiguousarray(np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, self.n_inter)))
        sp = None if species is None else np.ascontiguousarray(species, dtype=np.uint8)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        self._check(self.L.oracle_set_agents(self.h, n, _p(pos), _p(vel), _p(mp), _p(ip), _p(sp), _p(ids)))

    def set_noise(self, z):
        z = np.ascontiguousarray(z, dtype=np.float64)
        self._check(self.L.oracle_set_noise(self.h, _p(z), z.shape[0], z.shape[1]))

    @property
    def n_agents(self):
        return int(self.L.oracle_n_agents(self.h))

    def step(self, n_steps=1):
        self._check(self.L.oracle_step(self.h, n_steps))

    def forces(self):
        """-> (force[n,D], abs_sum[n]) in (voxel, in-voxel) order; state untouched."""
        n = self.n_agents
        f = np.empty((n, self.dim))
        a = np.empty(n)
        self._check(self.L.oracle_forces(self.h, _p(f), _p(a)))
        return f, a

    def get_agents(self):
        n = self.n_agents
        pos = np.empty((n, self.dim))
        vel = np.empty((n, self.dim))
        ids = np.empty(n, dtype=np.uint64)
        vox = np.empty(n, dtype=np.uint64)
        self._check(self.L.oracle_get_agents(self.h, _p(pos), _p(vel), _p(ids), _p(vox)))
        return pos, vel, ids, vox

    def get_aux(self):
        n = self.n_agents
        hx = np.empty((n, 2, self.dim))
        hv = np.empty((n, 2, self.dim))
        hl = np.empty(n, dtype=np.uint32)
        self._check(self.L.oracle_get_aux(self.h, _p(hx), _p(hv), _p(hl)))
        return hx, hv, hl

    def get_migrations(self):
        m = int(self.L.oracle_n_migrations(self.h))
        ids = np.empty(m, dtype=np.uint64)
        old = np.empty(m, dtype=np.uint64)
        new = np.empty(m, dtype=np.uint64)
        self._check(self.L.oracle_get_migrations(self.h, _p(ids), _p(old), _p(new)))
        return ids, old, new

    def get_voxel_counts(self):
        nv = 1
        for i in range(self.dim):
            nv *= int(self.cfg.n_voxels[i])
        c = np.empty(nv, dtype=np.uint32)
        self._check(self.L.oracle_get_voxel_counts(self.h, _p(c)))
        return c

    def count_candidate_pairs(self):
        return int(self.L.oracle_count_candidate_pairs(self.h))

    def sub_of(self, plain):
        return int(self.L.oracle_sub_of(self.h, plain))

    def voxel_index(self, pos, setup_variant=False):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        idx = np.zeros(3, dtype=np.uint64)
        rc = self.L.oracle_voxel_index(self.h, _p(pos), _p(idx), 1 if setup_variant else 0)
        return None if rc != 0 else idx[: self.dim].copy()

    def neighbors(self, plain):
        out = np.empty(27, dtype=np.uint64)
        n = self.L.oracle_neighbors(self.h, plain, _p(out))
        return out[:n].copy()


# ---- known-answer taps -----------------------------------------------------------------------------
def decomp_res(n_voxel, n_regions):
    n, m, a = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib().oracle_decomp_res(n_voxel, n_regions, C.byref(n), C.byref(m), C.byref(a))
    return None if rc != 0 else (n.value, m.value, a.value)


def domain_from_interaction_range(mn, mx, rng):
    mn = np.ascontiguousarray(mn, dtype=np.float64)
    mx = np.ascontiguousarray(mx, dtype=np.float64)
    d = len(mn)
    nv = np.zeros(d, dtype=np.uint64)
    dx = np.zeros(d)
    rc = lib().oracle_domain_from_interaction_range(d, _p(mn), _p(mx), rng, _p(nv), _p(dx))
    return None if rc != 0 else (nv, dx)


def domain_from_n_voxels(mn, mx, nv):
    mn = np.ascontiguousarray(mn, dtype=np.float64)
    mx = np.ascontiguousarray(mx, dtype=np.float64)
    nv = np.ascontiguousarray(nv, dtype=np.uint64)
    dx = np.zeros(len(mn))
    rc = lib().oracle_domain_from_n_voxels(len(mn), _p(mn), _p(mx), _p(nv), _p(dx))
    return None if rc != 0 else dx


def subdomain_sizes(n_voxel, n_regions):
    out = np.zeros(n_regions + 1, dtype=np.uint64)
    k = C.c_uint64()
    rc = lib().oracle_subdomain_sizes(n_voxel, n_regions, _p(out), C.byref(k))
    return None if rc != 0 else out[: k.value].copy()


def ring(values, cap):
    v = np.ascontiguousarray(values, dtype=np.float64)
    out = np.zeros(max(cap, 1))
    n = lib().oracle_ring(_p(v), len(v), cap, _p(out))
    return out[:n].copy()


def pair_force(kind, own_pos, ext_pos, own_params, own_species=0, ext_radius=0.0, ext_species=0):
    op = np.ascontiguousarray(own_pos, dtype=np.float64)
    ep = np.ascontiguousarray(ext_pos, dtype=np.float64)
    pr = np.ascontiguousarray(own_params, dtype=np.float64)
    d = len(op)
    fo = np.zeros(d)
    fe = np.zeros(d)
    rc = lib().oracle_pair_force(kind, d, _p(op), _p(ep), _p(pr), own_species, ext_radius, ext_species, _p(fo), _p(fe))
    if rc != 0:
        raise OracleError(rc, "CalcError")
    return fo, fe


def helper_euler(x, dx, dt, r=0.0):
    return lib().oracle_helper_euler(x, dx, dt, r)


def helper_ab2(x, dx0, dx1, dt, r=0.0):
    return lib().oracle_helper_ab2(x, dx0, dx1, dt, r)


def helper_ab3(x, dx0, dx1, dx2, dt, r=0.0):
    return lib().oracle_helper_ab3(x, dx0, dx1, dx2, dt, r)


def xapy(x, a, y):
    return lib().oracle_xapy(x, a, y)
