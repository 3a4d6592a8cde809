"""Thin object wrapper over the C ABI (``include/crb.h``): one :class:`DeviceSimulation` owns one
``crb_sim`` handle on one GPU.  numpy arrays in, numpy arrays out; all arithmetic happens in the
CUDA library.  Higher-level, reference-shaped API: :mod:`cellular_raza_b200.backend`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._abi import CrbConfig, SimulationError


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def make_config(
    *,
    dim: int,
    mechanics_kind: int,
    interaction_kind: int,
    domain_min,
    domain_max,
    n_voxels,
    dx,
    dt: float,
    capacity: int,
    solver_order: int = 2,
    pair_weight_mode: int = _abi.CRB_WEIGHT_CHILI,
    ref_n_subdomains: int = 1,
    noise_mode: int = _abi.CRB_NOISE_NONE,
    rng_seed: int = 0,
    device: int = 0,
    rod_n_vertices: int = 0,
    domain_force_kind: int = _abi.CRB_DOMAIN_FORCE_NONE,
    domain_force_strength=None,
    domain_force_anchor=None,
) -> CrbConfig:
    cfg = CrbConfig()
    cfg.abi_version = _abi.CRB_ABI_VERSION
    cfg.dim = dim
    cfg.mechanics_kind = mechanics_kind
    cfg.interaction_kind = interaction_kind
    cfg.solver_order = solver_order
    cfg.pair_weight_mode = pair_weight_mode
    cfg.ref_n_subdomains = ref_n_subdomains
    cfg.noise_mode = noise_mode
    cfg.rod_n_vertices = rod_n_vertices
    cfg.device = device
    cfg.rank = 0
    cfg.world_size = 1
    for i in range(dim):
        cfg.n_voxels[i] = int(n_voxels[i])
        cfg.domain_min[i] = float(domain_min[i])
        cfg.domain_max[i] = float(domain_max[i])
        cfg.dx[i] = float(dx[i])
    cfg.dt = float(dt)
    cfg.rng_seed = rng_seed
    cfg.capacity = int(capacity)
    cfg.domain_force_kind = domain_force_kind
    for i in range(dim):
        cfg.domain_force_strength[i] = float(domain_force_strength[i]) if domain_force_strength is not None else 0.0
        cfg.domain_force_anchor[i] = float(domain_force_anchor[i]) if domain_force_anchor is not None else 0.0
    return cfg


class DeviceSimulation:
    """Owns a ``crb_sim*``.  Methods map 1:1 to the entry points of ``include/crb.h``."""

    def __init__(self, cfg: CrbConfig):
        self._lib = _abi.load_library()
        self.cfg = cfg
        self.dim = int(cfg.dim)
        self.n_mech = _abi.N_MECH_PARAMS[int(cfg.mechanics_kind)]
        self.n_inter = _abi.N_INTER_PARAMS[int(cfg.interaction_kind)]
        h = C.c_void_p()
        rc = self._lib.crb_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise SimulationError(rc, (self._lib.crb_last_error(None) or b"").decode())
        self._h = h

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.crb_destroy(self._h)  # waits for a snapshot in flight
            self._h = None
            for ptr, _ in self.__dict__.pop("_pin_pool", {}).values():
                self._lib.crb_host_free(C.c_void_p(ptr))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise SimulationError(rc, (self._lib.crb_last_error(self._h) or b"").decode())

    # -- agents -------------------------------------------------------------------------------
    def set_agents(self, pos, vel, mech_params, inter_params, species=None, ids=None):
        pos = _f64(pos)
        n = pos.shape[0]
        pos = _f64(pos, (n, self.dim))
        vel = None if vel is None else _f64(vel, (n, self.dim))
        if mech_params is not None:
            mech_params = _f64(np.broadcast_to(np.asarray(mech_params, dtype=np.float64), (n, self.n_mech)))
        if inter_params is not None:
            inter_params = _f64(np.broadcast_to(np.asarray(inter_params, dtype=np.float64), (n, self.n_inter)))
        species = None if species is None else np.ascontiguousarray(species, dtype=np.uint8)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
        self._check(
            self._lib.crb_set_agents(
                self._h, n, _ptr(pos), _ptr(vel), _ptr(mech_params), _ptr(inter_params), _ptr(species), _ptr(ids)
            )
        )

    def set_uniform_params(self, mech_row=None, inter_row=None):
        """All agents share one parameter row; set_agents may then omit the parameter arrays."""
        if mech_row is None and inter_row is None:
            self._check(self._lib.crb_set_uniform_params(self._h, None, None))
            return
        m = _f64(mech_row, (self.n_mech,))
        i = _f64(inter_row, (self.n_inter,))
        self._check(self._lib.crb_set_uniform_params(self._h, _ptr(m), _ptr(i)))

    @property
    def stream(self) -> int:
        """cudaStream_t of the handle (for CUDA-event timing)."""
        return int(self._lib.crb_stream(self._h) or 0)

    @property
    def n_agents(self) -> int:
        return int(self._lib.crb_n_agents(self._h))

    @property
    def iteration(self) -> int:
        return int(self._lib.crb_iteration(self._h))

    def set_iteration(self, iteration: int):
        """step counter of the handle (keys the device generator of CRB_NOISE_PHILOX together with seed and id)"""
        self._check(self._lib.crb_set_iteration(self._h, int(iteration)))

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.crb_kernel_launches(self._h))

    @property
    def force_kernel_name(self) -> str:
        return self._lib.crb_force_kernel_name(self._h).decode()

    def get_agents(self):
        """-> (pos[n,D], vel[n,D], ids[n], voxel[n]) in device order."""
        n = self.n_agents
        pos = np.empty((n, self.dim))
        vel = np.empty((n, self.dim))
        ids = np.empty(n, dtype=np.uint64)
        vox = np.empty(n, dtype=np.uint64)
        self._check(self._lib.crb_get_agents(self._h, n, _ptr(pos), _ptr(vel), _ptr(ids), _ptr(vox)))
        return pos, vel, ids, vox

    # -- storage snapshots that do not stop the stepping loop (crb_snapshot_async / _wait) ------------
    def _pinned(self, name, shape, dtype):
        """numpy view of page-locked host memory (crb_host_alloc), one buffer per name, grown on demand"""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(count * dtype.itemsize, 1)
        pool = self.__dict__.setdefault("_pin_pool", {})
        ent = pool.get(name)
        if ent is None or ent[1] < nbytes:
            if ent is not None:
                self._lib.crb_host_free(C.c_void_p(ent[0]))
            ptr = self._lib.crb_host_alloc(nbytes)
            if not ptr:
                raise MemoryError("crb_host_alloc failed")
            ent = (ptr, nbytes)
            pool[name] = ent
        buf = (C.c_char * nbytes).from_address(ent[0])
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def snapshot_async(self, aux: bool = True):
        """starts a snapshot of the current state; ``step`` may be called while it is copied.  Returns a
        token for :meth:`snapshot_wait`."""
        n = self.n_agents
        bufs = {
            "pos": self._pinned("pos", (n, self.dim), np.float64),
            "vel": self._pinned("vel", (n, self.dim), np.float64),
            "ids": self._pinned("ids", (n,), np.uint64),
            "vox": self._pinned("vox", (n,), np.uint64),
        }
        if aux:
            bufs["hx"] = self._pinned("hx", (n, 2, self.dim), np.float64)
            bufs["hv"] = self._pinned("hv", (n, 2, self.dim), np.float64)
            bufs["hl"] = self._pinned("hl", (n,), np.uint32)
        p = lambda k: _ptr(bufs[k]) if k in bufs else None
        self._check(self._lib.crb_snapshot_async(self._h, n, p("pos"), p("vel"), p("ids"), p("vox"),
                                                 p("hx"), p("hv"), p("hl")))
        return bufs

    def snapshot_wait(self, bufs):
        """-> copies (pos, vel, ids, voxel[, hist_dx, hist_dv, hist_len]) of a finished snapshot"""
        n = C.c_uint64(0)
        self._check(self._lib.crb_snapshot_wait(self._h, C.byref(n)))
        m = int(n.value)
        out = [bufs[k][:m].copy() for k in ("pos", "vel", "ids", "vox")]
        if "hx" in bufs:
            out += [bufs[k][:m].copy() for k in ("hx", "hv", "hl")]
        return tuple(out)

    def get_params(self):
        n = self.n_agents
        mp = np.empty((n, self.n_mech))
        ip = np.empty((n, self.n_inter))
        sp = np.empty(n, dtype=np.uint8)
        self._check(self._lib.crb_get_params(self._h, n, _ptr(mp), _ptr(ip), _ptr(sp)))
        return mp, ip, sp

    def get_aux(self):
        """-> (hist_dx[n,2,D], hist_dv[n,2,D], hist_len[n]) ring buffers oldest-first."""
        n = self.n_agents
        hx = np.empty((n, 2, self.dim))
        hv = np.empty((n, 2, self.dim))
        hl = np.empty(n, dtype=np.uint32)
        self._check(self._lib.crb_get_aux(self._h, n, _ptr(hx), _ptr(hv), _ptr(hl)))
        return hx, hv, hl

    def set_aux(self, hist_dx, hist_dv, hist_len):
        n = self.n_agents
        hx = _f64(hist_dx, (n, 2, self.dim))
        hv = _f64(hist_dv, (n, 2, self.dim))
        hl = np.ascontiguousarray(hist_len, dtype=np.uint32)
        self._check(self._lib.crb_set_aux(self._h, n, _ptr(hx), _ptr(hv), _ptr(hl)))

    def set_noise(self, z):
        """z[n_steps, n_ids, D] standard normal numbers."""
        z = _f64(z)
        if z.ndim != 3 or z.shape[2] != self.dim:
            raise ValueError("noise must have shape (n_steps, n_ids, dim)")
        self._check(self._lib.crb_set_noise(self._h, _ptr(z), z.shape[0], z.shape[1]))

    # -- stepping -----------------------------------------------------------------------------
    def step(self, n_steps: int = 1):
        self._check(self._lib.crb_step(self._h, n_steps))

    def sync(self):
        self._check(self._lib.crb_sync(self._h))

    # -- parity taps --------------------------------------------------------------------------
    def debug_forces(self) -> np.ndarray:
        n = self.n_agents
        f = np.empty((n, self.dim))
        self._check(self._lib.crb_debug_forces(self._h, n, _ptr(f)))
        return f

    def get_migrations(self, cap: int | None = None):
        cap = self.n_agents if cap is None else cap
        ids = np.empty(cap, dtype=np.uint64)
        old = np.empty(cap, dtype=np.uint64)
        new = np.empty(cap, dtype=np.uint64)
        cnt = C.c_uint64(0)
        self._check(self._lib.crb_get_migrations(self._h, cap, _ptr(ids), _ptr(old), _ptr(new), C.byref(cnt)))
        m = cnt.value
        return ids[:m], old[:m], new[:m]

    def get_voxel_counts(self) -> np.ndarray:
        nv = 1
        for i in range(self.dim):
            nv *= int(self.cfg.n_voxels[i])
        c = np.empty(nv, dtype=np.uint32)
        self._check(self._lib.crb_get_voxel_counts(self._h, nv, _ptr(c)))
        return c

    def count_candidate_pairs(self) -> int:
        cnt = C.c_uint64(0)
        self._check(self._lib.crb_count_candidate_pairs(self._h, C.byref(cnt)))
        return cnt.value

    def enable_timers(self, on: bool = True):
        self._check(self._lib.crb_enable_timers(self._h, 1 if on else 0))

    def reset_timers(self):
        self._check(self._lib.crb_reset_timers(self._h))

    def stage_time_ms(self, stage: int):
        ms = C.c_double(0)
        ln = C.c_uint64(0)
        self._check(self._lib.crb_stage_time_ms(self._h, stage, C.byref(ms), C.byref(ln)))
        return ms.value, ln.value


def measure_fp64_peak(device: int = 0):
    """-> (TFLOP/s, ms) of the DFMA-only micro-kernel (roofline denominator of the force kernel)."""
    lib = _abi.load_library()
    t = C.c_double(0)
    ms = C.c_double(0)
    rc = lib.crb_measure_fp64_peak(device, C.byref(t), C.byref(ms))
    if rc != 0:
        raise SimulationError(rc, "crb_measure_fp64_peak failed")
    return t.value, ms.value
