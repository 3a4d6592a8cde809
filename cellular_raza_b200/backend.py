"""Host-side mirror of the reference's user-facing interface for the mechanics/interaction path.

A simulation written against cellular_raza's concepts API

    run_simulation!(domain: domain, agents: agents, settings: settings, aspects: [Mechanics, Interaction])

(cellular_raza-examples/cell_sorting/src/main.rs:165-170; macro kwargs
cellular_raza-core-proc-macro/src/run_sim.rs:131-159) maps 1:1 onto :func:`run_simulation` here: the
same names (``CartesianCuboid``, ``NewtonDamped3D``, ``Brownian3D``, ``Langevin3D``, ``MorsePotential``,
``BoundLennardJones``, ``MiePotential``, ``Settings``, ``FixedStepsize``, ``StorageBuilder``,
``StorageAccess``, ``SimulationError``), the same argument meaning and the same error behaviour.  The
generated main loop (run_sim.rs:380-403) becomes ``crb_step`` between save events and
``crb_get_agents``/``crb_get_aux`` at ``PartialSave`` iterations.  All arithmetic happens in
``libcrb200.so``; this file only packs agents into parameter rows and writes snapshots in the
reference's on-disk layout (cellular_raza-core/src/storage/concepts.rs:156-170, 702-790).

The Rust host crate a maintainer would add does exactly the same through the same C ABI
(INTEGRATION.md); it cannot be compiled in this image (no cargo).
"""
from __future__ import annotations

import json
import math
import os
import threading
import time as _time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Iterable, Sequence

import numpy as np

from . import _abi
from ._abi import SimulationError  # noqa: F401  (re-export: chili/errors.rs:73-121)
from .device import DeviceSimulation, make_config


# ---- domain (cellular_raza-building-blocks/src/domains/cartesian_cuboid_n.rs) ---------------------
@dataclass
class CartesianCuboid:
    min: np.ndarray
    max: np.ndarray
    dx: np.ndarray
    n_voxels: np.ndarray
    rng_seed: int = 0

    @staticmethod
    def _check(mn, mx):
        if np.any(mn >= mx):  # check_min_max (:149-162)
            raise SimulationError(_abi.CRB_ERR_BOUNDARY, "Min must be smaller than Max for domain boundaries")

    @classmethod
    def from_boundaries_and_interaction_range(cls, min, max, interaction_range: float) -> "CartesianCuboid":
        """:183-222 — n = floor((max - min) / range) clamped to >= 1, dx = (max - min) / n."""
        mn, mx = np.asarray(min, dtype=np.float64), np.asarray(max, dtype=np.float64)
        cls._check(mn, mx)
        n = np.floor((mx - mn) / interaction_range)
        nv = np.maximum(n.astype(np.uint64), 1)
        with np.errstate(divide="ignore"):
            dx = (mx - mn) / n
        return cls(mn, mx, dx, nv)

    @classmethod
    def from_boundaries_and_n_voxels(cls, min, max, n_voxels) -> "CartesianCuboid":
        """:226-256"""
        mn, mx = np.asarray(min, dtype=np.float64), np.asarray(max, dtype=np.float64)
        cls._check(mn, mx)
        nv = np.asarray(n_voxels, dtype=np.uint64)
        return cls(mn, mx, (mx - mn) / nv.astype(np.float64), nv)

    @property
    def dim(self) -> int:
        return len(self.min)


# ---- mechanics (cell_building_blocks/mechanics.rs) -------------------------------------------------
@dataclass
class _Mechanics:
    pos: Sequence[float]
    KIND = -1

    def row(self) -> list[float]:
        raise NotImplementedError

    def velocity(self):
        return getattr(self, "vel", [0.0] * len(self.pos))


@dataclass
class NewtonDamped3D(_Mechanics):  # :158-173 (2-D twin below)
    vel: Sequence[float] = (0.0, 0.0, 0.0)
    damping_constant: float = 0.0
    mass: float = 1.0
    KIND = _abi.CRB_MECH_NEWTON_DAMPED

    def row(self):
        return [self.damping_constant, self.mass]


class NewtonDamped2D(NewtonDamped3D):
    pass


class NewtonDamped1D(NewtonDamped3D):
    pass


@dataclass
class Brownian3D(_Mechanics):  # :336-357
    diffusion_constant: float = 0.0
    kb_temperature: float = 1.0
    KIND = _abi.CRB_MECH_BROWNIAN

    def row(self):
        return [self.diffusion_constant, self.kb_temperature]


class Brownian2D(Brownian3D):
    pass


class Brownian1D(Brownian3D):  # used by the reference's tests/brownian_diffusion_constant_approx.rs:230-258
    pass


@dataclass
class Langevin3D(_Mechanics):  # :434-463
    vel: Sequence[float] = (0.0, 0.0, 0.0)
    mass: float = 1.0
    damping: float = 0.0
    kb_temperature: float = 0.0
    KIND = _abi.CRB_MECH_LANGEVIN

    def row(self):
        return [self.mass, self.damping, self.kb_temperature]


class Langevin2D(Langevin3D):
    pass


class Langevin1D(Langevin3D):
    pass


# ---- interaction (cell_building_blocks/interaction.rs; examples/cell_sorting/src/main.rs) ----------
@dataclass
class MorsePotential:  # :226-260
    radius: float
    potential_stiffness: float
    cutoff: float
    strength: float
    KIND = _abi.CRB_INT_MORSE
    species = 0

    def row(self):
        return [self.radius, self.potential_stiffness, self.cutoff, self.strength]


@dataclass
class BoundLennardJones:  # :122-138
    epsilon: float
    sigma: float
    bound: float
    cutoff: float
    KIND = _abi.CRB_INT_BOUND_LJ
    species = 0

    def row(self):
        return [self.epsilon, self.sigma, self.bound, self.cutoff]


@dataclass
class MiePotential:  # :345-389
    radius: float
    strength: float
    bound: float
    cutoff: float
    en: float
    em: float
    KIND = _abi.CRB_INT_MIE
    species = 0

    def row(self):
        return [self.radius, self.strength, self.bound, self.cutoff, self.en, self.em]


@dataclass
class CellSpecificInteraction:  # cellular_raza-examples/cell_sorting/src/main.rs:26-97
    species: int  # 0 = RedCell, 1 = BlueCell
    cell_radius: float
    potential_strength: float
    relative_interaction_range: float
    KIND = _abi.CRB_INT_CELL_SORTING

    def row(self):
        return [self.cell_radius, self.potential_strength, self.relative_interaction_range]


# ---- the F32 variants of the building blocks ---------------------------------------------------------
# mechanics.rs:202-204, 385-387, 569-571 (NewtonDamped{1,2,3}DF32, Brownian{1,2,3}DF32, Langevin{1,2,3}DF32) and
# interaction.rs:144-145, 262-263, 405 (BoundLennardJonesF32, MorsePotentialF32, MiePotentialF32) are the same
# structs over f32.  The device state of this backend is f64 structure-of-arrays throughout (BASELINE.json
# north_star) and its parity contract (forces 1e-12, trajectories 1e-9) cannot be stated for f32
# arithmetic; an f32 state would be a second layout and a second set of kernels, not a parameter.  The decision
# (INTEGRATION.md "F32 building blocks"): the names exist so that a ported model fails at the first call with a
# message that says what to do, instead of an AttributeError or a silent widening to f64.
class _F32Refused:
    F64_TWIN = ""

    def __init__(self, *args, **kwargs):
        raise SimulationError(
            _abi.CRB_ERR_INVALID_ARG,
            f"{type(self).__name__}: the B200 backend keeps all agent state in f64 and does not run the F32 variants of "
            f"the building blocks (cellular_raza-building-blocks mechanics.rs:202-204, 385-387, 569-571; "
            f"interaction.rs:144-145, 262-263, 405); use {self.F64_TWIN} (results within f32 rounding of the F32 "
            "variant) or the chili backend")


def _f32(name: str, twin: str):
    return type(name, (_F32Refused,), {"F64_TWIN": twin, "__doc__": f"refused: see {twin} (f64)"})


for _d in (1, 2, 3):
    for _k in ("NewtonDamped", "Brownian", "Langevin"):
        globals()[f"{_k}{_d}DF32"] = _f32(f"{_k}{_d}DF32", f"{_k}{_d}D")
for _k in ("MorsePotential", "BoundLennardJones", "MiePotential"):
    globals()[f"{_k}F32"] = _f32(f"{_k}F32", _k)
del _d, _k


@dataclass
class Agent:
    """what ``#[derive(CellAgent)] struct Cell { #[Mechanics] mechanics, #[Interaction] interaction }``
    is in the reference (cell_sorting/src/main.rs:99-105)."""

    mechanics: _Mechanics
    interaction: object = None  # None: an agent without the Interaction aspect (domain_force_spring.rs:70-75)


# ---- bacterial rods (cell_building_blocks/bacterial_rods.rs; examples/bacterial_rods/src/main.rs) -----
MICRO_METRE = 1e-6
MINUTE = 1.0
GROWTH_BASE_RATE = 0.1 * MICRO_METRE / MINUTE  # examples/bacterial_rods/src/main.rs:12


@dataclass
class RodMechanics:  # bacterial_rods.rs:82-107
    pos: Sequence[Sequence[float]]  # nrows x 3 (one row per vertex, like nalgebra::MatrixXx3)
    vel: Sequence[Sequence[float]]
    diffusion_constant: float = 0.0
    spring_tension: float = 0.0
    rigidity: float = 0.0
    spring_length: float = 0.0
    damping: float = 0.0
    KIND = _abi.CRB_MECH_ROD


@dataclass
class RodInteraction:  # bacterial_rods.rs:272-290: newtype around a point potential
    potential: MorsePotential
    KIND = _abi.CRB_INT_MORSE

    def row(self):
        return self.potential.row()


@dataclass
class RodAgent:
    """``Agent`` of cellular_raza-examples/bacterial_rods/src/main.rs:14-109: RodMechanics +
    RodInteraction<MorsePotential> + the Cycle / NeighborSensing fields and rules of the example
    (growth_rate <- GROWTH_BASE_RATE * max(8 - neighbours, 0) / 8, division above the threshold)."""

    mechanics: RodMechanics
    interaction: RodInteraction
    growth_rate: float = GROWTH_BASE_RATE
    spring_length_threshold: float = 6.0 * MICRO_METRE
    growth_base_rate: float = GROWTH_BASE_RATE

    def mech_row(self):  # CRB_MECH_ROD row of include/crb.h
        m = self.mechanics
        return [m.diffusion_constant, m.spring_tension, m.rigidity, m.spring_length, m.damping,
                self.growth_rate, self.spring_length_threshold, self.growth_base_rate]


@dataclass
class CartesianCuboidRods:  # bacterial_rods.rs:338-411
    domain: CartesianCuboid
    gel_pressure: float = 0.0
    surface_friction: float = 0.0
    surface_friction_distance: float = 0.0  # create_subdomains passes surface_friction instead (:406)

    @property
    def dim(self) -> int:
        return self.domain.dim


@dataclass
class SpringDomainForce:
    """``SubDomainForce::calculate_custom_force`` of cellular_raza/tests/domain_force_spring.rs:56-68, one
    strength and anchor per axis: ``force[c] = -strength[c] * (pos[c] - anchor[c])``.  The reference lets
    the user write any force law; the device offers this closed form (include/crb.h crb_domain_force_kind)."""

    strength: Sequence[float]
    anchor: Sequence[float] | None = None


# ---- time (cellular_raza-core/src/time.rs) ---------------------------------------------------------
def _round_to_usize(x: float) -> int:
    """``F::round`` (half AWAY from zero, unlike Python's ``round``) followed by ``to_usize`` (time.rs:203-219)"""
    r = math.floor(abs(x) + 0.5)
    if x < 0.0 and r > 0:
        raise ValueError("An error in casting of float type to usize occurred")
    return int(r)


@dataclass
class FixedStepsize:
    t0: float
    dt: float
    save_iterations: list[int]
    maximum_iterations: int

    @classmethod
    def from_partial_save_points(cls, t0: float, dt: float, points: Iterable[float]) -> "FixedStepsize":
        """:173-239 — iteration of a save point = round((t - t0) / dt); the run ends at the last one."""
        pts = sorted(points)
        if not pts:
            raise ValueError("No savepoints specified. Simulation will not save any results.")
        if any(t0 > p for p in pts):
            raise ValueError("Invalid time configuration! Evaluation time point is before starting time point.")
        its = [_round_to_usize((p - t0) / dt) for p in pts]
        return cls(t0, dt, its, its[-1])

    @classmethod
    def from_partial_save_freq(cls, t0: float, dt: float, t_max: float, save_freq: int) -> "FixedStepsize":
        """:135-168 as written: ``maximum_iterations = round((t_max - t0) / dt)`` and one PartialSave event at every
        iteration n of ``(0..maximum_iterations).step_by(save_freq)`` — the end point itself is NOT a save event
        (exclusive range), so the run ends ``maximum_iterations`` iterations in without storing its last state unless
        ``maximum_iterations - 1`` is a multiple of ``save_freq``."""
        if save_freq <= 0:
            raise ValueError("save_freq must be positive")  # (step_by(0) panics)
        mx = _round_to_usize((t_max - t0) / dt)
        return cls(t0, dt, list(range(0, mx, save_freq)), mx)

    @classmethod
    def from_partial_save_steps(cls, t0: float, dt: float, n_steps: int, save_interval: int) -> "FixedStepsize":
        """:94-111 — save points t0 + k * save_interval * dt for k = 0 ..= ceil(n_steps / save_interval)."""
        k = -(-n_steps // save_interval)
        return cls.from_partial_save_points(t0, dt, [t0 + (save_interval * n) * dt for n in range(k + 1)])

    @classmethod
    def from_partial_save_interval(cls, t0: float, dt: float, t_max: float, save_interval: float) -> "FixedStepsize":
        pts, t = [], t0
        while t <= t_max:
            pts.append(t)
            t += save_interval
        return cls.from_partial_save_points(t0, dt, pts)


# ---- storage (cellular_raza-core/src/storage/concepts.rs) ------------------------------------------
@dataclass
class StorageBuilder:
    """``StorageBuilder::new().priority([...]).location(path)`` (:226-331).  priority entries:
    "Memory" and/or "SerdeJson"."""

    location: str = "out"
    priority: Sequence[str] = ("Memory",)
    add_date: bool = True
    suffix: str = ""


class StorageManager:
    """read side of ``StorageAccess.cells`` (``StorageInterfaceLoad``): iteration -> elements."""

    def __init__(self, json_dir: Path | None):
        self.json_dir = json_dir
        self.memory: dict[int, dict] = {}

    def get_all_iterations(self) -> list[int]:
        its = set(self.memory)
        if self.json_dir is not None and self.json_dir.exists():
            its |= {int(p.name) for p in self.json_dir.iterdir() if p.is_dir()}
        return sorted(its)

    @staticmethod
    def _key(identifier) -> tuple:
        """serde form of a ``CellIdentifier`` (``{"Initial": n}`` / ``{"Division": [voxel, counter]}``, cell.rs:18-25)
        or a tuple -> the tuple the Memory storage uses: ("Initial", n) / ("Division", voxel, counter)"""
        if isinstance(identifier, dict):
            (kind, val), = identifier.items()
            return (kind, *val) if isinstance(val, (list, tuple)) else (kind, val)
        return tuple(identifier) if isinstance(identifier, (list, tuple)) else identifier

    def load_all_elements_at_iteration(self, iteration: int) -> dict:
        """-> {identifier tuple: {"pos", "vel", "id", "voxel", "aux"}} (:937-944)"""
        if iteration in self.memory:
            return self.memory[iteration]
        out = {}
        if self.json_dir is None:
            return out
        for f in sorted((self.json_dir / f"{iteration:020d}").glob("batch_*.json")):
            for e in json.loads(f.read_text())["data"]:
                cell, aux = e["element"]
                out[self._key(e["identifier"])] = {"cell": cell["cell"], "aux": aux}
        return out

    def load_all_elements(self) -> dict[int, dict]:
        """:951-964 (a BTreeMap: ascending iterations)"""
        return {it: self.load_all_elements_at_iteration(it) for it in self.get_all_iterations()}

    def load_single_element(self, iteration: int, identifier):
        """:863-871: the element, or None when the identifier is not stored at that iteration"""
        return self.load_all_elements_at_iteration(iteration).get(self._key(identifier))

    def load_element_history(self, identifier) -> dict[int, dict]:
        """:899-931: every occurrence of the element in ascending iterations; like the reference, gathering stops at
        the first iteration that no longer holds the identifier once it has been seen (an element lives over one
        uninterrupted span)"""
        out, started = {}, False
        for it in self.get_all_iterations():
            e = self.load_single_element(it, identifier)
            if e is not None:
                started = True
                out[it] = e
            elif started:
                break
        return out

    def load_all_element_histories(self) -> dict[tuple, dict[int, dict]]:
        """:969-994: identifier -> {iteration: element}"""
        out: dict[tuple, dict[int, dict]] = {}
        for it, elements in self.load_all_elements().items():
            for ident, e in elements.items():
                out.setdefault(ident, {})[it] = e
        return out


@dataclass
class StorageAccess:  # chili/result.rs:5-10
    cells: StorageManager
    subdomains: StorageManager
    path: Path | None = None


@dataclass
class Settings:  # chili/setup.rs:17-27
    n_threads: int
    time: FixedStepsize
    storage: StorageBuilder = field(default_factory=StorageBuilder)
    progressbar: str | None = None


def _storage_names(agents, domain, n_threads: int) -> dict:
    """what the snapshot writers need besides the state: the agents' struct types (serde field names) and the map
    from a plain voxel index to its chili subdomain (= storage instance of the batch file)"""
    subs = chili_subdomains(domain.min, domain.max, domain.n_voxels, domain.dx, n_threads)
    firsts = np.cumsum([0] + [len(s["voxels"]) for s in subs])[:-1]
    inter = agents[0].interaction if agents else None
    return {"mech": type(agents[0].mechanics).__name__ if agents else "", "inter": type(inter).__name__,
            "mech_cls": type(agents[0].mechanics) if agents else None, "inter_cls": None if inter is None else type(inter),
            "subdomains": subs,
            "sub_of": (lambda vox: np.searchsorted(firsts, np.asarray(vox, dtype=np.int64), side="right") - 1)}


# ---- the entry point -------------------------------------------------------------------------------
def _pack(agents: Sequence[Agent], dim: int):
    n = len(agents)
    if n == 0:
        raise ValueError("no agents")
    mk = {type(a.mechanics).KIND for a in agents}
    ik = {(_abi.CRB_INT_NONE if a.interaction is None else type(a.interaction).KIND) for a in agents}
    if len(mk) != 1 or len(ik) != 1:
        raise SimulationError(_abi.CRB_ERR_INVALID_ARG, "all agents must use the same mechanics and interaction types")
    pos = np.array([a.mechanics.pos for a in agents], dtype=np.float64).reshape(n, dim)
    vel = np.array([a.mechanics.velocity() for a in agents], dtype=np.float64).reshape(n, dim)
    mp = np.array([a.mechanics.row() for a in agents], dtype=np.float64)
    ip = np.array([([] if a.interaction is None else a.interaction.row()) for a in agents], dtype=np.float64).reshape(n, -1)
    sp = np.array([getattr(a.interaction, "species", 0) for a in agents], dtype=np.uint8)
    return mk.pop(), ik.pop(), pos, vel, mp, ip, sp


def run_simulation(domain: CartesianCuboid, agents: Sequence[Agent], settings: Settings,
                   aspects: Sequence[str] = ("Mechanics", "Interaction"), *,
                   mechanics_solver_order: int = 2, determinism: bool = True,
                   pair_weight_mode: int = _abi.CRB_WEIGHT_CHILI, noise_mode: int | None = None,
                   noise_table: np.ndarray | None = None, device: int = 0,
                   domain_force: SpringDomainForce | None = None, capacity: int | None = None,
                   devices: Sequence[int] | None = None) -> StorageAccess:
    """``run_simulation!`` of the B200 backend.  ``settings.n_threads`` is chili's subdomain count: it
    does not create threads here, it selects the pair weights chili applies for that decomposition
    (SURVEY.md finding 2); ``pair_weight_mode=CRB_WEIGHT_PHYSICAL`` applies both halves everywhere.
    ``devices=[0, 1, ...]`` cuts the domain into one slab of voxel planes per listed GPU (one
    host thread drives all of them, halo and migration buffers travel by peer copies); the results are
    bit-identical to the single-device run.  A device may be listed several times (several slabs on it)."""
    agents = list(agents)
    if agents and isinstance(agents[0], RodAgent):
        return _run_rods(domain, agents, settings, aspects, mechanics_solver_order=mechanics_solver_order,
                         pair_weight_mode=pair_weight_mode, device=device, capacity=capacity, devices=devices)
    unsupported = set(aspects) - {"Mechanics", "Interaction", "DomainForce"}
    if unsupported or "Mechanics" not in aspects:
        raise SimulationError(_abi.CRB_ERR_INVALID_ARG, f"aspects {sorted(unsupported)} are outside the hot path (SURVEY.md §8)")
    if ("DomainForce" in aspects) != (domain_force is not None):
        raise SimulationError(_abi.CRB_ERR_INVALID_ARG, "aspect DomainForce needs domain_force=SpringDomainForce(...) and vice versa")
    dim = domain.dim
    if "Interaction" not in aspects:
        agents = [Agent(a.mechanics, None) for a in agents]
    mk, ik, pos, vel, mp, ip, sp = _pack(agents, dim)
    if ("Interaction" in aspects) == (ik == _abi.CRB_INT_NONE):
        raise SimulationError(_abi.CRB_ERR_INVALID_ARG, "aspect Interaction needs agents with an interaction")
    if noise_mode is None:
        noise_mode = _abi.CRB_NOISE_NONE if mk == _abi.CRB_MECH_NEWTON_DAMPED else (
            _abi.CRB_NOISE_TABLE if noise_table is not None else _abi.CRB_NOISE_PHILOX)
    cfg = make_config(dim=dim, mechanics_kind=mk, interaction_kind=ik, domain_min=domain.min, domain_max=domain.max,
                      n_voxels=domain.n_voxels, dx=domain.dx, dt=settings.time.dt, capacity=len(agents),
                      solver_order=mechanics_solver_order, pair_weight_mode=pair_weight_mode,
                      ref_n_subdomains=max(settings.n_threads, 1), noise_mode=noise_mode,
                      rng_seed=domain.rng_seed, device=device,
                      domain_force_kind=_abi.CRB_DOMAIN_FORCE_SPRING if domain_force is not None else _abi.CRB_DOMAIN_FORCE_NONE,
                      domain_force_strength=None if domain_force is None else domain_force.strength,
                      domain_force_anchor=None if domain_force is None else domain_force.anchor)
    if devices is not None and len(devices) > 1:
        if domain_force is not None:
            raise SimulationError(_abi.CRB_ERR_INVALID_ARG, "devices=[...] with DomainForce: use one device")
        return _run_slabs(cfg, list(devices), agents, settings, pos, vel, mp, ip, sp, noise_table, domain)
    sim = DeviceSimulation(cfg)
    try:
        sim.set_agents(pos, vel, mp, ip, sp)
        if noise_table is not None:
            sim.set_noise(noise_table)
        sb = settings.storage
        root = None
        if "SerdeJson" in sb.priority:
            root = Path(sb.location)
            if sb.add_date:
                root = root / _time.strftime("%Y-%m-%d-T%H-%M-%S")
            if sb.suffix:
                root = root / sb.suffix
        cells = StorageManager(None if root is None else root / "cells" / "json")
        subdomains = StorageManager(None if root is None else root / "subdomains" / "json")
        ts = settings.time
        save_its = sorted(set(ts.save_iterations))
        agent_rows = _storage_names(agents, domain, max(settings.n_threads, 1))
        # Snapshots do not stop the stepping loop: crb_snapshot_async packs the state on the device and
        # copies it to pinned host memory on a side stream while the next steps run; the JSON / memory
        # store is written by a writer thread (ctypes releases the GIL inside crb_step).
        # mechanics / interaction parameters do not change on this path: fetched once, looked up by
        # identifier for every snapshot (the device order changes with every cell-list rebuild)
        ids0 = sim.get_agents()[2]
        o0 = np.argsort(ids0, kind="stable")
        ids_sorted = ids0[o0]
        params_sorted = tuple(a[o0] for a in sim.get_params())
        writer: threading.Thread | None = None
        pending = None  # (iteration, snapshot token)
        errors: list[BaseException] = []

        def flush():
            nonlocal writer, pending
            if writer is not None:
                writer.join()
                writer = None
            if errors:
                raise errors[0]
            if pending is not None:
                it_p, token = pending
                pending = None
                state = sim.snapshot_wait(token)

                def work():
                    try:
                        k = np.searchsorted(ids_sorted, state[2])
                        _store(state, tuple(a[k] for a in params_sorted), cells, it_p, sb, agent_rows)
                    except BaseException as e:  # surfaced by the next flush()
                        errors.append(e)

                writer = threading.Thread(target=work, name=f"crb-save-{it_p}")
                writer.start()

        def save(iteration):
            nonlocal pending
            flush()  # one snapshot in flight: finish the previous one first
            pending = (iteration, sim.snapshot_async(aux=True))
            _store_subdomains(subdomains, iteration, sb, agent_rows["subdomains"])  # save_subdomains (run_sim.rs:322-325)

        # save_initial (time.rs:269-280, run_sim.rs:375-378): current_time == t0 when the loop starts, so the
        # initial state is ALWAYS stored, whether or not t0 is among the save points
        save(0)
        it = 0
        for target in [s for s in save_its if s > 0]:
            sim.step(target - it)  # replaces (target - it) iterations of run_sim.rs:380-403
            it = target
            save(it)
        if it < ts.maximum_iterations:
            sim.step(ts.maximum_iterations - it)
        flush()
        if writer is not None:
            writer.join()
        if errors:
            raise errors[0]
        return StorageAccess(cells, subdomains, root)
    finally:
        sim.close()


def _run_slabs(cfg, devices, agents, settings: Settings, pos, vel, mp, ip, sp, noise_table, domain) -> StorageAccess:
    """one slab per entry of ``devices`` driven by this host thread (``slabs.LocalSlabGroup``); snapshots are
    gathered slab by slab at the save points."""
    import copy

    from .slabs import LocalSlabGroup

    cfg.capacity = max(len(agents), 1)  # every slab can hold the whole population (clustering over time)
    grp = LocalSlabGroup(lambda: copy.copy(cfg), len(devices), devices=devices)
    try:
        grp.set_agents(pos, vel, mp, ip, sp)
        if noise_table is not None:
            grp.set_noise(noise_table)
        sb = settings.storage
        root = None
        if "SerdeJson" in sb.priority:
            root = Path(sb.location)
            if sb.add_date:
                root = root / _time.strftime("%Y-%m-%d-T%H-%M-%S")
            if sb.suffix:
                root = root / sb.suffix
        cells = StorageManager(None if root is None else root / "cells" / "json")
        subdomains = StorageManager(None if root is None else root / "subdomains" / "json")
        names = _storage_names(agents, domain, max(settings.n_threads, 1))
        ts = settings.time
        save_its = sorted(set(ts.save_iterations))

        def save(iteration):
            state = (*grp.get_agents(), *grp.get_aux())
            k = state[2].astype(np.int64)  # identifiers are the input indices (CellIdentifier::Initial(n))
            _store(state, (mp[k], ip[k], sp[k]), cells, iteration, sb, names)
            _store_subdomains(subdomains, iteration, sb, names["subdomains"])

        save(0)  # save_initial: always (time.rs:269-280, run_sim.rs:375-378)
        it = 0
        for target in [x for x in save_its if x > 0]:
            grp.step(target - it)
            it = target
            save(it)
        if it < ts.maximum_iterations:
            grp.step(ts.maximum_iterations - it)
        return StorageAccess(cells, subdomains, root)
    finally:
        grp.close()


def _save(sim: DeviceSimulation, cells: StorageManager, iteration: int, sb: StorageBuilder, names: dict):
    """blocking variant: download and store (used by tests)"""
    state = (*sim.get_agents(), *sim.get_aux())
    _store(state, sim.get_params(), cells, iteration, sb, names)


_SPECIES_NAMES = {0: "RedCell", 1: "BlueCell"}  # enum Species of cell_sorting/src/main.rs:25-29 (serde unit variants)


def _serde_fields(cls, skip=("pos", "vel", "species")):
    """field names of a building-block struct in declaration order = the keys its serde derive writes
    (mechanics.rs:54-65, 253-261, 413-426; interaction.rs:87-96, 215-224, 322-335; the dataclasses above declare the
    same fields in the same order, and their row() methods list the scalar ones in that order)"""
    import dataclasses

    return [f.name for f in dataclasses.fields(cls) if f.name not in skip]


def _cell_body(names: dict, pos, vel, mp_row, ip_row, species) -> dict:
    """the user's ``#[derive(CellAgent, Serialize)] struct Cell { interaction, mechanics }``
    (cell_sorting/src/main.rs:99-105) with the building-block structs' own serde field names:
    ``mechanics: {pos, vel, damping_constant, mass}``, ``interaction: {species: "BlueCell" | "RedCell", cell_radius,
    potential_strength, relative_interaction_range}`` — what cell_sorting/src/plot.py:43-57 reads."""
    mcls, icls = names["mech_cls"], names.get("inter_cls")
    mech = {"pos": [float(x) for x in pos]}
    if "vel" in {f for f in _serde_fields(mcls, skip=())}:
        mech["vel"] = [float(x) for x in vel]
    for k, v in zip(_serde_fields(mcls), mp_row):
        mech[k] = float(v)
    cell = {}
    if icls is not None:
        inter = {}
        if icls is CellSpecificInteraction:
            inter["species"] = _SPECIES_NAMES.get(int(species), int(species))
        for k, v in zip(_serde_fields(icls), ip_row):
            inter[k] = float(v)
        cell["interaction"] = inter
    cell["mechanics"] = mech
    return cell


def chili_subdomains(domain_min, domain_max, n_voxels, dx, n_threads: int) -> list[dict]:
    """``CartesianSubDomain`` of every chili subdomain as ``create_subdomains`` builds it
    (cartesian_cuboid_n.rs:415-423, 500-583): contiguous chunks of the voxel enumeration (last axis fastest), bounding
    box of the chunk as min / max; serde field names of the struct."""
    from itertools import product

    nv = [int(x) for x in n_voxels]
    dim = len(nv)
    lib = _abi.load_library()
    import ctypes as C

    n, m, avg = C.c_uint64(), C.c_uint64(), C.c_uint64()
    total = int(np.prod(nv))
    if lib.crb_decomp_res(total, max(int(n_threads), 1), C.byref(n), C.byref(m), C.byref(avg)) != 0:
        raise SimulationError(_abi.CRB_ERR_DECOMPOSE, "DecomposeError: Could not find a suiting decomposition")
    sizes = [int(avg.value)] * int(n.value) + [int(avg.value) - 1] * int(m.value)
    sizes = [x for x in sizes if x > 0]
    voxels = list(product(*[range(k) for k in nv]))
    out, first = [], 0
    for size in sizes:
        chunk = voxels[first:first + size]
        first += size
        lo = [min(v[i] for v in chunk) for i in range(dim)]
        hi = [max(v[i] for v in chunk) for i in range(dim)]
        out.append({
            "min": [float(domain_min[i] + lo[i] * dx[i]) for i in range(dim)],
            "max": [float(domain_min[i] + (hi[i] + 1.0) * dx[i]) for i in range(dim)],
            "dx": [float(x) for x in dx], "voxels": [list(v) for v in chunk],
            "domain_min": [float(x) for x in domain_min], "domain_max": [float(x) for x in domain_max],
            "domain_n_voxels": nv,
        })
    return out


def _store_subdomains(subdomains: "StorageManager", iteration: int, sb: "StorageBuilder", subs: list[dict], wrap=None):
    """save_subdomains (chili/datastructures.rs:349-371): one ``single_<instance>_<k>.json`` per subdomain holding
    ``{"identifier": SubDomainPlainIndex, "element": SubDomain}`` (storage/concepts.rs:156-162, 702-743)"""
    if "Memory" in sb.priority:
        subdomains.memory[iteration] = {i: (wrap(s) if wrap else s) for i, s in enumerate(subs)}
    if subdomains.json_dir is None:
        return
    d = subdomains.json_dir / f"{iteration:020d}"
    d.mkdir(parents=True, exist_ok=True)
    for i, sub in enumerate(subs):
        k = 0
        while (d / f"single_{i:020d}_{k:020d}.json").exists():
            k += 1
        (d / f"single_{i:020d}_{k:020d}.json").write_text(
            json.dumps({"identifier": i, "element": wrap(sub) if wrap else sub}, indent=2))


def _store(state, params, cells: StorageManager, iteration: int, sb: StorageBuilder, names: dict):
    """save_cells (chili/datastructures.rs:373-396) -> store_batch_elements (storage/concepts.rs:766-790): one batch
    file per storage instance = chili subdomain (``names["sub_of"]`` maps a plain voxel index to it)."""
    pos, vel, ids, vox, hx, hv, hl = state
    mp, ip, sp = params
    if "Memory" in sb.priority:
        cells.memory[iteration] = {
            ("Initial", int(i)): {"pos": pos[k], "vel": vel[k], "voxel": int(vox[k]),
                                  "aux": {"positions": hx[k, : hl[k]], "velocities": hv[k, : hl[k]]}}
            for k, i in enumerate(ids)
        }
    if cells.json_dir is not None:
        d = cells.json_dir / f"{iteration:020d}"
        d.mkdir(parents=True, exist_ok=True)
        dim = pos.shape[1]
        sub_of = names.get("sub_of")
        inst = np.zeros(len(ids), dtype=np.int64) if sub_of is None else sub_of(vox)
        for instance in np.unique(inst) if len(ids) else [0]:
            data = []
            for a in np.nonzero(inst == instance)[0]:
                ident = {"Initial": int(ids[a])}
                data.append({
                    "identifier": ident,
                    "element": [
                        {"identifier": ident, "parent": None,
                         "cell": _cell_body(names, pos[a], vel[a], mp[a], ip[a], sp[a])},
                        # _CrAuxStorage of aspects [Mechanics, Interaction]: only Mechanics contributes a field
                        # (cellular_raza-core-proc-macro/src/aux_storage.rs:807-823); RingBuffer = circ_buffer 0.1.8
                        # (un-vendored): written as the sequence of its entries, oldest first
                        {"mechanics": {"positions": hx[a, : hl[a]].tolist(), "velocities": hv[a, : hl[a]].tolist(),
                                       "current_force": [0.0] * dim, "zero_force": [0.0] * dim}},
                    ],
                })
            k = 0
            while (d / f"batch_{int(instance):020d}_{k:020d}.json").exists():
                k += 1
            (d / f"batch_{int(instance):020d}_{k:020d}.json").write_text(json.dumps({"data": data}, indent=2))


# ---- bacterial rods: aspects [Mechanics, Interaction, Cycle, NeighborSensing] (+ DomainForce) ------------
_IDENT_NAMES = {0: "Division", 1: "Initial", 2: "Inserted"}  # CellIdentifier, concepts/src/cell.rs:18-25


def _ident(t):
    kind, a, b = int(t[0]), int(t[1]), int(t[2])
    if kind == 1:
        return ("Initial", a)
    return (_IDENT_NAMES[kind], a, b)


def _ident_json(key):
    return {key[0]: key[1]} if key[0] == "Initial" else {key[0]: [key[1], key[2]]}


def _matrix_json(m: np.ndarray):
    """nalgebra 0.34 serde of Matrix<f64, Dyn, Const<3>, VecStorage>: the storage as the tuple
    (data column-major, nrows, ncols) with the Const dimension as unit (un-vendored dependency; layout
    restated from its published source, not pinned by any reference fixture)."""
    return [np.asarray(m).T.reshape(-1).tolist(), int(m.shape[0]), None]


def _run_rods(domain, agents: Sequence[RodAgent], settings: Settings, aspects, *, mechanics_solver_order: int,
              pair_weight_mode: int, device: int, capacity: int | None = None,
              devices: Sequence[int] | None = None) -> StorageAccess:
    """``run_simulation!(domain, agents, settings, aspects: [Mechanics, Interaction, Cycle, NeighborSensing],
    zero_force_default: |c| MatrixXx3::zeros(nrows))`` of examples/bacterial_rods/src/main.rs:202-210.
    The rod count changes when rods divide (``CycleEvent::Division``), identifiers follow
    chili/update_cycle.rs:33-66; snapshots are taken after the step of every save point."""
    from .rods import RodSimulation

    need = {"Mechanics", "Interaction", "Cycle", "NeighborSensing"}
    aspects = set(aspects)
    if not need <= aspects or aspects - need - {"DomainForce"}:
        raise SimulationError(_abi.CRB_ERR_INVALID_ARG,
                              "rod agents run aspects [Mechanics, Interaction, Cycle, NeighborSensing] (+ DomainForce)")
    gel = domain if isinstance(domain, CartesianCuboidRods) else None
    cub = gel.domain if gel is not None else domain
    if ("DomainForce" in aspects) and gel is None:
        raise SimulationError(_abi.CRB_ERR_INVALID_ARG, "aspect DomainForce needs a CartesianCuboidRods domain")
    if cub.dim != 3:
        raise SimulationError(_abi.CRB_ERR_INVALID_ARG, "RodMechanics<f64, 3> needs a 3-D domain")
    n = len(agents)
    pos = np.array([a.mechanics.pos for a in agents], dtype=np.float64)
    vel = np.array([a.mechanics.vel for a in agents], dtype=np.float64)
    if pos.ndim != 3 or pos.shape[2] != 3 or vel.shape != pos.shape:
        raise SimulationError(_abi.CRB_ERR_INVALID_ARG, "all rods need the same nrows x 3 position/velocity matrices")
    V = pos.shape[1]
    mp = np.array([a.mech_row() for a in agents], dtype=np.float64)
    ip = np.array([a.interaction.row() for a in agents], dtype=np.float64)
    ts = settings.time
    save_its = sorted(set(ts.save_iterations))
    # device capacity (the reference grows its Vecs; the device needs a bound): head room for three
    # doublings unless the caller passes `capacity`; exceeding it is a CapacityError, never a silent drop
    capacity = int(capacity or max(8 * n, 1024))
    use_force = "DomainForce" in aspects
    cfg = make_config(dim=3, mechanics_kind=_abi.CRB_MECH_ROD, interaction_kind=_abi.CRB_INT_MORSE,
                      domain_min=cub.min, domain_max=cub.max, n_voxels=cub.n_voxels, dx=cub.dx, dt=ts.dt,
                      capacity=capacity, solver_order=mechanics_solver_order, pair_weight_mode=pair_weight_mode,
                      ref_n_subdomains=max(settings.n_threads, 1), rng_seed=cub.rng_seed, device=device,
                      rod_n_vertices=V,
                      domain_force_kind=_abi.CRB_DOMAIN_FORCE_ROD_GEL if use_force else _abi.CRB_DOMAIN_FORCE_NONE,
                      # create_subdomains hands surface_friction to surface_friction_distance (bacterial_rods.rs:406)
                      domain_force_strength=[gel.gel_pressure, gel.surface_friction, gel.surface_friction] if use_force else None)
    rod_subs = chili_subdomains(cub.min, cub.max, cub.n_voxels, cub.dx, max(settings.n_threads, 1))
    _firsts = np.cumsum([0] + [len(x["voxels"]) for x in rod_subs])[:-1]

    def rod_sub_of(vox):
        return np.searchsorted(_firsts, np.asarray(vox, dtype=np.int64), side="right") - 1

    if devices is not None and len(devices) > 1:
        # one slab of voxel planes per listed GPU, driven by this host thread; the gathered views come back in the
        # order of the single-device run (rank order = ascending voxel), so the snapshots are the same bit for bit
        import copy

        from .slabs import LocalRodSlabGroup

        sim = LocalRodSlabGroup(lambda: copy.copy(cfg), len(devices), devices=list(devices))
    else:
        sim = RodSimulation(cfg)
    try:
        sim.set_agents(pos, vel, mp, ip)
        sb = settings.storage
        root = None
        if "SerdeJson" in sb.priority:
            root = Path(sb.location)
            if sb.add_date:
                root = root / _time.strftime("%Y-%m-%d-T%H-%M-%S")
            if sb.suffix:
                root = root / sb.suffix
        cells = StorageManager(None if root is None else root / "cells" / "json")
        subdomains = StorageManager(None if root is None else root / "subdomains" / "json")
        inter_row = ip[0].tolist()

        def save(iteration):
            p, v, ident, parent, vox = sim.get_agents()
            m = sim.get_params()
            hx, hv, hl = sim.get_aux()
            keys = [_ident(t) for t in ident]
            if "Memory" in sb.priority:
                cells.memory[iteration] = {
                    k: {"pos": p[a], "vel": v[a], "voxel": int(vox[a]), "spring_length": float(m[a, 3]),
                        "growth_rate": float(m[a, 5]),
                        "parent": None if int(parent[a, 0]) > 2 else _ident(parent[a]),
                        "aux": {"positions": hx[a, : hl[a]], "velocities": hv[a, : hl[a]]}}
                    for a, k in enumerate(keys)
                }
            if cells.json_dir is not None:
                d = cells.json_dir / f"{iteration:020d}"
                d.mkdir(parents=True, exist_ok=True)
                data = []
                zero = _matrix_json(np.zeros((V, 3)))
                for a, k in enumerate(keys):
                    ident_j = _ident_json(k)
                    par = None if int(parent[a, 0]) > 2 else _ident_json(_ident(parent[a]))
                    cell = {"mechanics": {"pos": _matrix_json(p[a]), "vel": _matrix_json(v[a]),
                                          "diffusion_constant": float(m[a, 0]), "spring_tension": float(m[a, 1]),
                                          "rigidity": float(m[a, 2]), "spring_length": float(m[a, 3]),
                                          "damping": float(m[a, 4])},
                            "interaction": {"radius": inter_row[0], "potential_stiffness": inter_row[1],
                                            "cutoff": inter_row[2], "strength": inter_row[3]},
                            "growth_rate": float(m[a, 5]), "spring_length_threshold": float(m[a, 6])}
                    aux = {"cycle": {"cycle_events": []},
                           "mechanics": {"positions": [_matrix_json(x) for x in hx[a, : hl[a]]],
                                         "velocities": [_matrix_json(x) for x in hv[a, : hl[a]]],
                                         "current_force": zero, "zero_force": zero},
                           "neighbor_sensing": {"accumulator": 0}}
                    data.append({"identifier": ident_j,
                                 "element": [{"identifier": ident_j, "parent": par, "cell": cell}, aux]})
                # one batch file per storage instance = chili subdomain of the rod's voxel
                inst = rod_sub_of(vox) if len(keys) else np.zeros(0, dtype=np.int64)
                for instance in (np.unique(inst) if len(keys) else [0]):
                    part = [data[a] for a in np.nonzero(inst == instance)[0]] if len(keys) else []
                    (d / f"batch_{int(instance):020d}_{0:020d}.json").write_text(json.dumps({"data": part}, indent=2))
            # save_subdomains: CartesianSubDomainRods { subdomain, gel_pressure, surface_friction,
            # surface_friction_distance } (bacterial_rods.rs:425-435) or the plain CartesianSubDomain
            _store_subdomains(subdomains, iteration, sb, rod_subs, wrap=(
                (lambda sub: {"subdomain": sub, "gel_pressure": float(gel.gel_pressure),
                              "surface_friction": float(gel.surface_friction),
                              # create_subdomains hands surface_friction to the distance field (:406)
                              "surface_friction_distance": float(gel.surface_friction)}) if gel is not None else None))

        save(0)  # save_initial: always (time.rs:269-280, run_sim.rs:375-378)
        it = 0
        for target in [x for x in save_its if x > 0]:
            sim.step(target - it)
            it = target
            save(it)
        if it < ts.maximum_iterations:
            sim.step(ts.maximum_iterations - it)
        return StorageAccess(cells, subdomains, root)
    finally:
        sim.close()
