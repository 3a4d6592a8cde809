"""Synthetic populations of the BASELINE.json configurations.

The reference seeds its examples with ChaCha8 (cellular_raza-examples/cell_sorting/src/main.rs:113-124),
a stream that cannot be reproduced here; positions therefore come from numpy's PCG64 with a fixed
seed and are handed to the oracle and to the device alike.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _abi
from .device import make_config


@dataclass
class Workload:
    name: str
    cfg_kwargs: dict
    pos: np.ndarray
    vel: np.ndarray
    mech_params: np.ndarray
    inter_params: np.ndarray
    species: np.ndarray | None = None
    noise_steps: int = 0
    extra: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return self.pos.shape[0]

    def config(self, **overrides):
        kw = dict(self.cfg_kwargs)
        kw.update(overrides)
        kw.setdefault("capacity", max(self.n, 1))
        return make_config(**kw)


def _domain_from_interaction_range(mn, mx, rng):
    """CartesianCuboid::from_boundaries_and_interaction_range (cartesian_cuboid_n.rs:183-222)."""
    mn = np.asarray(mn, dtype=np.float64)
    mx = np.asarray(mx, dtype=np.float64)
    n = np.floor((mx - mn) / rng)
    nv = np.maximum(n.astype(np.uint64), 1)
    with np.errstate(divide="ignore"):
        dx = (mx - mn) / n
    return nv, dx


def _uniform_positions(rng, n, dim, lo, hi):
    pos = rng.uniform(lo, hi, size=(n, dim))
    # check_min_max(min, pos) rejects pos <= min at setup (cartesian_cuboid_n.rs:149-162, 321)
    tiny = np.nextafter(lo, hi)
    np.maximum(pos, tiny, out=pos)
    return pos


def cell_sorting(
    n_cells: int = 1600,
    domain_size: float = 110.0,
    *,
    seed: int = 1,
    dim: int = 3,
    dt: float = 0.2,
    n_threads: int = 4,
    bench_sign: bool = False,
    pair_weight_mode: int = _abi.CRB_WEIGHT_CHILI,
    solver_order: int = 2,
) -> Workload:
    """cellular_raza-examples/cell_sorting as shipped (src/main.rs:9-24, 113-170): NewtonDamped3D
    (damping 2, mass 1) + species potential (R 6, strength 2, range 1.5); voxels from
    interaction range 18.  ``n <= N_CELLS_1`` are BlueCell => one more blue than red (:132)."""
    rng = np.random.default_rng(seed)
    pos = _uniform_positions(rng, n_cells, dim, 0.0, domain_size)
    vel = np.zeros((n_cells, dim))
    half = n_cells // 2
    species = (np.arange(n_cells) <= half).astype(np.uint8)  # 1 = BlueCell, 0 = RedCell
    mn = [0.0] * dim
    mx = [domain_size] * dim
    nv, dx = _domain_from_interaction_range(mn, mx, 1.5 * 6.0 * 2.0)
    cfg = dict(
        dim=dim,
        mechanics_kind=_abi.CRB_MECH_NEWTON_DAMPED,
        interaction_kind=_abi.CRB_INT_CELL_SORTING_BENCH if bench_sign else _abi.CRB_INT_CELL_SORTING,
        domain_min=mn,
        domain_max=mx,
        n_voxels=nv,
        dx=dx,
        dt=dt,
        solver_order=solver_order,
        pair_weight_mode=pair_weight_mode,
        ref_n_subdomains=n_threads,
    )
    return Workload(
        name=f"cell_sorting_{dim}d_{n_cells}",
        cfg_kwargs=cfg,
        pos=pos,
        vel=vel,
        mech_params=np.tile([2.0, 1.0], (n_cells, 1)),
        inter_params=np.tile([6.0, 2.0, 1.5], (n_cells, 1)),
        species=species,
    )


def cell_sorting_scaled(n_cells: int, density_per_voxel: float = 7.1, **kw) -> Workload:
    """cell_sorting at constant density: L is a multiple of the 18-unit voxel edge (SURVEY.md §8d:
    C2 = 1 M cells in L = 936 = 52 * 18 => rho ~ 7.1; the benchmark rule of
    cellular_raza-benchmarks/src/cell_sorting.rs:280-285 is rho ~ 2.5)."""
    dim = kw.get("dim", 3)
    n_side = max(int(round((n_cells / density_per_voxel) ** (1.0 / dim))), 1)
    return cell_sorting(n_cells, 18.0 * n_side, **kw)


def brownian_morse(
    n_cells: int,
    domain_size: float,
    *,
    seed: int = 2,
    dim: int = 3,
    dt: float = 0.01,
    n_threads: int = 1,
    noise_mode: int = _abi.CRB_NOISE_TABLE,
    noise_steps: int = 0,
    pair_weight_mode: int = _abi.CRB_WEIGHT_CHILI,
    solver_order: int = 2,
    langevin: bool = False,
) -> Workload:
    """C3: Brownian3D{D 0.3, kT 0.3} (cell_sorting/src/main-brownian.rs:115-116) +
    MorsePotential{radius 3, stiffness 0.5, strength 0.01} (tests/determinism_mechanics_interaction.rs:84-89)
    with cutoff = voxel edge = 12."""
    rng = np.random.default_rng(seed)
    pos = _uniform_positions(rng, n_cells, dim, 0.0, domain_size)
    if langevin:
        vel = rng.normal(0.0, 0.1, size=(n_cells, dim))
        mech = np.tile([1.0, 0.5, 0.3], (n_cells, 1))  # mass, damping, kT
        kind = _abi.CRB_MECH_LANGEVIN
    else:
        vel = np.zeros((n_cells, dim))
        mech = np.tile([0.3, 0.3], (n_cells, 1))
        kind = _abi.CRB_MECH_BROWNIAN
    mn = [0.0] * dim
    mx = [domain_size] * dim
    nv, dx = _domain_from_interaction_range(mn, mx, 12.0)
    cfg = dict(
        dim=dim,
        mechanics_kind=kind,
        interaction_kind=_abi.CRB_INT_MORSE,
        domain_min=mn,
        domain_max=mx,
        n_voxels=nv,
        dx=dx,
        dt=dt,
        solver_order=solver_order,
        pair_weight_mode=pair_weight_mode,
        ref_n_subdomains=n_threads,
        noise_mode=noise_mode,
        rng_seed=seed,
    )
    return Workload(
        name=f"{'langevin' if langevin else 'brownian'}_morse_{dim}d_{n_cells}",
        cfg_kwargs=cfg,
        pos=pos,
        vel=vel,
        mech_params=mech,
        inter_params=np.tile([3.0, 0.5, 12.0, 0.01], (n_cells, 1)),
        species=None,
        noise_steps=noise_steps,
    )


def noise_table(n_steps: int, n_ids: int, dim: int, seed: int = 7) -> np.ndarray:
    """Standard-normal table z[step][agent id][d] fed to oracle and device alike."""
    return np.random.default_rng(seed).standard_normal((n_steps, n_ids, dim))


def cell_sorting_slab(cells_per_rank: int, rank: int, world: int, density_per_voxel: float = 7.1, *,
                      seed: int = 1, n_threads: int = 1, dt: float = 0.2,
                      pair_weight_mode: int = _abi.CRB_WEIGHT_CHILI) -> Workload:
    """Weak-scaling population (BASELINE.json configs[4]): the global domain is ``world`` cubes of
    ``cells_per_rank`` agents stacked along axis 0; only THIS rank's agents are generated (independent
    PCG64 stream per rank), ids are globally unique.  ``cfg_kwargs`` describe the GLOBAL domain."""
    n_side = max(int(round((cells_per_rank / density_per_voxel) ** (1.0 / 3.0))), 1)
    edge = 18.0 * n_side
    rng = np.random.default_rng([seed, rank])
    pos = _uniform_positions(rng, cells_per_rank, 3, 0.0, edge)
    pos[:, 0] += rank * edge
    if rank > 0:
        # the plane index of x == rank * edge must be the slab's first plane
        np.maximum(pos[:, 0], rank * edge, out=pos[:, 0])
    half = cells_per_rank // 2
    species = (np.arange(cells_per_rank) <= half).astype(np.uint8)
    mn = [0.0, 0.0, 0.0]
    mx = [edge * world, edge, edge]
    nv, dx = _domain_from_interaction_range(mn, mx, 18.0)
    cfg = dict(
        dim=3,
        mechanics_kind=_abi.CRB_MECH_NEWTON_DAMPED,
        interaction_kind=_abi.CRB_INT_CELL_SORTING,
        domain_min=mn, domain_max=mx, n_voxels=nv, dx=dx, dt=dt, solver_order=2,
        pair_weight_mode=pair_weight_mode, ref_n_subdomains=n_threads,
    )
    w = Workload(
        name=f"cell_sorting_slab_{cells_per_rank}x{world}",
        cfg_kwargs=cfg,
        pos=pos,
        vel=np.zeros((cells_per_rank, 3)),
        mech_params=np.tile([2.0, 1.0], (cells_per_rank, 1)),
        inter_params=np.tile([6.0, 2.0, 1.5], (cells_per_rank, 1)),
        species=species,
    )
    w.extra["ids"] = (np.arange(cells_per_rank, dtype=np.uint64) + np.uint64(rank) * np.uint64(cells_per_rank))
    w.extra["n_side"] = n_side
    return w


def brownian_morse_scaled(n_cells: int, density_per_voxel: float = 4.0, **kw) -> Workload:
    """C3 at constant density: L is a multiple of the 12-unit voxel edge (8 M cells in
    L = 1,512 = 126 * 12 => rho ~ 4)."""
    n_side = max(int(round((n_cells / density_per_voxel) ** (1.0 / 3.0))), 1)
    return brownian_morse(n_cells, 12.0 * n_side, **kw)


def brownian_morse_slab(cells_total: int, rank: int, world: int, density_per_voxel: float = 4.0, *,
                        seed: int = 2, dt: float = 0.01, noise_mode: int = _abi.CRB_NOISE_PHILOX,
                        n_threads: int = 1) -> Workload:
    """BASELINE.json configs[2]: Brownian3D + MorsePotential, ``cells_total`` agents in ONE cube of
    n_side^3 voxels (8 M: 126^3) cut into ``world`` slabs of whole planes (16,16,16,16,16,16,15,15 for 8
    GPUs); only THIS rank's agents are generated, in proportion to its planes."""
    from .slabs import slab_planes

    n_side = max(int(round((cells_total / density_per_voxel) ** (1.0 / 3.0))), 1)
    edge = 12.0 * n_side
    first, count = slab_planes(n_side, world)[rank]
    firsts = [f for f, _ in slab_planes(n_side, world)]
    lo_id = int(round(cells_total * firsts[rank] / n_side))
    hi_id = int(round(cells_total * (first + count) / n_side))
    n = hi_id - lo_id
    rng = np.random.default_rng([seed, rank])
    pos = _uniform_positions(rng, n, 3, 0.0, edge)
    x_lo, x_hi = first * 12.0, (first + count) * 12.0
    pos[:, 0] = x_lo + (pos[:, 0] / edge) * (x_hi - x_lo)
    np.clip(pos[:, 0], x_lo if rank else np.nextafter(0.0, 1.0), np.nextafter(x_hi, x_lo), out=pos[:, 0])
    mn, mx = [0.0] * 3, [edge] * 3
    nv, dx = _domain_from_interaction_range(mn, mx, 12.0)
    cfg = dict(dim=3, mechanics_kind=_abi.CRB_MECH_BROWNIAN, interaction_kind=_abi.CRB_INT_MORSE,
               domain_min=mn, domain_max=mx, n_voxels=nv, dx=dx, dt=dt, solver_order=2,
               pair_weight_mode=_abi.CRB_WEIGHT_CHILI, ref_n_subdomains=n_threads, noise_mode=noise_mode,
               rng_seed=seed)
    w = Workload(name=f"brownian_morse_slab_{cells_total}/{world}", cfg_kwargs=cfg, pos=pos,
                 vel=np.zeros((n, 3)), mech_params=np.tile([0.3, 0.3], (n, 1)),
                 inter_params=np.tile([3.0, 0.5, 12.0, 0.01], (n, 1)), species=None)
    w.extra["ids"] = np.arange(lo_id, hi_id, dtype=np.uint64)
    w.extra["n_side"] = n_side
    return w


# ---------------------------------------------------------------------------------------------------
MICRO_METRE = 1e-6
MINUTE = 1.0


def bacterial_rods(
    n_tiles_x: int = 1,
    n_tiles_z: int = 1,
    *,
    rods_per_tile: int = 5,
    seed: int = 5,
    n_vertices: int = 5,
    dt: float = 0.1 * MINUTE,
    n_threads: int = 8,
    spring_length_range=(1.5, 2.5),
    pair_weight_mode: int = _abi.CRB_WEIGHT_CHILI,
    solver_order: int = 2,
    capacity_factor: float = 2.0,
    tile_x0: int = 0,
    n_tiles_x_total: int | None = None,
) -> Workload:
    """cellular_raza-examples/bacterial_rods as shipped (src/main.rs:111-212), tiled: every tile is the
    shipped domain (200 x 15 x 45 um, [8, 1, 2] voxels) holding ``rods_per_tile`` rods placed by the
    shipped rule (:143-163: first vertex in [dx, 2dx] x [dx/3, 2dx/3] x [dx, 2dx] of the tile with
    dx = 15 um, direction in the x-y plane with +-pi/8 jitter per segment, 3 um segments, spring length
    drawn from ``spring_length_range`` um).  Tiles repeat along x and z, so the rod density of the shipped
    example is kept (BASELINE.json configs[3]: "2 M rods ... domain scaled at constant rod density").
    Units: metres and minutes like the example.  ChaCha8 cannot be reproduced here: numpy PCG64."""
    rng = np.random.default_rng(seed)
    V = n_vertices
    um = MICRO_METRE
    seg = 3.0 * um
    delta_x = seg * V
    domain_size = 50.0 * um
    tile = np.array([4.0 * domain_size, delta_x, 3.0 * delta_x])
    n = n_tiles_x * n_tiles_z * rods_per_tile
    # (slabs: only the tiles [tile_x0, tile_x0 + n_tiles_x) of a domain n_tiles_x_total tiles long are populated)
    n_tot_x = n_tiles_x if n_tiles_x_total is None else n_tiles_x_total
    tx = np.repeat(np.arange(n_tiles_x), n_tiles_z * rods_per_tile) + tile_x0
    tz = np.tile(np.repeat(np.arange(n_tiles_z), rods_per_tile), n_tiles_x)
    pos = np.zeros((n, V, 3))
    pos[:, 0, 0] = rng.uniform(delta_x, 2.0 * delta_x, n) + tx * tile[0]
    pos[:, 0, 1] = rng.uniform(delta_x / 3.0, delta_x * 2.0 / 3.0, n)
    pos[:, 0, 2] = rng.uniform(delta_x, 2.0 * delta_x, n) + tz * tile[2]
    # directions are redrawn until every vertex lies inside the domain: a rod that starts across a wall
    # is folded by the reflective boundary in the first step (bacterial_rods.rs:489-533) and the
    # curvature force 2 * rigidity * tan(angle / 2) of a folded joint diverges
    todo = np.arange(n)
    margin = 0.25 * um
    while todo.size:
        theta = rng.uniform(0.0, 2.0 * np.pi, todo.size)
        for i in range(1, V):
            phi = theta + rng.uniform(-np.pi / 8.0, np.pi / 8.0, todo.size)
            pos[todo, i, 0] = pos[todo, i - 1, 0] + seg * np.cos(phi)
            pos[todo, i, 1] = pos[todo, i - 1, 1] + seg * np.sin(phi)
            pos[todo, i, 2] = pos[todo, i - 1, 2]
        y, x = pos[todo, :, 1], pos[todo, :, 0]
        bad = ((y.min(axis=1) < margin) | (y.max(axis=1) > tile[1] - margin) | (x.min(axis=1) < margin)
               | (x.max(axis=1) > tile[0] * n_tot_x - margin))
        todo = todo[bad]
    spring_length = rng.uniform(spring_length_range[0], spring_length_range[1], n) * um
    growth_base = 0.1 * um / MINUTE
    mech = np.tile(
        [0.0, 10.0 / MINUTE**2, 1.0 * um / MINUTE**2, 0.0, 1.5 / MINUTE, growth_base, 6.0 * um, growth_base],
        (n, 1),
    )
    mech[:, 3] = spring_length
    inter = np.tile([3.0 * um, 0.5 / um, 5.0 * um, 0.1 * um**2 / MINUTE**2], (n, 1))
    mn = [0.0, 0.0, 0.0]
    mx = [tile[0] * n_tot_x, tile[1], tile[2] * n_tiles_z]
    nv = np.array([8 * n_tot_x, 1, 2 * n_tiles_z], dtype=np.uint64)
    dx = (np.asarray(mx) - np.asarray(mn)) / nv.astype(np.float64)  # from_boundaries_and_n_voxels
    cfg = dict(
        dim=3,
        mechanics_kind=_abi.CRB_MECH_ROD,
        interaction_kind=_abi.CRB_INT_MORSE,
        domain_min=mn, domain_max=mx, n_voxels=nv, dx=dx, dt=dt, solver_order=solver_order,
        pair_weight_mode=pair_weight_mode, ref_n_subdomains=n_threads, rod_n_vertices=V,
        capacity=max(int(n * capacity_factor), 16),
    )
    return Workload(
        name=f"bacterial_rods_{n}",
        cfg_kwargs=cfg,
        pos=pos,
        vel=np.zeros((n, V, 3)),
        mech_params=mech,
        inter_params=inter,
        species=None,
    )


def bacterial_rods_slab(rods_per_rank: int, rank: int, world: int, *, rods_per_tile: int = 5, seed: int = 5,
                        n_threads: int = 8, spring_length_range=(4.0, 6.0)) -> Workload:
    """Weak-scaling rod population (BASELINE.json configs[3] per GPU): the domain of :func:`bacterial_rods` at the
    shipped rod density, ``world`` times longer along x; only THIS rank's tiles are generated (independent PCG64
    stream per rank), ids are globally unique.  ``cfg_kwargs`` describe the GLOBAL domain."""
    side = max(int(round((rods_per_rank / rods_per_tile) ** 0.5)), 1)
    w = bacterial_rods(side, side, rods_per_tile=rods_per_tile, seed=[seed, rank], n_threads=n_threads,
                       spring_length_range=spring_length_range, tile_x0=rank * side, n_tiles_x_total=world * side)
    w.extra["ids"] = np.arange(w.n, dtype=np.uint64) + np.uint64(rank) * np.uint64(w.n)
    return w
