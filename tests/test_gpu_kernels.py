"""Kernel-level checks on the device: the reciprocal square root used by every potential, and the
two force kernels (per-agent over global memory / shared-memory tiled) against each other."""
import ctypes as C
import os

import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads
from cellular_raza_b200.device import DeviceSimulation

pytestmark = pytest.mark.gpu


def test_rsqrt_accuracy(crb_lib):
    rng = np.random.default_rng(3)
    x = np.concatenate([
        rng.uniform(1e-6, 4e3, 200_000),
        10.0 ** rng.uniform(-300, 300, 100_000),
        np.array([1.0, 4.0, 0.25, 324.0, 2.0, 3.0, 1e-300, 1e300]),
    ])
    y = np.empty_like(x)
    rc = crb_lib.crb_debug_rsqrt(0, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), x.size)
    assert rc == 0
    ref = 1.0 / np.sqrt(x.astype(np.longdouble))
    rel = np.abs((y.astype(np.longdouble) - ref) / ref).astype(np.float64)
    assert rel.max() <= 2.5e-16, f"max rel err {rel.max():.3e} at x={x[rel.argmax()]!r}"


def _forces(w, kernel, monkeypatch):
    monkeypatch.setenv("CRB_FORCE_KERNEL", kernel)
    sim = DeviceSimulation(w.config())
    sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    f = sim.debug_forces()
    ids = sim.get_agents()[2]
    sim.close()
    return f[np.argsort(ids)]


@pytest.mark.parametrize("make", [
    lambda: workloads.cell_sorting(),
    lambda: workloads.cell_sorting(n_threads=4),
    lambda: workloads.cell_sorting(n_cells=3000, domain_size=90.0, dim=2),
    lambda: workloads.cell_sorting_scaled(200_000, 7.1, n_threads=3),
    lambda: workloads.cell_sorting_scaled(100_000, 2.5),
    lambda: workloads.brownian_morse(50_000, 12.0 * 23, noise_mode=_abi.CRB_NOISE_NONE),
])
def test_tile_kernel_equals_agent_kernel(crb_lib, monkeypatch, make):
    """same candidates, same coef(), same ascending-index summation order => identical forces"""
    w = make()
    fa = _forces(w, "agent", monkeypatch)
    scale = np.abs(fa).max()
    for kernel in ("tile", "tile2", "tile3"):
        ft = _forces(w, kernel, monkeypatch)
        assert np.abs(fa - ft).max() <= 1e-14 * scale, kernel
        assert np.array_equal(fa, ft), f"{kernel} and per-agent kernels should agree bitwise"


def test_tile_kernel_dense_voxels_fall_back(crb_lib, monkeypatch):
    """a cluster whose 27-voxel neighbourhood exceeds the staging capacity takes the per-agent path
    inside k_force_tile; results must not change."""
    w = workloads.cell_sorting(n_cells=6000, domain_size=54.0)  # 3^3 voxels, 222 agents per voxel
    fa = _forces(w, "agent", monkeypatch)
    for kernel in ("tile", "tile2", "tile3"):
        assert np.array_equal(fa, _forces(w, kernel, monkeypatch)), kernel


def _box_workload(n_voxels, density, seed, n_threads, clustered=False):
    """cell_sorting population in a NON-cubic box of n_voxels voxels of edge 18 (ragged tiles: odd
    plane / row counts, single-voxel axes) at `density` agents per voxel; `clustered` piles a tenth of
    the agents into one voxel so that a staging window overflows."""
    base = workloads.cell_sorting(n_cells=8, domain_size=36.0, n_threads=n_threads)
    nv = np.asarray(n_voxels, dtype=np.uint64)
    n = max(int(density * float(nv.prod())), 2)
    rng = np.random.default_rng(seed)
    mx = 18.0 * nv.astype(np.float64)
    pos = rng.uniform(1e-9, 1.0, (n, 3)) * (mx - 2e-9)
    if clustered:
        k = n // 10
        pos[:k] = rng.uniform(0.05, 0.95, (k, 3)) * 18.0 + 18.0 * (nv // 2).astype(np.float64)
    cfg = dict(base.cfg_kwargs)
    cfg.update(domain_max=mx.tolist(), n_voxels=nv, dx=np.full(3, 18.0))
    return workloads.Workload(
        name="box", cfg_kwargs=cfg, pos=pos, vel=np.zeros((n, 3)), mech_params=np.tile([2.0, 1.0], (n, 1)),
        inter_params=np.tile([6.0, 2.0, 1.5], (n, 1)), species=(rng.random(n) < 0.5).astype(np.uint8))


@pytest.mark.parametrize("n_voxels,density,n_threads,clustered", [
    ((1, 1, 1), 9.0, 1, False),       # a single voxel: every tile row but one is invalid
    ((7, 1, 2), 6.0, 2, False),       # the rods example's shape: single-voxel axis 1 (SURVEY.md §7 vii)
    ((3, 5, 11), 7.1, 3, False),      # odd plane and row counts: half-empty 2 x 2 tiles on two faces
    ((5, 4, 37), 2.5, 1, False),      # long fastest axis, benchmark density: several chunks per tile
    ((4, 6, 9), 31.0, 4, False),      # dense: candidates exceed the staging capacity => z-windows
    ((6, 6, 6), 7.1, 1, True),        # one overfull voxel: per-agent path inside the tile kernels
    ((9, 2, 64), 0.3, 5, False),      # nearly empty voxels, most tiles without targets
])
def test_tile_kernels_ragged_geometries(crb_lib, monkeypatch, n_voxels, density, n_threads, clustered):
    """k_force_tile2 / k_force_tile == k_force_integrate BITWISE on ragged grids, and the same again
    after steps (reflection at the walls, migration, cell-list rebuild feed the next force pass)."""
    w = _box_workload(n_voxels, density, seed=11, n_threads=n_threads, clustered=clustered)
    fa = _forces(w, "agent", monkeypatch)
    for kernel in ("tile", "tile2", "tile3"):
        assert np.array_equal(fa, _forces(w, kernel, monkeypatch)), kernel
    states = []
    for kernel in ("agent", "tile2"):
        monkeypatch.setenv("CRB_FORCE_KERNEL", kernel)
        sim = DeviceSimulation(w.config())
        sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
        sim.step(3 if (clustered or density > 20) else 12)
        pos, vel, ids, vox = sim.get_agents()
        o = np.argsort(ids)
        states.append((pos[o], vel[o], vox[o]))
        sim.close()
    for other in states[1:]:
        for a, b in zip(states[0], other):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("n_steps", [4, 5, 21, 40])
def test_graph_batches_equal_plain_loop(crb_lib, monkeypatch, n_steps):
    """crb_step(n) replays CUDA graphs of two steps each for deterministic mechanics (csrc/crb_api.cu
    crb_step); the same kernels in the same order must give the bits of the plain launch loop, for even and
    odd batch lengths (both buffer orientations) and across consecutive batches."""
    w = workloads.cell_sorting(n_threads=4)

    def run(no_graph):
        if no_graph:
            monkeypatch.setenv("CRB_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("CRB_NO_GRAPH", raising=False)
        d = DeviceSimulation(w.config())
        d.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
        l0 = d.kernel_launches
        d.step(n_steps)
        launches = d.kernel_launches - l0
        d.step(3)          # below the graph threshold: plain
        d.step(n_steps)    # second batch, possibly on the other buffer orientation
        out = d.get_agents() + d.get_aux() + (d.get_migrations(),)
        d.close()
        return out, launches

    a, la = run(False)
    b, lb = run(True)
    # the launch counter counts the kernels inside the graphs: force + scan (one launch: 216 voxels fit one
    # scan tile) + fill + reorder per step
    assert la == lb == 4 * n_steps
    for x, y in zip(a[:7], b[:7]):
        np.testing.assert_array_equal(x, y)
    oa, ob = np.argsort(a[7][0]), np.argsort(b[7][0])  # the migration tap is unordered
    for x, y in zip(a[7], b[7]):
        np.testing.assert_array_equal(x[oa], y[ob])


def test_exp_bounded_accuracy(crb_lib):
    """the lean exponential of the Morse potential (csrc/crb_kernels.cuh exp_bounded) against long double:
    relative error below 2.3e-16 over the whole admitted range |x| < 700."""
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-700.0, 700.0, 300_000), rng.uniform(-5.0, 5.0, 300_000),
                        np.array([0.0, 1.0, -1.0, 0.5, 3.0, -3.0, 1e-300, -1e-17, 699.9, -699.9,
                                  0.34657359027997264, -0.34657359027997264])])
    y = np.empty_like(x)
    rc = crb_lib.crb_debug_exp(0, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), x.size)
    assert rc == 0
    ref = np.exp(x.astype(np.longdouble))
    rel = np.abs((y.astype(np.longdouble) - ref) / ref).astype(np.float64)
    assert rel.max() <= 2.3e-16, f"max rel err {rel.max():.3e} at x={x[rel.argmax()]!r}"


def test_graphs_are_dropped_when_the_population_changes(crb_lib, monkeypatch):
    """the kernel parameters baked into the step graphs (agent count, tile plan, uniform rows) change with
    crb_set_agents: a second population on the same handle must not replay the first one's graphs."""
    w1 = workloads.cell_sorting(n_threads=1)
    w2 = workloads.cell_sorting(n_cells=700, domain_size=72.0, seed=4, n_threads=1)

    def run(no_graph):
        if no_graph:
            monkeypatch.setenv("CRB_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("CRB_NO_GRAPH", raising=False)
        d = DeviceSimulation(w1.config())
        outs = []
        for w in (w1, w2, w1):
            if w is w2:
                # a different domain needs its own handle in general; same handle is legal when the grid is the
                # same, so w2 reuses w1's domain with fewer agents
                pos = w2.pos / 72.0 * 110.0
                d.set_agents(pos, w2.vel, w2.mech_params * [1.0, 1.5], w2.inter_params, w2.species)
            else:
                d.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
            d.step(9)
            outs.append(d.get_agents())
        d.close()
        return outs

    for a, b in zip(run(False), run(True)):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
