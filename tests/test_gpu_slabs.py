"""Slab decomposition on ONE GPU: several `crb_sim` handles with (rank, world_size)
exchange halo and migration buffers by device copies and must reproduce the single-handle run BIT FOR
BIT (same candidate order, same pair weights, same integration) — the multi-device test the reference
never had (its integration tests all use n_threads: 1)."""
import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads
from cellular_raza_b200.device import DeviceSimulation
from cellular_raza_b200.slabs import LocalSlabGroup

import parity

pytestmark = pytest.mark.gpu


def _single(w):
    sim = DeviceSimulation(w.config())
    sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    return sim


def _group(w, world, **caps):
    grp = LocalSlabGroup(lambda: w.config(), world, **caps)
    grp.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    return grp


def _by_id(sim):
    pos, vel, ids, vox = sim.get_agents()
    o = np.argsort(ids, kind="stable")
    return pos[o], vel[o], ids[o], vox[o]


@pytest.mark.parametrize("world", [2, 3, 6])
@pytest.mark.parametrize("n_threads", [1, 4])
def test_c1_slabs_equal_single_handle(crb_lib, world, n_threads):
    w = workloads.cell_sorting(n_threads=n_threads)
    one, grp = _single(w), _group(w, world)
    assert grp.n_agents == one.n_agents
    assert grp.count_candidate_pairs() == one.count_candidate_pairs()
    f1 = one.debug_forces()[np.argsort(one.get_agents()[2])]
    fg = grp.debug_forces()[np.argsort(grp.get_agents()[2])]
    assert np.array_equal(f1, fg)
    migrated = 0
    for _ in range(4):
        one.step(10)
        grp.step(10)
        migrated += sum(sum(rk.last_migration) for rk in grp.ranks)
        p1, v1, i1, x1 = _by_id(one)
        pg, vg, ig, xg = _by_id(grp)
        assert np.array_equal(i1, ig)
        assert np.array_equal(x1, xg), "voxel assignment differs"
        assert np.array_equal(p1, pg) and np.array_equal(v1, vg)
    hx1, hv1, hl1 = one.get_aux()
    hxg, hvg, hlg = grp.get_aux()
    o1, og = np.argsort(one.get_agents()[2]), np.argsort(grp.get_agents()[2])
    assert np.array_equal(hx1[o1], hxg[og]) and np.array_equal(hv1[o1], hvg[og]) and np.array_equal(hl1[o1], hlg[og])
    np.testing.assert_array_equal(one.get_voxel_counts(), grp.get_voxel_counts())
    one.close()
    grp.close()


def test_slabs_large_with_migration(crb_lib):
    w = workloads.cell_sorting_scaled(150_000, 7.1, n_threads=3)
    one, grp = _single(w), _group(w, 4)
    crossed = 0
    for _ in range(5):
        one.step(8)
        grp.step(8)
        crossed += sum(sum(rk.last_migration[:2]) for rk in grp.ranks)
    p1, v1, i1, x1 = _by_id(one)
    pg, vg, ig, xg = _by_id(grp)
    assert np.array_equal(i1, ig) and np.array_equal(x1, xg)
    assert np.array_equal(p1, pg) and np.array_equal(v1, vg)
    assert grp.n_agents == w.n
    # agents did change slab during the run (last step of every batch is recorded)
    assert crossed > 0
    one.close()
    grp.close()


def test_slabs_brownian_noise_table(crb_lib):
    w = workloads.brownian_morse(40_000, 12.0 * 22, noise_mode=_abi.CRB_NOISE_TABLE)
    z = workloads.noise_table(12, w.n, 3)
    one, grp = _single(w), _group(w, 3)
    one.set_noise(z)
    grp.set_noise(z)
    one.step(12)
    grp.step(12)
    p1, v1, i1, x1 = _by_id(one)
    pg, vg, ig, xg = _by_id(grp)
    assert np.array_equal(i1, ig) and np.array_equal(x1, xg) and np.array_equal(p1, pg)
    one.close()
    grp.close()


def test_slabs_heterogeneous_parameters(crb_lib):
    w = workloads.cell_sorting(n_cells=4000, domain_size=144.0)
    rng = np.random.default_rng(5)
    w.inter_params[:, 0] = rng.uniform(4.0, 6.0, w.n)  # radii
    w.mech_params[:, 0] = rng.uniform(1.5, 2.5, w.n)   # damping
    one, grp = _single(w), _group(w, 2)
    one.step(15)
    grp.step(15)
    p1, v1, i1, x1 = _by_id(one)
    pg, vg, ig, xg = _by_id(grp)
    assert np.array_equal(i1, ig) and np.array_equal(x1, xg)
    assert np.array_equal(p1, pg) and np.array_equal(v1, vg)
    one.close()
    grp.close()


def test_slabs_against_oracle(oracle_lib, crb_lib):
    """and the slab run still matches the CPU restatement of chili (n_threads = 4 weights)."""
    from oracle.chili_oracle import ChiliOracle

    w = workloads.cell_sorting(n_threads=4)
    grp = _group(w, 3)
    ora = ChiliOracle(w.config())
    ora.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    grp.step(30)
    ora.step(30)
    perr, verr, key_diff = parity.state_errors(grp, ora, 110.0)
    assert perr <= 1e-9 and not key_diff.any()
    grp.close()


def test_migration_capacity_error(crb_lib):
    w = workloads.cell_sorting(n_threads=1)
    w.vel[:] = np.array([40.0, 0.0, 0.0])  # everybody rushes along axis 0
    grp = _group(w, 2, migration_capacity=4)
    with pytest.raises(_abi.SimulationError) as e:
        grp.step(3)
    assert e.value.code == _abi.CRB_ERR_CAPACITY
    grp.close()


def test_overlapped_step_equals_plain_step(crb_lib):
    """crb_slab_compute_interior + crb_slab_compute_boundary (halo exchange hidden behind the interior
    planes, second stream) is bit-identical to crb_slab_compute after the halo import."""
    w = workloads.cell_sorting_scaled(120_000, 7.1, n_threads=2)
    a = LocalSlabGroup(lambda: w.config(), 3, overlap=True)
    b = LocalSlabGroup(lambda: w.config(), 3, overlap=False)
    for g in (a, b):
        g.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    for _ in range(3):
        a.step(7)
        b.step(7)
        pa, va, ia, xa = _by_id(a)
        pb, vb, ib, xb = _by_id(b)
        assert np.array_equal(ia, ib) and np.array_equal(xa, xb)
        assert np.array_equal(pa, pb) and np.array_equal(va, vb)
    # protocol errors: boundary without interior, compute while a boundary call is pending
    rk = a.ranks[0]
    with pytest.raises(Exception):
        rk.compute_boundary()
    rk.compute_interior(True)
    with pytest.raises(Exception):
        rk.compute(True)
    a.halo_exchange()
    rk.compute_boundary()
    a.close()
    b.close()


@pytest.mark.parametrize("world", [2, 5])
def test_1d_slabs_equal_single_handle(crb_lib, world):
    """1-D domain (NewtonDamped1D + Morse): the only axis is the slab axis, a halo 'plane' is ONE voxel."""
    n, L = 1400, 12.0 * 400
    w = workloads.brownian_morse(n, L, dim=1, n_threads=3, noise_mode=_abi.CRB_NOISE_NONE)
    w.cfg_kwargs["mechanics_kind"] = _abi.CRB_MECH_NEWTON_DAMPED
    w.mech_params = np.tile([0.1, 1.0], (n, 1))
    w.vel = np.random.default_rng(3).normal(0.0, 60.0, size=(n, 1))  # fast enough to cross slab faces
    one, grp = _single(w), _group(w, world, ghost_capacity=256, migration_capacity=256)
    f1 = one.debug_forces()[np.argsort(one.get_agents()[2])]
    fg = grp.debug_forces()[np.argsort(grp.get_agents()[2])]
    assert np.array_equal(f1, fg)
    migrated = 0
    for _ in range(5):
        one.step(20)
        grp.step(20)
        migrated += sum(sum(rk.last_migration) for rk in grp.ranks)
        for a, b in zip(_by_id(one), _by_id(grp)):
            assert np.array_equal(a, b)
    assert migrated > 0
    one.close()
    grp.close()
