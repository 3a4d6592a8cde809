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


# ---- crb_slab_step: exchange over peer memory, device-side counts, no host synchronisation inside a batch ------
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("n_threads", [1, 4])
def test_native_slab_step_equals_single_handle(crb_lib, world, n_threads):
    """the whole step behind crb_slab_step (halo and migration stored straight into the neighbours' arenas, flags
    instead of host synchronisation) reproduces the single handle bit for bit, incl. AuxStorage and voxel counts"""
    w = workloads.cell_sorting(n_threads=n_threads)
    one, grp = _single(w), _group(w, world, transport="p2p")
    for _ in range(4):
        one.step(10)
        grp.step(10)
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
    # in-voxel (save) order as well: slabs are contiguous plain-index ranges, ranks in order
    assert np.array_equal(one.get_agents()[2], grp.get_agents()[2])
    one.close()
    grp.close()


def test_native_slab_step_large_with_migration_and_mixed_protocols(crb_lib):
    """150 k cells over 3 slabs: agents cross the faces every step; the native step and the six-phase protocol can
    alternate on the same handles"""
    w = workloads.cell_sorting_scaled(150_000, 7.1, n_threads=3)
    one, grp = _single(w), _group(w, 3, transport="p2p")
    n0 = [rk.n_agents for rk in grp.ranks]
    for k in range(4):
        one.step(6)
        if k == 2:
            grp.transport = "copy"
        grp.step(6)
        grp.transport = "p2p"
    p1, v1, i1, x1 = _by_id(one)
    pg, vg, ig, xg = _by_id(grp)
    assert np.array_equal(i1, ig) and np.array_equal(x1, xg)
    assert np.array_equal(p1, pg) and np.array_equal(v1, vg)
    assert grp.n_agents == w.n and [rk.n_agents for rk in grp.ranks] != n0  # the populations of the slabs changed
    one.close()
    grp.close()


def test_native_slab_step_with_per_agent_parameters_and_noise(crb_lib):
    """heterogeneous radii (per-agent kernel with the device-side count) + Brownian noise table over 2 slabs"""
    w = workloads.brownian_morse(20_000, 12.0 * 12, noise_mode=_abi.CRB_NOISE_TABLE)
    rng = np.random.default_rng(5)
    ip = w.inter_params.copy()
    ip[:, 0] *= rng.uniform(0.8, 1.2, w.n)
    z = rng.standard_normal((12, w.n, 3))
    one = DeviceSimulation(w.config())
    one.set_agents(w.pos, w.vel, w.mech_params, ip, w.species)
    one.set_noise(z)
    grp = LocalSlabGroup(lambda: w.config(), 2, transport="p2p")
    grp.set_agents(w.pos, w.vel, w.mech_params, ip, w.species)
    grp.set_noise(z)
    one.step(12)
    grp.step(12)
    p1, v1, i1, x1 = _by_id(one)
    pg, vg, ig, xg = _by_id(grp)
    assert np.array_equal(i1, ig) and np.array_equal(x1, xg) and np.array_equal(p1, pg)
    one.close()
    grp.close()


# ---- uniform vs per-agent parameters is a property of the WHOLE population -----------------
@pytest.mark.parametrize("transport", ["copy", "p2p"])
def test_slab_that_starts_empty_receives_agents_with_their_parameters(crb_lib, transport):
    """all agents start in the lower half of the domain: the upper slab is created empty and must integrate the
    agents that migrate into it with THEIR mass and radius (it used to fabricate an all-zero uniform row)"""
    w = workloads.cell_sorting(n_cells=600, domain_size=72.0, n_threads=2)
    pos = w.pos.copy()
    pos[:, 0] = pos[:, 0] * 0.45 + 1.0     # x in (1, 33.4): planes 0 and 1 of 4 => rank 0 of 2
    vel = np.zeros_like(pos)
    vel[:, 0] = 12.0                        # everybody heads for the upper slab (no one crosses two planes in a step)
    one = DeviceSimulation(w.config())
    one.set_agents(pos, vel, w.mech_params, w.inter_params, w.species)
    grp = LocalSlabGroup(lambda: w.config(), 2, transport=transport)
    grp.set_agents(pos, vel, w.mech_params, w.inter_params, w.species)
    assert grp.ranks[1].n_agents == 0
    one.step(12)
    grp.step(12)
    assert grp.ranks[1].n_agents > 0
    p1, v1, i1, x1 = _by_id(one)
    pg, vg, ig, xg = _by_id(grp)
    assert np.isfinite(pg).all()
    assert np.array_equal(i1, ig) and np.array_equal(x1, xg) and np.array_equal(p1, pg) and np.array_equal(v1, vg)
    one.close()
    grp.close()


@pytest.mark.parametrize("transport", ["copy", "p2p"])
def test_locally_homogeneous_slab_next_to_a_heterogeneous_one(crb_lib, transport):
    """the agents of the lower slab share one parameter row, those of the upper slab do not: both slabs must take
    the per-agent path (a slab used to decide from its own subset and ship halos without parameter columns)"""
    w = workloads.cell_sorting(n_cells=800, domain_size=72.0, n_threads=1)
    ip = w.inter_params.copy()
    upper = w.pos[:, 0] >= 36.0
    ip[upper, 0] *= np.random.default_rng(9).uniform(0.7, 1.3, int(upper.sum()))
    one = DeviceSimulation(w.config())
    one.set_agents(w.pos, w.vel, w.mech_params, ip, w.species)
    grp = LocalSlabGroup(lambda: w.config(), 2, transport=transport)
    grp.set_agents(w.pos, w.vel, w.mech_params, ip, w.species)
    f1 = one.debug_forces()[np.argsort(one.get_agents()[2])]
    fg = grp.debug_forces()[np.argsort(grp.get_agents()[2])]
    assert np.array_equal(f1, fg)
    one.step(10)
    grp.step(10)
    p1, v1, i1, x1 = _by_id(one)
    pg, vg, ig, xg = _by_id(grp)
    assert np.array_equal(i1, ig) and np.array_equal(x1, xg) and np.array_equal(p1, pg)
    one.close()
    grp.close()


def test_parameter_mode_mismatch_between_slabs_is_an_error(crb_lib):
    """a rank that declares a uniform row next to one that does not: the import of the halo latches an error
    instead of reading parameter columns that were never sent"""
    from cellular_raza_b200._abi import SimulationError
    from cellular_raza_b200.slabs import SlabRank, split_agents_by_slab

    w = workloads.cell_sorting(n_cells=600, domain_size=72.0, n_threads=1)
    ranks = [SlabRank(w.config(), r, 2, 0) for r in range(2)]
    owner = split_agents_by_slab(ranks[0].cfg, 2, w.pos)
    ids = np.arange(w.n, dtype=np.uint64)
    ranks[0].set_uniform_params(w.mech_params[0], w.inter_params[0])
    ranks[0].set_agents(w.pos[owner == 0], None, None, None, w.species[owner == 0], ids[owner == 0])
    ranks[1].set_agents(w.pos[owner == 1], None, w.mech_params[owner == 1], w.inter_params[owner == 1],
                        w.species[owner == 1], ids[owner == 1])
    ranks[0].halo_export()
    ranks[1].halo_in[0].copy_(ranks[0].halo_out[1])
    ranks[1].halo_import()
    with pytest.raises(SimulationError):
        ranks[1].sync()
    for rk in ranks:
        rk.close()


def test_agent_that_jumps_over_a_whole_slab_plane_is_an_error(crb_lib):
    """the slab protocol exchanges with the adjacent slabs only: an agent whose step carries it into a plane that is
    neither owned nor a ghost plane is an IndexError, never a silently misplaced agent"""
    from cellular_raza_b200._abi import SimulationError

    w = workloads.cell_sorting(n_cells=300, domain_size=108.0, n_threads=1)   # 6 planes of 18
    vel = np.zeros_like(w.pos)
    k = int(np.argmin(np.abs(w.pos[:, 0] - 60.0)))   # an agent of the upper slab ...
    vel[k, 0] = -250.0                                # ... thrown 50 units downwards in one step (dt = 0.2)
    grp = LocalSlabGroup(lambda: w.config(), 2, transport="copy")
    grp.set_agents(w.pos, vel, w.mech_params, w.inter_params, w.species)
    with pytest.raises(SimulationError, match="IndexError"):
        grp.step(1)
    grp.close()
