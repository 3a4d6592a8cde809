"""Bacterial rods cut into slabs (SURVEY.md §8e for BASELINE.json configs[3]): W slabs on ONE GPU driven through the
six-phase protocol (crb_slab_halo_export .. crb_slab_finish with mechanics_kind == CRB_MECH_ROD) must reproduce the
single-handle run bit for bit — positions, velocities, identifiers incl. Division(voxel, counter), parents, voxel
assignment, spring lengths and growth rates — in the same (voxel, in-voxel) order, with rods interacting across the
slab faces, migrating over them and dividing next to them."""
import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads
from cellular_raza_b200.rods import RodSimulation
from cellular_raza_b200.slabs import LocalRodSlabGroup

pytestmark = pytest.mark.gpu

UM = 1e-6


def _population(world_planes=32, shift=175.0 * UM, **kw):
    """the shipped rod placement, tiled 4 x 3 and shifted along x so that the rod clusters straddle the planes
    x = 200, 400, 600 um (the faces of 2 and 4 slabs); rods pushed out of the domain by the shift are dropped"""
    w = workloads.bacterial_rods(world_planes // 8, 3, rods_per_tile=8, spring_length_range=(5.5, 5.99),
                                 capacity_factor=4.0, **kw)
    pos = w.pos.copy()
    pos[:, :, 0] += shift
    Lx = w.cfg_kwargs["domain_max"][0]
    keep = (pos[:, :, 0].min(axis=1) > 1 * UM) & (pos[:, :, 0].max(axis=1) < Lx - 1 * UM)
    w.pos, w.vel = pos[keep], w.vel[keep]
    w.mech_params, w.inter_params = w.mech_params[keep], w.inter_params[keep]
    return w


def _single(w, **over):
    sim = RodSimulation(w.config(**over))
    sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    return sim


def _group(w, world, **over):
    grp = LocalRodSlabGroup(lambda: w.config(**over), world)
    grp.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    return grp


def _assert_same(grp, one):
    a, b = grp.get_agents(), one.get_agents()
    assert len(a[2]) == len(b[2]) == grp.n_agents == one.n_agents
    for x, y, name in zip(a, b, ("pos", "vel", "ident", "parent", "voxel")):
        assert np.array_equal(x, y), name
    pa, pb = grp.get_params(), one.get_params()
    assert np.array_equal(pa, pb)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("mode,n_threads", [(_abi.CRB_WEIGHT_CHILI, 8), (_abi.CRB_WEIGHT_PHYSICAL, 1)])
def test_forces_and_neighbor_counts_across_the_faces(crb_lib, world, mode, n_threads):
    w = _population(n_threads=n_threads, pair_weight_mode=mode)
    one, grp = _single(w), _group(w, world)
    _assert_same(grp, one)
    f1, nb1 = one.debug_forces(), one.debug_neighbor_counts()
    fg, nbg = grp.debug_forces(), grp.debug_neighbor_counts()
    assert nb1.sum() > 100
    assert np.array_equal(nbg, nb1)
    assert np.array_equal(fg, f1)  # same candidates in the same order: the same bits
    assert grp.count_candidate_pairs() == one.count_candidate_pairs()
    # the population does interact across a face: the ghost planes of the first slab hold rods
    n0 = grp.ranks[0].n_agents
    assert 0 < n0 < one.n_agents
    one.close()
    grp.close()


@pytest.mark.parametrize("world,order", [(2, 2), (4, 2), (4, 0)])
def test_trajectory_with_divisions_and_migration(crb_lib, world, order):
    w = _population(solver_order=order)
    one, grp = _single(w), _group(w, world)
    migrated = 0
    divisions = 0
    for _ in range(120):
        one.step(1)
        grp.step(1)
        migrated += sum(sum(rk.last_migration) for rk in grp.ranks)
        divisions += grp.n_divisions
        assert grp.n_divisions == one.n_divisions
    _assert_same(grp, one)
    assert divisions > 10, "the test population must divide"
    assert migrated > 0, "no rod crossed a slab face: the test does not exercise migration"
    hx1, hv1, hl1 = one.get_aux()
    parts = [rk.get_aux() for rk in grp.ranks]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), hx1)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), hl1)
    one.close()
    grp.close()


def test_random_contribution_over_slabs(crb_lib):
    """the shared noise table is indexed by identifier, not by slot: slabs draw the same numbers"""
    w = _population()
    mech = w.mech_params.copy()
    mech[:, 0] = 0.5e-6
    z = np.random.default_rng(5).standard_normal((40, 509, 15))
    one = RodSimulation(w.config(noise_mode=_abi.CRB_NOISE_TABLE))
    one.set_agents(w.pos, w.vel, mech, w.inter_params)
    one.set_noise(z)
    grp = LocalRodSlabGroup(lambda: w.config(noise_mode=_abi.CRB_NOISE_TABLE), 3)
    grp.set_agents(w.pos, w.vel, mech, w.inter_params)
    for rk in grp.ranks:
        rk.set_noise(z)
    for _ in range(40):
        one.step(1)
        grp.step(1)
    _assert_same(grp, one)
    one.close()
    grp.close()


def test_an_empty_slab_and_error_paths(crb_lib):
    w = _population()
    # 8 slabs of 4 planes: some of them start without a single rod and receive their first one by migration or never
    one, grp = _single(w), _group(w, 8)
    assert min(rk.n_agents for rk in grp.ranks) == 0
    for _ in range(30):
        one.step(1)
        grp.step(1)
    _assert_same(grp, one)
    # crb_slab_step (peer-memory form) and the interior / boundary split are refused for rods, loudly
    with pytest.raises(_abi.SimulationError):
        grp.ranks[0].slab_step(1)
    with pytest.raises(_abi.SimulationError):
        grp.ranks[0].compute_interior(True)
    one.close()
    grp.close()
