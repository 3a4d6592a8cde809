"""Known-answer tests of the rods oracle (oracle/rods_oracle.cpp) against what the reference pins for
the bacterial-rods path, plus self-consistency of the restatement.  CPU only."""
import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads
from oracle import rods_oracle as ro

UM = 1e-6


def test_nearest_point_doctest():
    # cellular_raza-building-blocks/src/cell_building_blocks/interaction.rs:478-496
    dist, point, t = ro.nearest_point([0.0, 0.0], [1.0, 0.0], [-1.0, 2.0])
    assert abs(dist - 1.0 / np.sqrt(2.0)) < 1e-10
    assert abs(t - 0.25) < 1e-10
    assert abs(point[0] - 0.5) < 1e-10 and abs(point[1] - 0.5) < 1e-10


def test_closest_points_exact():
    # interaction.rs:820-851 (assert_eq!, exact)
    p1, p2 = [0.0, 0.0], [2.0, 0.0]
    cases = [([0.5, 1.0], [0.5, 0.0], 1.0), ([-1.0, 2.0], p1, np.sqrt(5.0)), ([3.0, -2.0], p2, np.sqrt(5.0))]
    for q, r, d in cases:
        dist, nearest, _ = ro.nearest_point(q, p1, p2)
        assert dist == d
        assert np.array_equal(nearest, np.asarray(r))


def test_divide_doctest():
    # bacterial_rods.rs:560-585: 7 vertices 0.5 apart, spring_length 0.5, radius 0.25
    pos = np.zeros((7, 2))
    pos[:, 0] = np.arange(7) * 0.5
    m1, m2, new_len = ro.divide(pos, 0.5, 0.25)
    gap = np.linalg.norm(m1[6] - m2[0])
    assert abs(gap - 2 * 0.25) < 1e-3
    # div_factor = 1/2 - radius / (n_rows * spring_length)
    assert new_len == (0.5 - 0.25 / (7 * 0.5)) * 0.5
    # both halves stay on the original polyline, mother keeps the first vertex, daughter the last
    assert np.array_equal(m1[0], pos[0]) and np.array_equal(m2[6], pos[6])
    assert np.allclose(m1[:, 1], 0) and np.allclose(m2[:, 1], 0)
    assert np.all(np.diff(m1[:, 0]) > 0) and np.all(np.diff(m2[:, 0]) > 0)


def test_divide_3d_bent_rod_arclength():
    rng = np.random.default_rng(3)
    pos = np.cumsum(rng.normal(0, 1, (5, 3)) + [3, 0, 0], axis=0)
    seg = np.linalg.norm(np.diff(pos, axis=0), axis=1)
    radius = 0.4
    m1, m2, _ = ro.divide(pos, 1.0, radius)
    # each half spans (L - 2 radius) / 2 of arc length in 4 equal pieces measured along the polyline
    seglen = (seg.sum() - 2 * radius) / 4 / 2
    assert seglen > 0
    assert np.allclose(m1[0], pos[0]) and np.allclose(m2[-1], pos[-1])


def test_divide_error_when_too_short():
    pos = np.zeros((5, 3))
    pos[:, 0] = np.arange(5) * 0.01  # total length 0.04 < 2 * radius
    with pytest.raises(ro.OracleError):
        ro.divide(pos, 0.01, 0.25)


def _row(**kw):
    d = dict(diffusion_constant=0.0, spring_tension=10.0, rigidity=1e-6, spring_length=3e-6, damping=1.5,
             growth_rate=1e-7, spring_length_threshold=6e-6, growth_base_rate=1e-7)
    d.update(kw)
    return [d[k] for k in ("diffusion_constant", "spring_tension", "rigidity", "spring_length", "damping",
                           "growth_rate", "spring_length_threshold", "growth_base_rate")]


def test_increment_relaxed_straight_rod_is_force_free():
    pos = np.zeros((5, 3))
    pos[:, 0] = np.arange(5) * 3e-6
    dx, dv = ro.increment(pos, np.zeros((5, 3)), _row())
    assert np.array_equal(dx, np.zeros((5, 3)))
    assert np.abs(dv).max() < 1e-18


def test_increment_stretched_rod_pulls_inwards_and_sums_to_zero():
    pos = np.zeros((5, 3))
    pos[:, 0] = np.arange(5) * 4e-6  # stretched by 1 um per segment
    _, dv = ro.increment(pos, np.zeros((5, 3)), _row(rigidity=0.0))
    # F = -dir * tension * (|d| - l), +-1/2 on the two vertices (bacterial_rods.rs:177-187)
    assert dv[0, 0] == pytest.approx(0.5 * 10.0 * 1e-6, rel=1e-12)
    assert dv[4, 0] == pytest.approx(-0.5 * 10.0 * 1e-6, rel=1e-12)
    assert abs(dv[1:4, 0]).max() < 1e-20
    assert abs(dv.sum(axis=0)).max() < 1e-20


def test_increment_bent_rod_curvature_force_balances():
    pos = np.array([[0, 0, 0], [3, 0, 0], [6, 1, 0], [9, 3, 0], [12, 6, 0]], dtype=float) * UM
    _, dv = ro.increment(pos, np.zeros((5, 3)), _row(spring_tension=0.0))
    assert abs(dv.sum(axis=0)).max() < 1e-18  # -1/2, +1, -1/2 per joint
    assert np.abs(dv).max() > 1e-9
    # damping: dv -= vel * damping, dx = vel
    vel = np.full((5, 3), 2e-6)
    dx, dv2 = ro.increment(pos, vel, _row(spring_tension=0.0))
    assert np.array_equal(dx, vel)
    assert np.allclose(dv2, dv - vel * 1.5, rtol=0, atol=1e-20)


def test_pair_force_symmetry_and_cutoff():
    inter = [3 * UM, 0.5 / UM, 5 * UM, 0.1 * UM**2]
    a = np.zeros((5, 3))
    a[:, 0] = np.arange(5) * 3 * UM
    b = a.copy()
    b[:, 1] += 4 * UM  # parallel, 4 um apart (< cutoff): repulsive since 4 < 2 * radius
    fo, fe = ro.pair_force(a, b, inter, 3 * UM)
    assert np.all(fo[:, 1] < 0) and np.all(fe[:, 1] >= 0)
    assert np.allclose(fo.sum(axis=0), -fe.sum(axis=0), rtol=1e-12, atol=1e-30)
    b[:, 1] += 2 * UM  # 6 um > cutoff: exactly zero
    fo, fe = ro.pair_force(a, b, inter, 3 * UM)
    assert not fo.any() and not fe.any()


def _run(n_threads, mode, steps, **kw):
    w = workloads.bacterial_rods(2, 2, n_threads=n_threads, pair_weight_mode=mode, **kw)
    o = ro.RodsOracle(w.config(), n_threads=2)
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    o.step(steps)
    return w, o


def test_example_runs_and_divides_with_chili_ids():
    # spring lengths close to the threshold: divisions within a few hundred steps
    w, o = _run(8, _abi.CRB_WEIGHT_CHILI, 400, spring_length_range=(5.6, 5.95))
    pos, vel, ident, parent, vox = o.get_agents()
    assert o.n_agents > w.n
    div = ident[:, 0] == 0
    assert div.any()
    # Division(voxel, counter): counter > number of initial cells of that voxel never collides
    assert len({tuple(r) for r in ident}) == len(ident)
    # every divided cell has a parent, initial cells do not
    assert np.all(parent[div, 0] != np.uint64(0xFFFFFFFFFFFFFFFF))
    assert np.all(parent[~div, 0] == np.uint64(0xFFFFFFFFFFFFFFFF))
    # all vertices inside the domain
    for c in range(3):
        assert pos[:, :, c].min() >= 0 and pos[:, :, c].max() <= w.cfg_kwargs["domain_max"][c]
    mp = o.get_params()
    assert np.all(mp[:, 3] <= 6e-6 + 1e-12)


def test_physical_mode_is_independent_of_thread_count():
    _, a = _run(1, _abi.CRB_WEIGHT_PHYSICAL, 60)
    _, b = _run(8, _abi.CRB_WEIGHT_PHYSICAL, 60)
    pa, pb = a.get_agents()[0], b.get_agents()[0]
    assert np.abs(pa - pb).max() < 1e-15


def test_determinism():
    _, a = _run(8, _abi.CRB_WEIGHT_CHILI, 50)
    _, b = _run(8, _abi.CRB_WEIGHT_CHILI, 50)
    assert np.array_equal(a.get_agents()[0], b.get_agents()[0])


def test_neighbor_counts_symmetric_for_equal_radii():
    w = workloads.bacterial_rods(2, 2)
    o = ro.RodsOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    f, a, nb = o.forces()
    assert nb.sum() % 2 == 0 and nb.sum() > 0


def test_domain_force_gel_pressure_and_surface_friction():
    """SubDomainForce of CartesianSubDomainRods (bacterial_rods.rs:433-477): every vertex feels
    (0, 0, -gel_pressure); a MOVING vertex within surface_friction_distance of a domain wall additionally
    feels -v/|v| * gel_pressure * surface_friction.  Straight relaxed rods without neighbours in range, so
    the step-1 force is the domain force alone."""
    from cellular_raza_b200 import _abi

    um = workloads.MICRO_METRE
    w = workloads.bacterial_rods(1, 1, rods_per_tile=3)
    V = 5
    seg = 3.0 * um
    pos = np.zeros((3, V, 3))
    vel = np.zeros((3, V, 3))
    for k, (x0, y0, z0) in enumerate([(20 * um, 7 * um, 20 * um), (100 * um, 0.5 * um, 20 * um), (150 * um, 7 * um, 20 * um)]):
        pos[k, :, 0] = x0 + seg * np.arange(V)
        pos[k, :, 1] = y0
        pos[k, :, 2] = z0
    vel[1, :, 0] = 3.0   # rod 1 lies 0.5 um from the y = 0 wall and moves along (3, 0, 4) / 5
    vel[1, :, 2] = 4.0
    vel[2, :, 1] = 1.0   # rod 2 moves but is far from every wall
    gel, fric, dist = 0.25, 0.5, 1.0 * um
    cfg = w.config(domain_force_kind=_abi.CRB_DOMAIN_FORCE_ROD_GEL, domain_force_strength=[gel, fric, dist])
    o = ro.RodsOracle(cfg, n_threads=2)
    o.set_agents(pos, vel, w.mech_params[:3], w.inter_params[:3])
    f, _, _ = o.forces()
    order = np.argsort(o.get_agents()[2][:, 1])  # Initial(n) identifiers -> input order
    f = f[order]
    np.testing.assert_array_equal(f[0], np.tile([0.0, 0.0, -gel], (V, 1)))
    np.testing.assert_array_equal(f[2], np.tile([0.0, 0.0, -gel], (V, 1)))
    expect = np.array([-(3.0 / 5.0) * gel * fric, 0.0, -gel - (4.0 / 5.0) * gel * fric])
    np.testing.assert_allclose(f[1], np.tile(expect, (V, 1)), rtol=1e-15, atol=0)
    # without the aspect the same rods feel nothing
    o2 = ro.RodsOracle(w.config(), n_threads=2)
    o2.set_agents(pos, vel, w.mech_params[:3], w.inter_params[:3])
    assert np.abs(o2.forces()[0]).max() == 0.0


def test_random_contribution_known_answer():
    """RodMechanics::get_random_contribution (bacterial_rods.rs:215-240): dpos = M * 2^-2 * D / dt with M filled
    column by column from Normal(0, sqrt(dt)), dvel = 0; the Euler helper (solvers.rs:303-312) adds dpos * dt.  A
    relaxed straight rod at rest feels no force, so one step moves vertex v, coordinate c by exactly
    ((sqrt(dt) z[c V + v]) * 0.25 * D / dt) * dt."""
    V, D, dt = 5, 0.5 * UM, 0.1
    w = workloads.bacterial_rods(1, 1, rods_per_tile=1, n_threads=1, solver_order=0)
    pos = np.zeros((1, V, 3))
    pos[0, :, 0] = 30 * UM + 3 * UM * np.arange(V)
    pos[0, :, 1] = 7 * UM
    pos[0, :, 2] = 20 * UM
    mech = w.mech_params.copy()
    mech[:, 0] = D
    mech[:, 3] = 3 * UM   # spring_length == segment length: no spring force
    mech[:, 5] = 0.0      # no growth
    z = np.random.default_rng(3).standard_normal((2, 1, 3 * V))
    o = ro.RodsOracle(w.config(noise_mode=_abi.CRB_NOISE_TABLE), n_threads=1)
    o.set_agents(pos, np.zeros_like(pos), mech, w.inter_params)
    o.set_noise(z)
    o.step(1)
    got = o.get_agents()[0][0]
    m = (0.0 + np.sqrt(dt) * z[0, 0].reshape(3, V).T)          # column-major fill: entry c V + v -> (v, c)
    want = (np.zeros((V, 3)) * dt + pos[0]) + ((m * 0.25 * D) / dt) * dt
    np.testing.assert_array_equal(got, want)
    # velocities get no random part (what is left is the rounding of the vertex spacing through the spring: 1e-21)
    assert np.abs(o.get_agents()[1]).max() < 1e-18
    # the table runs out after its two steps
    o.step(1)
    with pytest.raises(Exception):
        o.step(1)


def test_random_contribution_zero_diffusion_is_deterministic():
    w = workloads.bacterial_rods(2, 2, n_threads=2)
    a = ro.RodsOracle(w.config(noise_mode=_abi.CRB_NOISE_TABLE), n_threads=2)
    b = ro.RodsOracle(w.config(), n_threads=2)
    for o in (a, b):
        o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    a.set_noise(np.random.default_rng(1).standard_normal((5, w.n, 15)))
    a.step(5)
    b.step(5)
    assert np.array_equal(a.get_agents()[0], b.get_agents()[0])
