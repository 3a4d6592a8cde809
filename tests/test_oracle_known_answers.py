"""Pins the oracle (oracle/chili_oracle.cpp) against every known-answer test the reference holds for
the mechanics/interaction path.  The reference ships no golden vectors; these are
its in-crate tests and doctests re-expressed against the CPU restatement."""
import math

import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads


def test_n_voxels_from_interaction_range(oracle_lib):
    # cartesian_cuboid_n.rs:167-182 (doctest)
    nv, dx = oracle_lib.domain_from_interaction_range([2.0, 3.0, 1.0], [10.0, 10.0, 20.0], 2.0)
    assert nv.tolist() == [4, 3, 9]
    np.testing.assert_array_equal(dx, [8.0 / 4.0, 7.0 / 3.0, 19.0 / 9.0])


def test_cell_sorting_domain(oracle_lib):
    # cell_sorting/src/main.rs:148-152: [0,110]^3 with range 18 -> 6^3 voxels, dx = 110/6
    nv, dx = oracle_lib.domain_from_interaction_range([0.0] * 3, [110.0] * 3, 18.0)
    assert nv.tolist() == [6, 6, 6]
    np.testing.assert_array_equal(dx, [110.0 / 6.0] * 3)


def test_generate_subdomains(oracle_lib):
    # cartesian_cuboid_n.rs:380-402: 5^3 voxels into 4 subdomains; counts sum to 125
    nv, _ = oracle_lib.domain_from_interaction_range([0.0] * 3, [100.0] * 3, 20.0)
    assert nv.tolist() == [5, 5, 5]
    sizes = oracle_lib.subdomain_sizes(125, 4)
    assert len(sizes) == 4
    assert int(sizes.sum()) == 125
    assert sizes.tolist() == [32, 31, 31, 31]


def test_get_decomp_res_invariants(oracle_lib):
    # cartesian_cuboid_n.rs:670-699 (debug-assertion ranges)
    for n_voxel in range(1, 500):
        for n_regions in range(1, 100):
            res = oracle_lib.decomp_res(n_voxel, n_regions)
            assert res is not None, (n_voxel, n_regions)
            n, m, avg = res
            assert n + m == n_regions
            assert n * avg + m * (avg - 1) == n_voxel


def test_get_decomp_res_table(oracle_lib):
    # rows of the comment table at cartesian_cuboid_n.rs:12-18 that satisfy the invariants the
    # reference's own test asserts (rows 1-3 of that comment do not: e.g. "10 3 -> 1*4 + 3*3" is
    # the decomposition of 13 into 4).
    assert oracle_lib.decomp_res(225, 16) == (1, 15, 15)
    assert oracle_lib.decomp_res(225, 17) == (4, 13, 14)
    assert oracle_lib.decomp_res(13, 4) == (1, 3, 4)
    assert oracle_lib.decomp_res(10, 3) == (1, 2, 4)


def test_ring_buffer_order(oracle_lib):
    # chili/aux_storage.rs:862-868: push 1,3 -> [1,3]; :894-903 push 0,3 -> [0,3], size 2
    assert oracle_lib.ring([1.0, 3.0], 2).tolist() == [1.0, 3.0]
    assert oracle_lib.ring([0.0, 3.0], 10 if False else 2).tolist() == [0.0, 3.0]
    assert oracle_lib.ring([10.0], 2).tolist() == [10.0]
    # circ_buffer drops the oldest entry when full
    assert oracle_lib.ring([1.0, 2.0, 3.0], 2).tolist() == [2.0, 3.0]
    assert oracle_lib.ring([1.0, 2.0, 3.0], 1).tolist() == [3.0]
    assert oracle_lib.ring([1.0, 2.0], 0).tolist() == []


def test_xapy(oracle_lib):
    # concepts/src/reactions.rs:54-59
    rng = np.random.default_rng(0)
    for x, a, y in rng.normal(size=(100, 3)):
        assert oracle_lib.xapy(x, a, y) == x * a + y


def test_euler_exp_decay(oracle_lib):
    # chili/solvers.rs:588-617
    y0, lam, dt = 27.0, 0.3, 0.001
    y, t = y0, 0.0
    for _ in range(100):
        y = oracle_lib.helper_euler(y, -lam * y, dt, 0.0)
        t += dt
        e = dt * (y0 * lam**2) / (2.0 * lam) * (math.exp(lam * t) - 1.0)
        assert abs(y - y0 * math.exp(-lam * t)) < e


def test_adams_bashforth_2_harmonic_oscillator(oracle_lib):
    # chili/solvers.rs:619-710 (there in f32; the bound holds in f64 as well)
    y0, omega, dt = 2.8347, 0.319, 0.045
    rhs = lambda y: (y[1], -(omega**2) * y[0])
    exact = lambda t: (y0 * math.cos(omega * t), -y0 * omega * math.sin(omega * t))
    third = y0 * omega**3
    lte = third * (5.0 / 12.0 * dt**2)
    lip = (1.0, omega)
    gte = lambda t: tuple((math.exp(l * t) - 1.0) * lte / dt / l for l in lip)
    t = 0.0
    y = (y0, 0.0)
    store = [rhs(exact(t - dt)), rhs(exact(t - 2 * dt))]
    for i in range(100):
        dy = rhs(y)
        store[1] = store[0]
        store[0] = dy
        y = tuple(oracle_lib.helper_ab2(y[k], store[0][k], store[1][k], dt, 0.0) for k in range(2))
        t += dt
        e = gte(t)
        if i > 0:
            assert abs(y[0] - exact(t)[0]) < e[0]
            assert abs(y[1] - exact(t)[1]) < e[1]


def test_ab3_helper_formula(oracle_lib):
    # chili/solvers.rs:303-326: ((dx0*f0 + (dx1*f1 + dx2*f2))*dt + x)*1 + rand*dt
    x, d0, d1, d2, dt, r = 0.7, 1.3, -0.4, 2.2, 0.2, 0.05
    f0, f1, f2 = 23.0 / 12.0, -16.0 / 12.0, 5.0 / 12.0
    expect = ((d0 * f0 + (d1 * f1 + d2 * f2)) * dt + x) * 1.0 + r * dt
    assert oracle_lib.helper_ab3(x, d0, d1, d2, dt, r) == expect


def _free_motion_workload(dim=3):
    # chili/mod.rs:110-157: agents at pos = n with vel = 0.1, no interaction, dt = 0.1, tmax = 20
    n = 10
    pos = np.zeros((n, dim))
    pos[:, 0] = np.arange(n) + 0.5
    pos[:, 1:] = 50.0
    vel = np.zeros((n, dim))
    vel[:, 0] = 0.1
    from cellular_raza_b200.workloads import Workload

    cfg = dict(dim=dim, mechanics_kind=_abi.CRB_MECH_NEWTON_DAMPED, interaction_kind=_abi.CRB_INT_MORSE,
               domain_min=[0.0] * dim, domain_max=[100.0] * dim, n_voxels=[5] * dim, dx=[20.0] * dim, dt=0.1,
               ref_n_subdomains=1)
    return Workload("free", cfg, pos, vel, np.tile([0.0, 1.0], (n, 1)), np.tile([1.0, 1.0, 0.0, 0.0], (n, 1)))


def test_free_motion(oracle_lib):
    # chili/mod.rs:134-157: pos = key + iteration * 0.1 * vel within 1e-5
    w = _free_motion_workload()
    o = oracle_lib.ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    for it in range(1, 201):
        o.step(1)
        if it % 20 == 0:
            pos, _, ids, _ = o.get_agents()
            calc = ids.astype(float) + 0.5 + it * 0.1 * 0.1
            assert np.abs(pos[:, 0] - calc).max() < 1e-5


@pytest.mark.parametrize("dim,diff,seed", [(3, 1.0, 1), (3, 0.5, 2), (3, 0.25, 3), (2, 1.0, 4), (2, 0.5, 5), (2, 0.25, 6),
                                           (1, 1.0, 7), (1, 0.5, 8), (1, 0.25, 9)])
def test_brownian_msd(oracle_lib, dim, diff, seed):
    # tests/brownian_diffusion_constant_approx.rs:8-36, 110-124: 160 particles from the domain
    # centre, dt 1e-3, 200 steps, MSD = 2 dim D t within 3.5 sigma.  Noise comes from the table
    # contract instead of ChaCha8.
    n, L, dt, steps = 160, 200.0, 1e-3, 200
    from cellular_raza_b200.workloads import Workload

    pos = np.full((n, dim), L / 2.0)
    cfg = dict(dim=dim, mechanics_kind=_abi.CRB_MECH_BROWNIAN, interaction_kind=_abi.CRB_INT_MORSE,
               domain_min=[0.0] * dim, domain_max=[L] * dim, n_voxels=[3] * dim, dx=[L / 3.0] * dim, dt=dt,
               ref_n_subdomains=1, noise_mode=_abi.CRB_NOISE_TABLE)
    w = Workload("msd", cfg, pos, np.zeros((n, dim)), np.tile([diff, 1.0], (n, 1)),
                 np.tile([1.0, 1.0, 0.0, 0.0], (n, 1)))
    o = oracle_lib.ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    o.set_noise(workloads.noise_table(steps, n, dim, seed))
    for it in range(1, steps + 1):
        o.step(1)
        p = o.get_agents()[0]
        sq = ((p - L / 2.0) ** 2).sum(axis=1)
        assert abs(2.0 * dim * diff * it * dt - sq.mean()) <= 3.5 * sq.std()


@pytest.mark.parametrize("dim,diff,damping,seed", [(3, 80.0, 10.0, 1), (3, 40.0, 10.0, 2), (3, 20.0, 10.0, 3),
                                                   (3, 40.0, 1.0, 4), (3, 40.0, 0.1, 5), (2, 80.0, 10.0, 6),
                                                   (2, 40.0, 10.0, 7), (2, 20.0, 10.0, 8)])
def test_langevin_msd(oracle_lib, dim, diff, damping, seed):
    # tests/brownian_diffusion_constant_approx.rs:260-340 (test_langevin!, diffusion_constant_langevin_*):
    # mass 1, kT = D * damping * mass, v(0) = 0, 160 particles, dt 1e-3, 200 steps; for iteration > 1
    # MSD = -dim kT/m/g^2 (1 - e^{-g t})(3 - e^{-g t}) + 2 dim kT/m t/g within 3.5 sigma
    n, L, dt, steps = 160, 200.0, 1e-3, 200
    from cellular_raza_b200.workloads import Workload

    kT = diff * damping * 1.0
    pos = np.full((n, dim), L / 2.0)
    cfg = dict(dim=dim, mechanics_kind=_abi.CRB_MECH_LANGEVIN, interaction_kind=_abi.CRB_INT_MORSE,
               domain_min=[0.0] * dim, domain_max=[L] * dim, n_voxels=[3] * dim, dx=[L / 3.0] * dim, dt=dt,
               ref_n_subdomains=1, noise_mode=_abi.CRB_NOISE_TABLE)
    w = Workload("msd_langevin", cfg, pos, np.zeros((n, dim)), np.tile([1.0, damping, kT], (n, 1)),
                 np.tile([1.0, 1.0, 0.0, 0.0], (n, 1)))
    o = oracle_lib.ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    o.set_noise(workloads.noise_table(steps, n, dim, seed))
    for it in range(1, steps + 1):
        o.step(1)
        if it > 1:
            p = o.get_agents()[0]
            sq = ((p - L / 2.0) ** 2).sum(axis=1)
            t = it * dt
            e = np.exp(-damping * t)
            expected = -dim * kT / damping ** 2 * (1.0 - e) * (3.0 - e) + 2.0 * dim * kT * t / damping
            assert abs(expected - sq.mean()) <= 3.5 * sq.std()


def test_determinism(oracle_lib):
    # tests/determinism_mechanics_interaction.rs:44-62: repeated runs are bitwise identical,
    # with one and with several worker threads
    w = workloads.cell_sorting(n_cells=600, domain_size=90.0, n_threads=4)
    outs = []
    for threads in (1, 4, 4):
        o = oracle_lib.ChiliOracle(w.config(), n_threads=threads)
        o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
        o.step(30)
        outs.append(o.get_agents())
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            np.testing.assert_array_equal(a, b)


def test_voxel_index_semantics(oracle_lib):
    # cartesian_cuboid_n.rs:469-490 + num-traits to_usize: (-1,0) -> 0, NaN -> error,
    # pos == max -> index n (no upper bound check); setup twin rejects pos <= min (:320-337)
    w = workloads.cell_sorting()
    o = oracle_lib.ChiliOracle(w.config())
    dx = 110.0 / 6.0
    assert o.voxel_index([0.0, 0.0, 0.0]).tolist() == [0, 0, 0]
    assert o.voxel_index([-0.5 * dx, 1.0, 1.0]).tolist() == [0, 0, 0]
    assert o.voxel_index([-1.5 * dx, 1.0, 1.0]) is None
    assert o.voxel_index([110.0, 1.0, 1.0]).tolist() == [6, 0, 0]
    assert o.voxel_index([float("nan"), 1.0, 1.0]) is None
    assert o.voxel_index([0.0, 1.0, 1.0], setup_variant=True) is None
    assert o.voxel_index([1e-300, 1.0, 1.0], setup_variant=True).tolist() == [0, 0, 0]
    for k in range(1, 6):
        face = k * dx
        q = int((face - 0.0) / dx)
        assert o.voxel_index([face, 1.0, 1.0]).tolist()[0] == q


def test_neighbor_sets(oracle_lib):
    # cartesian_cuboid_n.rs:641-662: clipped 3^D box, self excluded, ascending plain order
    w = workloads.cell_sorting()
    o = oracle_lib.ChiliOracle(w.config())
    assert len(o.neighbors(0)) == 7
    assert len(o.neighbors((1 * 6 + 1) * 6 + 1)) == 26
    nb = o.neighbors((2 * 6 + 3) * 6 + 5)
    assert len(nb) == 17 and np.all(np.diff(nb.astype(np.int64)) > 0)
    # subdomain chunks of the shipped example: 216 voxels on 4 threads -> 54 each
    assert [o.sub_of(v) for v in (0, 53, 54, 107, 108, 215)] == [0, 0, 1, 1, 2, 3]


def test_pair_weights_chili_as_written(oracle_lib):
    """SURVEY.md finding 2: same voxel 1, other voxel same subdomain 1/2, other subdomain 1."""
    from cellular_raza_b200.workloads import Workload

    def run(pos, nsub, mode=_abi.CRB_WEIGHT_CHILI):
        cfg = dict(dim=3, mechanics_kind=_abi.CRB_MECH_NEWTON_DAMPED, interaction_kind=_abi.CRB_INT_CELL_SORTING,
                   domain_min=[0.0] * 3, domain_max=[72.0] * 3, n_voxels=[4] * 3, dx=[18.0] * 3, dt=0.1,
                   ref_n_subdomains=nsub, pair_weight_mode=mode)
        w = Workload("pair", cfg, np.asarray(pos), np.zeros((2, 3)), np.tile([2.0, 1.0], (2, 1)),
                     np.tile([6.0, 2.0, 1.5], (2, 1)), np.array([1, 1], dtype=np.uint8))
        o = oracle_lib.ChiliOracle(w.config())
        o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
        f, _ = o.forces()
        ids = o.get_agents()[2]
        return f[np.argsort(ids)]

    f_pair, _ = oracle_lib.pair_force(_abi.CRB_INT_CELL_SORTING, [10.0, 10.0, 10.0], [10.0, 10.0, 17.0],
                                      [6.0, 2.0, 1.5], 1, 6.0, 1)
    same_voxel = run([[10.0, 10.0, 10.0], [10.0, 10.0, 17.0]], 1)
    np.testing.assert_allclose(same_voxel[0], f_pair, rtol=1e-15)
    # z = 17 and z = 24 are in voxels 0 and 1 along the last axis, distance 7 again
    f_pair2, _ = oracle_lib.pair_force(_abi.CRB_INT_CELL_SORTING, [10.0, 10.0, 17.0], [10.0, 10.0, 24.0],
                                       [6.0, 2.0, 1.5], 1, 6.0, 1)
    cross_same_sub = run([[10.0, 10.0, 17.0], [10.0, 10.0, 24.0]], 1)
    np.testing.assert_allclose(cross_same_sub[0], 0.5 * f_pair2, rtol=1e-15)
    physical = run([[10.0, 10.0, 17.0], [10.0, 10.0, 24.0]], 1, _abi.CRB_WEIGHT_PHYSICAL)
    np.testing.assert_allclose(physical[0], f_pair2, rtol=1e-15)
    # 64 subdomains of one voxel each: the two voxels are in different subdomains -> weight 1
    cross_other_sub = run([[10.0, 10.0, 17.0], [10.0, 10.0, 24.0]], 64)
    np.testing.assert_allclose(cross_other_sub[0], f_pair2, rtol=1e-15)
    np.testing.assert_allclose(cross_other_sub[1], -f_pair2, rtol=1e-15)


def test_ab_wrapper_recurrence(oracle_lib):
    """SURVEY.md finding 3: push-before-read + oldest-first iteration => step 0 is AB2 with
    [dx0, dx0], step k >= 1 is AB3 with [dx_k, dx_{k-1}, dx_k] (chili/solvers.rs:171-230)."""
    w = _free_motion_workload()
    w.mech_params[:] = [0.5, 1.0]  # damping 0.5: dv = -0.5 v
    o = oracle_lib.ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    x, v, dt = 0.5, 0.1, 0.1
    f0, f1, f2 = 23.0 / 12.0, -16.0 / 12.0, 5.0 / 12.0
    prev = None
    for k in range(5):
        dx, dv = v, 0.0 / 1.0 - 0.5 * v
        if k == 0:
            xn = ((dx * 1.5 + dx * (-0.5)) * dt + x) * 1.0 + 0.0 * dt
            vn = ((dv * 1.5 + dv * (-0.5)) * dt + v) * 1.0 + 0.0 * dt
        else:
            xn = ((dx * f0 + (prev[0] * f1 + dx * f2)) * dt + x) * 1.0 + 0.0 * dt
            vn = ((dv * f0 + (prev[1] * f1 + dv * f2)) * dt + v) * 1.0 + 0.0 * dt
        prev = (dx, dv)
        x, v = xn, vn
        o.step(1)
        pos, vel, ids, _ = o.get_agents()
        i = int(np.where(ids == 0)[0][0])
        assert pos[i, 0] == x and vel[i, 0] == v
    hx, hv, hl = o.get_aux()
    assert hl.tolist() == [2] * 10


def test_potentials_known_values(oracle_lib):
    # Morse: zero beyond cutoff and at distance 0 (interaction.rs:163-169); force at r = R1+R2 is 0
    fo, fe = oracle_lib.pair_force(_abi.CRB_INT_MORSE, [0, 0, 0.0], [0, 0, 6.0], [3.0, 0.5, 12.0, 0.01], 0, 3.0)
    assert np.all(fo == 0) and np.all(fe == 0)
    fo, _ = oracle_lib.pair_force(_abi.CRB_INT_MORSE, [0, 0, 0.0], [0, 0, 13.0], [3.0, 0.5, 12.0, 0.01], 0, 3.0)
    assert np.all(fo == 0)
    fo, _ = oracle_lib.pair_force(_abi.CRB_INT_MORSE, [1, 1, 1.0], [1, 1, 1.0], [3.0, 0.5, 12.0, 0.01], 0, 3.0)
    assert np.all(fo == 0)
    fo, fe = oracle_lib.pair_force(_abi.CRB_INT_MORSE, [0, 0, 0.0], [0, 0, 4.0], [3.0, 0.5, 12.0, 0.01], 0, 3.0)
    e = math.exp(-0.5 * (4.0 - 6.0))
    assert fo[2] == pytest.approx(-1.0 * (-2.0 * 0.01 * 0.5 * e * (1 - e)), rel=1e-15)
    np.testing.assert_array_equal(fe, -fo)
    # Mie: CalcError at r == 0 (interaction.rs:363-367)
    with pytest.raises(oracle_lib.OracleError):
        oracle_lib.pair_force(_abi.CRB_INT_MIE, [1, 1, 1.0], [1, 1, 1.0], [1.0, 1.0, 5.0, 3.0, 4.0, 2.0], 0, 1.0)
    # BoundLennardJones: NaN at r == 0 (no guard, interaction.rs:122-138)
    fo, _ = oracle_lib.pair_force(_abi.CRB_INT_BOUND_LJ, [1, 1, 1.0], [1, 1, 1.0], [1.0, 1.0, 5.0, 3.0])
    assert np.all(np.isnan(fo))
    # cell sorting: coincident -> 0; between species only repulsion (main.rs:60-63, 88-96)
    fo, _ = oracle_lib.pair_force(_abi.CRB_INT_CELL_SORTING, [1, 1, 1.0], [1, 1, 1.0], [6.0, 2.0, 1.5], 0, 6.0, 0)
    assert np.all(fo == 0)
    fa, _ = oracle_lib.pair_force(_abi.CRB_INT_CELL_SORTING, [0, 0, 0.0], [0, 0, 15.0], [6.0, 2.0, 1.5], 0, 6.0, 0)
    fb, _ = oracle_lib.pair_force(_abi.CRB_INT_CELL_SORTING, [0, 0, 0.0], [0, 0, 15.0], [6.0, 2.0, 1.5], 0, 6.0, 1)
    assert fa[2] > 0 and np.all(fb == 0)  # attraction pulls own (+z) towards ext; other species: none
    fc, _ = oracle_lib.pair_force(_abi.CRB_INT_CELL_SORTING, [0, 0, 0.0], [0, 0, 8.0], [6.0, 2.0, 1.5], 0, 6.0, 1)
    assert fc[2] < 0  # repulsion acts between species
    fd, _ = oracle_lib.pair_force(_abi.CRB_INT_CELL_SORTING, [0, 0, 0.0], [0, 0, 18.5], [6.0, 2.0, 1.5], 0, 6.0, 0)
    assert np.all(fd == 0)  # beyond range * (R1 + R2) = 18
    # benchmark copy has the opposite sign (cellular_raza-benchmarks/src/cell_sorting.rs:79-86)
    fe_, _ = oracle_lib.pair_force(_abi.CRB_INT_CELL_SORTING_BENCH, [0, 0, 0.0], [0, 0, 15.0], [6.0, 2.0, 1.5], 0, 6.0, 0)
    np.testing.assert_array_equal(fe_, -fa)


def _spring_workload(n_extra: int = 0):
    """cellular_raza/tests/domain_force_spring.rs:76-117: ONE NewtonDamped2D particle (damping 0, mass 0.5)
    at x0 = 10 in [-77, 77]^2 with 3 x 3 voxels, dt 1e-3, aspects [Mechanics, DomainForce] with
    force_x = -0.0032 * (pos.x - 0)."""
    from cellular_raza_b200.workloads import Workload

    n = 1 + n_extra
    pos = np.zeros((n, 2))
    pos[:, 0] = 10.0 + 3.0 * np.arange(n)
    cfg = dict(dim=2, mechanics_kind=_abi.CRB_MECH_NEWTON_DAMPED, interaction_kind=_abi.CRB_INT_NONE,
               domain_min=[-77.0] * 2, domain_max=[77.0] * 2, n_voxels=[3] * 2, dx=[154.0 / 3.0] * 2, dt=1e-3,
               ref_n_subdomains=1, solver_order=2, domain_force_kind=_abi.CRB_DOMAIN_FORCE_SPRING,
               domain_force_strength=[0.0032, 0.0], domain_force_anchor=[0.0, 0.0])
    return Workload("spring", cfg, pos, np.zeros((n, 2)), np.tile([0.0, 0.5], (n, 1)), np.zeros((n, 0)))


def test_domain_force_spring(oracle_lib):
    # domain_force_spring.rs:107-116: |x(t) - x0 cos(omega t)| < 1e-3 at every saved iteration (every 100
    # steps) over 10,000 steps, omega = sqrt(k / m)
    w = _spring_workload()
    o = oracle_lib.ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    omega = np.sqrt(0.0032 / 0.5)
    f, _ = o.forces()
    assert f[0, 0] == -0.0032 * 10.0 and f[0, 1] == 0.0
    for it in range(100, 10_001, 100):
        o.step(100)
        x = o.get_agents()[0][0, 0]
        assert abs(x - 10.0 * np.cos(omega * it * 1e-3)) < 1e-3
