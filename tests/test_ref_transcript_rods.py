"""CPU: the C++ rods oracle (oracle/rods_oracle.cpp) against the independent Python transcription of the Rust
(tests/ref_transcript_rods.py, which never reads oracle/): the rod primitives on random inputs, and the chili step for
rod agents — forces with the 1 / 0.5 / 1 pair weights over emulated subdomains, neighbour counts, the Adams–Bashforth
wrapper, growth — on a small population.  Agreement pins what the doctests of the reference do not: the vertex–segment
force distribution of RodInteraction, the spring / curvature increment, the division geometry on bent rods."""
import numpy as np
import pytest

import ref_transcript as R
import ref_transcript_rods as T
from cellular_raza_b200 import _abi, workloads
from oracle import rods_oracle as ro

UM = 1e-6
V = 5


def _bent_rod(rng, seg=3 * UM, origin=(30 * UM, 7 * UM, 20 * UM), jitter=0.5):
    pos = np.zeros((V, 3))
    pos[0] = origin
    phi, th = rng.uniform(0, 2 * np.pi), rng.uniform(-0.3, 0.3)
    for i in range(1, V):
        phi += rng.uniform(-jitter, jitter)
        th += rng.uniform(-0.2, 0.2)
        pos[i] = pos[i - 1] + seg * rng.uniform(0.7, 1.3) * np.array([np.cos(phi) * np.cos(th), np.sin(phi) * np.cos(th), np.sin(th)])
    return pos


def test_nearest_point_and_pair_force_primitives():
    rng = np.random.default_rng(1)
    inter = [3 * UM, 0.5 / UM, 5 * UM, 0.1 * UM**2]
    morse = R.MorsePotential(*inter)
    ri = T.RodInteraction(morse, V)
    worst = 0.0
    for _ in range(40):
        a = _bent_rod(rng)
        b = _bent_rod(rng, origin=(30 * UM + rng.uniform(-6, 6) * UM, 7 * UM + rng.uniform(-6, 6) * UM, 20 * UM + rng.uniform(-3, 3) * UM))
        d_o, x_o, t_o = ro.nearest_point(a[2], b[1], b[2])
        d_t, x_t, t_t = T.nearest_point_from_point_to_line(list(a[2]), (list(b[1]), list(b[2])))
        assert d_o == d_t and t_o == t_t and list(x_o) == x_t
        fo_o, fe_o = ro.pair_force(a, b, inter, inter[0])
        fo_t, fe_t = ri.calculate_force_between(T.flat(a.tolist()), [0.0] * 15, T.flat(b.tolist()), [0.0] * 15, inter[0])
        fo_t, fe_t = np.array(fo_t).reshape(V, 3), np.array(fe_t).reshape(V, 3)
        scale = max(np.abs(fo_t).max(), np.abs(fe_t).max(), 1e-300)
        worst = max(worst, np.abs(fo_o - fo_t).max() / scale, np.abs(fe_o - fe_t).max() / scale)
    assert worst <= 1e-13, worst


def test_increment_primitive():
    rng = np.random.default_rng(2)
    for _ in range(40):
        pos, vel = _bent_rod(rng), rng.normal(0, 1e-6, (V, 3))
        force = rng.normal(0, 1e-6, (V, 3))
        mech = [0.0, 10.0, 1.0 * UM, rng.uniform(2.5, 3.5) * UM, 1.5, 0.1 * UM, 6 * UM, 0.1 * UM]
        dx_o, dv_o = ro.increment(pos, vel, mech, force)
        m = T.RodMechanics(T.flat(pos.tolist()), T.flat(vel.tolist()), mech[0], mech[1], mech[2], mech[3], mech[4], V)
        dx_t, dv_t = m.calculate_increment(T.flat(force.tolist()))
        assert np.array_equal(dx_o.reshape(-1), np.array(dx_t))
        dv_t = np.array(dv_t).reshape(V, 3)
        assert np.abs(dv_o - dv_t).max() <= 1e-13 * np.abs(dv_t).max()


def test_divide_primitive_on_bent_rods():
    rng = np.random.default_rng(3)
    for _ in range(40):
        pos = _bent_rod(rng, jitter=0.8)
        length, radius = rng.uniform(5.0, 6.5) * UM, 3 * UM
        p1_o, p2_o, nl_o = ro.divide(pos, length, radius)
        m = T.RodMechanics(T.flat(pos.tolist()), [0.0] * 15, 0.0, 10.0, 1e-6, length, 1.5, V)
        c1, c2, nl_t = m.divide(radius)
        assert nl_o == nl_t
        assert np.abs(p1_o - np.array(c1)).max() <= 1e-15 and np.abs(p2_o - np.array(c2)).max() <= 1e-15


def _population(n_threads, order=2, mode=_abi.CRB_WEIGHT_CHILI, diffusion=0.0):
    w = workloads.bacterial_rods(2, 2, rods_per_tile=7, n_threads=n_threads, solver_order=order, pair_weight_mode=mode,
                                 spring_length_range=(2.0, 3.5))
    mech = w.mech_params.copy()
    mech[:, 0] = diffusion
    return w, mech


def _transcript(w, mech, n_threads, order):
    kw = w.cfg_kwargs
    cub = R.CartesianCuboid.from_boundaries_and_n_voxels(list(kw["domain_min"]), list(kw["domain_max"]), [int(x) for x in kw["n_voxels"]])
    t = T.RodsTranscript(cub, V, n_threads, kw["dt"], order)
    cells, states = [], {}
    for k in range(w.n):
        mp, ip = mech[k], w.inter_params[k]
        m = T.RodMechanics(T.flat(w.pos[k].tolist()), T.flat(w.vel[k].tolist()), mp[0], mp[1], mp[2], mp[3], mp[4], V)
        cells.append(R.Cell(k, m, T.RodInteraction(R.MorsePotential(*ip), V)))
        states[k] = T.RodCellState(mp[5], mp[6], mp[7])
    t.set_cells(cells)
    return t, states


def _oracle_by_key(o):
    pos, vel, ident, parent, vox = o.get_agents()
    assert (ident[:, 0] == 1).all()  # Initial(k): no division in these tests
    return ident[:, 1].astype(int), pos, vel, vox


@pytest.mark.parametrize("n_threads", [1, 2, 4])
def test_forces_and_neighbor_counts_of_the_chili_step(n_threads):
    w, mech = _population(n_threads)
    o = ro.RodsOracle(w.config(), n_threads=2)
    o.set_agents(w.pos, w.vel, mech, w.inter_params)
    fo, ab, nbo = o.forces()
    keys, *_ = _oracle_by_key(o)
    t, _ = _transcript(w, mech, n_threads, 2)
    ft = t.forces()
    nbt = t.neighbor_counts()
    assert nbo.sum() > 50
    worst = 0.0
    for row, k in enumerate(keys):
        f_t, abs_sum = ft[k]
        f_t = np.array(f_t).reshape(V, 3)
        den = max(np.abs(f_t).max(), abs_sum, 1e-300)
        worst = max(worst, np.abs(fo[row] - f_t).max() / den)
        assert int(nbo[row]) == nbt[k]
    assert worst <= 1e-12, worst


@pytest.mark.parametrize("order", [2, 1, 0])
def test_three_steps_with_growth(order):
    w, mech = _population(3, order)
    o = ro.RodsOracle(w.config(), n_threads=2)
    o.set_agents(w.pos, w.vel, mech, w.inter_params)
    t, states = _transcript(w, mech, 3, order)
    L = max(w.cfg_kwargs["domain_max"])
    for _ in range(3):
        o.step(1)
        t.step(states=states)
        keys, pos, vel, vox = _oracle_by_key(o)
        mp = o.get_params()
        got = {c.key: (c, p) for p, c, _ in t.cells_in_order()}
        for row, k in enumerate(keys):
            c, plain = got[k]
            assert np.abs(pos[row].reshape(-1) - np.array(c.mech.pos)).max() / L <= 1e-12
            assert np.abs(vel[row].reshape(-1) - np.array(c.mech.vel)).max() <= 1e-12 * max(np.abs(vel).max(), 1e-300)
            assert int(vox[row]) == plain
            assert mp[row, 5] == states[k].growth_rate
            assert abs(mp[row, 3] - c.mech.spring_length) <= 1e-15 * 6e-6


def test_random_contribution_of_the_step():
    w, mech = _population(2, 0, diffusion=0.5 * UM)
    z = np.random.default_rng(4).standard_normal((2, w.n, 15))
    o = ro.RodsOracle(w.config(noise_mode=_abi.CRB_NOISE_TABLE), n_threads=2)
    o.set_agents(w.pos, w.vel, mech, w.inter_params)
    o.set_noise(z)
    t, states = _transcript(w, mech, 2, 0)
    L = max(w.cfg_kwargs["domain_max"])
    for s in range(2):
        o.step(1)
        t.step(noise_row=lambda key, s=s: list(z[s, key]), states=states)
    keys, pos, vel, vox = _oracle_by_key(o)
    got = {c.key: c for _, c, _ in t.cells_in_order()}
    for row, k in enumerate(keys):
        assert np.abs(pos[row].reshape(-1) - np.array(got[k].mech.pos)).max() / L <= 1e-12


@pytest.mark.parametrize("seed", range(6))
def test_random_rod_populations(seed):
    """randomly drawn rod populations (seed of the generator, 1 - 3 tiles per axis, 4 - 9 rods per tile, 1 - 6 emulated
    subdomains, solver orders 0 / 1 / 2): forces, neighbour counts and two growing steps, oracle == transcript"""
    rng = np.random.default_rng(500 + seed)
    n_threads, order = int(rng.integers(1, 7)), int(rng.integers(0, 3))
    w = workloads.bacterial_rods(int(rng.integers(1, 4)), int(rng.integers(1, 4)), rods_per_tile=int(rng.integers(4, 10)),
                                 seed=600 + seed, n_threads=n_threads, solver_order=order, spring_length_range=(2.0, 3.5))
    mech = w.mech_params.copy()
    o = ro.RodsOracle(w.config(), n_threads=2)
    o.set_agents(w.pos, w.vel, mech, w.inter_params)
    t, states = _transcript(w, mech, n_threads, order)
    fo, ab, nbo = o.forces()
    keys, *_ = _oracle_by_key(o)
    ft, nbt = t.forces(), t.neighbor_counts()
    for row, k in enumerate(keys):
        f_t, abs_sum = ft[k]
        f_t = np.array(f_t).reshape(V, 3)
        assert np.abs(fo[row] - f_t).max() / max(np.abs(f_t).max(), abs_sum, 1e-300) <= 1e-12
        assert int(nbo[row]) == nbt[k]
    L = max(w.cfg_kwargs["domain_max"])
    for _ in range(2):
        o.step(1)
        t.step(states=states)
    keys, pos, vel, vox = _oracle_by_key(o)
    got = {c.key: (c, p) for p, c, _ in t.cells_in_order()}
    for row, k in enumerate(keys):
        c, plain = got[k]
        assert np.abs(pos[row].reshape(-1) - np.array(c.mech.pos)).max() / L <= 1e-12
        assert int(vox[row]) == plain
