"""1-D building blocks (NewtonDamped1D, Brownian1D, Langevin1D) against the oracle, and the reference's own
statistical known-answer tests (cellular_raza/tests/brownian_diffusion_constant_approx.rs) run on the DEVICE
through the reference-shaped ``run_simulation`` with the device's counter-based (Philox) noise."""
import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads
from cellular_raza_b200 import backend as cr
from cellular_raza_b200.device import DeviceSimulation

import parity

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-12
TRAJ_TOL = 1e-9


# ---- 1-D parity ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_threads", [1, 3])
@pytest.mark.parametrize("kind", ["newton", "brownian", "langevin"])
def test_1d_morse_against_oracle(oracle_lib, crb_lib, kind, n_threads):
    """a 1-D line of agents (3 neighbour voxels): voxel keys and membership bit-exact, forces 1e-12,
    100-step trajectory 1e-9 (BASELINE.json north_star tolerances)."""
    L = 12.0 * 400
    n = 1400
    w = workloads.brownian_morse(n, L, dim=1, n_threads=n_threads, langevin=(kind == "langevin"))
    if kind == "newton":
        w.cfg_kwargs["mechanics_kind"] = _abi.CRB_MECH_NEWTON_DAMPED
        w.cfg_kwargs["noise_mode"] = _abi.CRB_NOISE_NONE
        w.mech_params = np.tile([0.4, 1.3], (n, 1))
        w.vel = np.random.default_rng(3).normal(0.0, 0.2, size=(n, 1))
    assert tuple(w.cfg_kwargs["n_voxels"][:1]) == (400,)
    dev, ora = parity.make_pair(w)
    if kind != "newton":
        z = workloads.noise_table(100, n, 1, 11)
        dev.set_noise(z)
        ora.set_noise(z)
    parity.assert_same_membership(dev, ora)
    err, _, fo = parity.force_errors(dev, ora)
    assert np.abs(fo).max() > 0  # the agents interact
    assert err.max() <= FORCE_TOL
    dev.step(100)
    ora.step(100)
    perr, verr, key_diff = parity.state_errors(dev, ora, L)
    assert perr <= TRAJ_TOL, perr
    assert verr <= 1e-9
    assert key_diff.sum() <= 1
    parity.assert_same_membership(dev, ora) if key_diff.sum() == 0 else None
    dev.close()
    ora.close()


def test_1d_reflective_boundary_and_migration(oracle_lib, crb_lib):
    """fast 1-D agents without interaction partners in range: reflection at both walls and voxel
    migration follow the oracle exactly (positions bitwise: no pair sums are involved)."""
    n, L = 300, 90.0
    rng = np.random.default_rng(5)
    pos = rng.uniform(0.5, L - 0.5, (n, 1))
    vel = rng.uniform(-30.0, 30.0, (n, 1))
    cfg = dict(dim=1, mechanics_kind=_abi.CRB_MECH_NEWTON_DAMPED, interaction_kind=_abi.CRB_INT_MORSE,
               domain_min=[0.0], domain_max=[L], n_voxels=[9], dx=[L / 9.0], dt=0.05, ref_n_subdomains=2)
    w = workloads.Workload("line", cfg, pos, vel, np.tile([0.0, 1.0], (n, 1)), np.tile([0.1, 1.0, 0.0, 0.0], (n, 1)))
    dev, ora = parity.make_pair(w)
    for _ in range(20):
        dev.step(5)
        ora.step(5)
        pd, vd, ids_d, xd = dev.get_agents()
        po, vo, ids_o, xo = ora.get_agents()
        a, b = np.argsort(ids_d), np.argsort(ids_o)
        np.testing.assert_array_equal(pd[a], po[b])
        np.testing.assert_array_equal(vd[a], vo[b])
        np.testing.assert_array_equal(xd[a], xo[b])
        np.testing.assert_array_equal(dev.get_voxel_counts(), ora.get_voxel_counts())
    dev.close()
    ora.close()


# ---- the reference's statistical tests on the device -----------------------------------------------
def _positions(acc):
    hist = acc.cells.load_all_elements()
    return {it: np.array([c["pos"] for c in cells.values()]) for it, cells in hist.items()}


BROWNIAN = {1: cr.Brownian1D, 2: cr.Brownian2D, 3: cr.Brownian3D}
LANGEVIN = {1: cr.Langevin1D, 2: cr.Langevin2D, 3: cr.Langevin3D}


@pytest.mark.parametrize("dim,diff,seed", [(3, 1.0, 1), (3, 0.5, 2), (3, 0.25, 3), (2, 1.0, 4), (2, 0.5, 5), (2, 0.25, 6),
                                           (1, 1.0, 7), (1, 0.5, 8), (1, 0.25, 9)])
def test_brownian_diffusion_constant_like_reference(crb_lib, dim, diff, seed):
    """brownian_diffusion_constant_approx.rs:8-36, 110-258 (diffusion_constant_brownian_{3d,2d,1d}_{1,2,3}):
    160 particles from the domain centre, 3^dim voxels, dt 1e-3, 200 steps saved every step, aspects
    [Mechanics]; |2 dim D t - MSD| <= 3.5 sigma at every saved iteration."""
    n, L, dt, steps = 160, 200.0, 1e-3, 200
    domain = cr.CartesianCuboid.from_boundaries_and_n_voxels([0.0] * dim, [L] * dim, [3] * dim)
    domain.rng_seed = seed
    agents = [cr.Agent(BROWNIAN[dim]([L / 2.0] * dim, diff, 1.0), None) for _ in range(n)]
    settings = cr.Settings(1, cr.FixedStepsize.from_partial_save_steps(0.0, dt, steps, 1),
                           cr.StorageBuilder(priority=["Memory"]))
    acc = cr.run_simulation(domain, agents, settings, aspects=["Mechanics"])
    pos = _positions(acc)
    assert len(pos) >= steps
    for it, p in pos.items():
        sq = ((p - L / 2.0) ** 2).sum(axis=1)
        assert abs(2.0 * dim * diff * it * dt - sq.mean()) <= 3.5 * sq.std() + 1e-300


@pytest.mark.parametrize("dim,diff,damping,seed", [(3, 80.0, 10.0, 1), (3, 40.0, 10.0, 2), (3, 20.0, 10.0, 3),
                                                   (3, 40.0, 1.0, 4), (3, 40.0, 0.1, 5), (2, 80.0, 10.0, 6),
                                                   (2, 40.0, 10.0, 7), (2, 20.0, 10.0, 8), (1, 40.0, 10.0, 9)])
def test_langevin_diffusion_constant_like_reference(crb_lib, dim, diff, damping, seed):
    """brownian_diffusion_constant_approx.rs:260-420 (diffusion_constant_langevin_*): free Langevin
    particles with v(0) = 0; MSD(t) of the Ornstein-Uhlenbeck process within 3.5 sigma for iteration > 1."""
    n, L, dt, steps = 160, 200.0, 1e-3, 200
    kT = diff * damping * 1.0
    domain = cr.CartesianCuboid.from_boundaries_and_n_voxels([0.0] * dim, [L] * dim, [3] * dim)
    domain.rng_seed = seed
    agents = [cr.Agent(LANGEVIN[dim]([L / 2.0] * dim, [0.0] * dim, 1.0, damping, kT), None) for _ in range(n)]
    settings = cr.Settings(1, cr.FixedStepsize.from_partial_save_steps(0.0, dt, steps, 1),
                           cr.StorageBuilder(priority=["Memory"]))
    acc = cr.run_simulation(domain, agents, settings, aspects=["Mechanics"])
    for it, p in _positions(acc).items():
        if it <= 1:
            continue
        sq = ((p - L / 2.0) ** 2).sum(axis=1)
        t = it * dt
        e = np.exp(-damping * t)
        expected = -dim * kT / damping ** 2 * (1.0 - e) * (3.0 - e) + 2.0 * dim * kT * t / damping
        assert abs(expected - sq.mean()) <= 3.5 * sq.std()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_device_noise_is_standard_normal(crb_lib, dim):
    """a sharper check of the device generator than the reference's 3.5-sigma-of-the-population bound:
    40,000 free Brownian walkers, MSD within 5 standard errors of 2 dim D t after 50 steps, per-axis mean
    displacement within 5 standard errors of 0, and no correlation between axes."""
    n, L, dt, steps, diff = 40_000, 400.0, 1e-3, 50, 0.7
    nv = {1: 400, 2: 40, 3: 10}[dim]
    cfg = dict(dim=dim, mechanics_kind=_abi.CRB_MECH_BROWNIAN, interaction_kind=_abi.CRB_INT_NONE,
               domain_min=[0.0] * dim, domain_max=[L] * dim, n_voxels=[nv] * dim, dx=[L / nv] * dim, dt=dt,
               ref_n_subdomains=1, noise_mode=_abi.CRB_NOISE_PHILOX, rng_seed=123)
    pos0 = np.random.default_rng(9).uniform(0.25 * L, 0.75 * L, (n, dim))  # far from the reflecting walls
    w = workloads.Workload("walkers", cfg, pos0, np.zeros((n, dim)), np.tile([diff, 1.0], (n, 1)), np.zeros((n, 0)))
    dev = DeviceSimulation(w.config())
    dev.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, None)
    dev.step(steps)
    p, _, ids, _ = dev.get_agents()
    dev.close()
    d = p[np.argsort(ids)] - pos0
    t = steps * dt
    sq = (d ** 2).sum(axis=1)
    assert abs(sq.mean() - 2.0 * dim * diff * t) <= 5.0 * sq.std() / np.sqrt(n)
    sigma = np.sqrt(2.0 * diff * t)
    assert np.abs(d.mean(axis=0)).max() <= 5.0 * sigma / np.sqrt(n)
    if dim > 1:
        c = np.corrcoef(d.T)
        assert np.abs(c - np.eye(dim)).max() <= 5.0 / np.sqrt(n)
