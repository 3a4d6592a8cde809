"""Device vs oracle for the other building-block agents named in BASELINE.json north_star:
Brownian / Langevin mechanics driven by a shared noise table, Morse (config 3),
BoundLennardJones and Mie potentials, in 2-D and 3-D, plus the error paths that abort a chili run."""
import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads
from cellular_raza_b200.device import DeviceSimulation

import parity

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-12
TRAJ_TOL = 1e-9


def _with_noise(w, n_steps, seed=7):
    z = workloads.noise_table(n_steps, w.n, w.cfg_kwargs["dim"], seed)
    dev, ora = parity.make_pair(w)
    dev.set_noise(z)
    ora.set_noise(z)
    return dev, ora


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("n_threads", [1, 3])
def test_c3_brownian_morse(oracle_lib, crb_lib, dim, n_threads):
    L = 12.0 * (9 if dim == 3 else 20)
    w = workloads.brownian_morse(3000, L, dim=dim, n_threads=n_threads)
    dev, ora = _with_noise(w, 100)
    parity.assert_same_membership(dev, ora)
    err, _, _ = parity.force_errors(dev, ora)
    assert err.max() <= FORCE_TOL
    dev.step(100)
    ora.step(100)
    perr, verr, key_diff = parity.state_errors(dev, ora, L)
    assert perr <= TRAJ_TOL, perr
    assert verr == 0.0  # Brownian agents have no velocity (mechanics.rs:371-378)
    assert key_diff.sum() <= 1
    assert dev.n_agents == w.n


def test_brownian_philox_mode_runs_and_is_reproducible(crb_lib):
    w = workloads.brownian_morse(2000, 12.0 * 8, noise_mode=_abi.CRB_NOISE_PHILOX)
    outs = []
    for _ in range(2):
        d = DeviceSimulation(w.config())
        d.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
        d.step(50)
        outs.append(d.get_agents())
        d.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    # the walk moved the agents: MSD ~ 2 * dim * D * t (brownian_diffusion_constant_approx.rs:110-124)
    pos, _, ids, _ = outs[0]
    disp = pos[np.argsort(ids)] - w.pos
    msd = (disp ** 2).sum(axis=1).mean()
    expect = 2 * 3 * 0.3 * 50 * 0.01
    assert 0.7 * expect < msd < 1.4 * expect


def test_langevin_morse(oracle_lib, crb_lib):
    w = workloads.brownian_morse(2500, 12.0 * 8, langevin=True, n_threads=2)
    dev, ora = _with_noise(w, 100)
    err, _, _ = parity.force_errors(dev, ora)
    assert err.max() <= FORCE_TOL
    dev.step(100)
    ora.step(100)
    perr, verr, key_diff = parity.state_errors(dev, ora, 96.0)
    assert perr <= TRAJ_TOL and verr <= 1e-9
    assert key_diff.sum() <= 1


def _generic(kind, inter_row, n=2500, L=60.0, dim=3, cutoff=6.0, seed=11, mech=_abi.CRB_MECH_NEWTON_DAMPED,
             n_threads=2, min_sep=None):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(1e-3, L, size=(n, dim))
    if min_sep:  # thin out close pairs so that the run stays in the physical regime
        from scipy.spatial import cKDTree

        bad = {j for _, j in cKDTree(pos).query_pairs(min_sep)}
        pos = pos[[k for k in range(n) if k not in bad]]
        n = len(pos)
    mn, mx = [0.0] * dim, [L] * dim
    nv, dx = workloads._domain_from_interaction_range(mn, mx, cutoff)
    cfg = dict(dim=dim, mechanics_kind=mech, interaction_kind=kind, domain_min=mn, domain_max=mx,
               n_voxels=nv, dx=dx, dt=0.002, solver_order=2, ref_n_subdomains=n_threads)
    return workloads.Workload(name="generic", cfg_kwargs=cfg, pos=pos, vel=np.zeros((n, dim)),
                              mech_params=np.tile([1.0, 1.0], (n, 1)),
                              inter_params=np.tile(inter_row, (n, 1)), species=None)


@pytest.mark.parametrize("dim", [2, 3])
def test_bound_lennard_jones(oracle_lib, crb_lib, dim):
    # [epsilon, sigma, bound, cutoff]  (interaction.rs:122-138)
    w = _generic(_abi.CRB_INT_BOUND_LJ, [0.5, 2.0, 4.0, 6.0], dim=dim, n=2500 if dim == 3 else 800, min_sep=1.2)
    dev, ora = parity.make_pair(w)
    err, _, _ = parity.force_errors(dev, ora)
    assert err.max() <= FORCE_TOL
    dev.step(60)
    ora.step(60)
    perr, verr, key_diff = parity.state_errors(dev, ora, 60.0)
    assert perr <= TRAJ_TOL and key_diff.sum() <= 1


def test_mie_potential(oracle_lib, crb_lib):
    # [radius, strength, bound, cutoff, en, em]  (interaction.rs:345-389)
    w = _generic(_abi.CRB_INT_MIE, [1.0, 0.3, 5.0, 6.0, 4.0, 2.0], min_sep=0.9)
    dev, ora = parity.make_pair(w)
    err, _, _ = parity.force_errors(dev, ora)
    assert err.max() <= 2e-12  # pow() on both sides
    dev.step(40)
    ora.step(40)
    perr, _, key_diff = parity.state_errors(dev, ora, 60.0)
    assert perr <= TRAJ_TOL and key_diff.sum() <= 1


def test_mie_coincident_agents_is_calc_error(oracle_lib, crb_lib):
    w = _generic(_abi.CRB_INT_MIE, [1.0, 0.3, 5.0, 6.0, 4.0, 2.0], n=200)
    w.pos[17] = w.pos[3]
    dev = DeviceSimulation(w.config())
    dev.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, None)
    with pytest.raises(_abi.SimulationError) as e:
        dev.step(1)
    assert e.value.code == _abi.CRB_ERR_CALC
    from oracle.chili_oracle import ChiliOracle, OracleError

    ora = ChiliOracle(w.config())
    ora.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, None)
    with pytest.raises(OracleError):
        ora.step(1)


def test_lj_coincident_agents_is_nan_like_reference(oracle_lib, crb_lib):
    w = _generic(_abi.CRB_INT_BOUND_LJ, [0.5, 2.0, 4.0, 6.0], n=300, min_sep=1.2)
    w.pos[5] = w.pos[9]
    dev, ora = parity.make_pair(w)
    fd = dev.debug_forces()[np.argsort(dev.get_agents()[2])]
    fo = ora.forces()[0][np.argsort(ora.get_agents()[2])]
    assert np.isnan(fd[5]).all() and np.isnan(fd[9]).all()
    assert np.array_equal(np.isnan(fd), np.isnan(fo))


def test_agent_leaving_domain_twice_is_boundary_error(crb_lib):
    """one reflection per axis; still outside => BoundaryError (cartesian_cuboid_n.rs:616-625)."""
    w = workloads.cell_sorting(n_cells=50, domain_size=54.0)
    w.vel[0] = [2000.0, 0.0, 0.0]
    dev = DeviceSimulation(w.config())
    dev.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    with pytest.raises(_abi.SimulationError) as e:
        dev.step(1)
    assert e.value.code == _abi.CRB_ERR_BOUNDARY


def test_setup_rejects_agents_outside(crb_lib):
    w = workloads.cell_sorting(n_cells=50, domain_size=54.0)
    w.pos[3, 1] = 0.0  # pos <= min is rejected by check_min_max (cartesian_cuboid_n.rs:149-162, 321)
    dev = DeviceSimulation(w.config())
    with pytest.raises(_abi.SimulationError) as e:
        dev.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    assert e.value.code == _abi.CRB_ERR_BOUNDARY
    w.pos[3, 1] = 60.0  # beyond max: no voxel with that index (IndexError)
    with pytest.raises(_abi.SimulationError) as e:
        dev.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    assert e.value.code == _abi.CRB_ERR_INDEX


def test_reflective_boundary_matches(oracle_lib, crb_lib):
    w = workloads.cell_sorting(n_cells=300, domain_size=54.0, n_threads=2)
    rng = np.random.default_rng(3)
    w.vel = rng.normal(0, 15.0, w.vel.shape)  # many agents hit the walls
    dev, ora = parity.make_pair(w)
    dev.step(40)
    ora.step(40)
    perr, verr, key_diff = parity.state_errors(dev, ora, 54.0)
    assert perr <= TRAJ_TOL and verr <= 1e-9 and not key_diff.any()
    pos = dev.get_agents()[0]
    assert (pos >= 0).all() and (pos <= 54.0).all()


def test_empty_population_and_empty_voxels(crb_lib):
    w = workloads.cell_sorting(n_cells=3, domain_size=180.0)  # 1000 voxels, 3 agents
    dev = DeviceSimulation(w.config())
    dev.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    dev.step(5)
    assert dev.n_agents == 3
    dev.set_agents(w.pos[:0], w.vel[:0], w.mech_params[:0], w.inter_params[:0], w.species[:0])
    dev.step(2)
    assert dev.n_agents == 0 and dev.get_agents()[0].shape == (0, 3)
    dev.close()


@pytest.mark.parametrize("n_threads", [1, 4])
def test_domain_force_with_interaction(oracle_lib, crb_lib, n_threads):
    """aspects [Mechanics, Interaction, DomainForce]: the spring of domain_force_spring.rs:56-68 (per-axis
    strengths, anchor at the domain centre) on top of the cell_sorting pair forces — forces 1e-12 (the
    tile kernels add it in their fused epilogue), 100-step trajectory 1e-9, cell lists bit-exact."""
    w = workloads.cell_sorting(n_threads=n_threads)
    kw = dict(domain_force_kind=_abi.CRB_DOMAIN_FORCE_SPRING, domain_force_strength=[0.02, 0.0, 0.035],
              domain_force_anchor=[55.0, 55.0, 55.0])
    dev, ora = parity.make_pair(w, n_threads_oracle=n_threads, **kw)
    err, fd, fo = parity.force_errors(dev, ora)
    assert err.max() <= FORCE_TOL
    plain, _ = parity.make_pair(w, n_threads_oracle=n_threads)[0], None
    f0 = plain.debug_forces()[np.argsort(plain.get_agents()[2])]
    assert np.abs(fd - f0).max() > 1e-3, "the domain force did not act"
    plain.close()
    dev.step(100)
    ora.step(100)
    perr, _, key_diff = parity.state_errors(dev, ora, 110.0)
    assert perr <= 1e-9 and not key_diff.any()
    parity.assert_same_membership(dev, ora)
    dev.close()
    ora.close()


def test_no_interaction_aspect(oracle_lib, crb_lib):
    """CRB_INT_NONE + spring: several free particles oscillate independently; device == oracle to 1e-12."""
    from test_oracle_known_answers import _spring_workload

    w = _spring_workload(n_extra=7)
    dev, ora = parity.make_pair(w)
    assert dev.force_kernel_name == "k_force_integrate"
    assert np.array_equal(dev.debug_forces()[np.argsort(dev.get_agents()[2])],
                          ora.forces()[0][np.argsort(ora.get_agents()[2])])
    dev.step(500)
    ora.step(500)
    perr, verr, key_diff = parity.state_errors(dev, ora, 154.0)
    assert perr <= 1e-13 and verr <= 1e-13 and not key_diff.any()
    dev.close()
    ora.close()


def test_set_iteration_rewinds_the_device_generator(crb_lib):
    """CRB_NOISE_PHILOX is keyed by (rng_seed, agent id, iteration): crb_set_agents keeps the step counter (a host that
    re-uploads its population every step must keep drawing fresh numbers), crb_set_iteration puts it back — the same
    population from the same counter repeats its trajectory bit for bit, from another counter it does not."""
    w = workloads.brownian_morse(20_000, 12.0 * 13, noise_mode=_abi.CRB_NOISE_PHILOX)
    sim = DeviceSimulation(w.config(rng_seed=11))
    sim.set_uniform_params(w.mech_params[0], w.inter_params[0])

    def run(start_iteration=None):
        sim.set_agents(w.pos, w.vel, None, None, w.species)
        if start_iteration is not None:
            sim.set_iteration(start_iteration)
        it0 = sim.iteration
        sim.step(6)
        pos, _, ids, _ = sim.get_agents()
        return it0, pos[np.argsort(ids)]

    it_a, a = run(0)
    it_b, b = run()          # the counter went on: other numbers
    it_c, c = run(0)         # rewound: the first run again
    assert (it_a, it_b, it_c) == (0, 6, 0)
    assert np.array_equal(a, c) and not np.array_equal(a, b)
    sim.close()
