"""Bacterial rods (BASELINE.json configs[3]): device vs oracle
through the C ABI.  Device order must equal chili's (voxel, in-voxel) order, so arrays are compared
position by position; identifiers, voxel assignment, neighbour counts, growth and division sets are
bit-exact, forces <= 1e-12 (metric of SURVEY §8d), trajectories <= 1e-9 of the domain extent."""
import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads
from cellular_raza_b200._abi import SimulationError
from cellular_raza_b200.rods import RodSimulation
from oracle.rods_oracle import RodsOracle

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-12
TRAJ_TOL = 1e-9


def make_pair(w, **over):
    dev = RodSimulation(w.config(**over))
    ora = RodsOracle(w.config(**over), n_threads=4)
    for sim in (dev, ora):
        sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    return dev, ora


def assert_same_population(dev, ora):
    pd, vd, idd, pad, xd = dev.get_agents()
    po, vo, ido, pao, xo = ora.get_agents()
    assert len(idd) == len(ido)
    np.testing.assert_array_equal(idd, ido)   # same identifiers in the same (voxel, in-voxel) order
    np.testing.assert_array_equal(pad, pao)
    np.testing.assert_array_equal(xd, xo)
    np.testing.assert_array_equal(dev.get_voxel_counts(), ora.get_voxel_counts())
    return pd, vd, po, vo


def force_check(dev, ora):
    fd = dev.debug_forces()
    nbd = dev.debug_neighbor_counts()
    fo, ab, nbo = ora.forces()
    np.testing.assert_array_equal(nbd, nbo)
    num = np.abs(fd - fo).reshape(len(fd), -1).max(axis=1)
    den = np.maximum(np.abs(fo).reshape(len(fo), -1).max(axis=1), ab)
    den = np.where(den == 0, 1.0, den)
    return (num / den).max(), int(nbo.sum())


@pytest.mark.parametrize("mode,n_threads", [(_abi.CRB_WEIGHT_CHILI, 8), (_abi.CRB_WEIGHT_CHILI, 1),
                                            (_abi.CRB_WEIGHT_PHYSICAL, 3)])
def test_forces_and_neighbor_counts(crb_lib, mode, n_threads):
    w = workloads.bacterial_rods(6, 5, rods_per_tile=24, n_threads=n_threads, pair_weight_mode=mode)
    dev, ora = make_pair(w)
    assert_same_population(dev, ora)
    err, nb = force_check(dev, ora)
    assert nb > 1000  # the population does interact
    assert err <= FORCE_TOL, err
    assert dev.count_candidate_pairs() == ora.count_candidate_pairs()


def test_example_as_shipped_100_steps(crb_lib):
    w = workloads.bacterial_rods(1, 1)  # 5 rods, [8, 1, 2] voxels, n_threads 8
    dev, ora = make_pair(w)
    dev.step(100)
    ora.step(100)
    pd, vd, po, vo = assert_same_population(dev, ora)
    L = max(w.cfg_kwargs["domain_max"])
    assert np.abs(pd - po).max() / L <= TRAJ_TOL
    np.testing.assert_array_equal(dev.get_params()[:, [3, 5]], ora.get_params()[:, [3, 5]])


@pytest.mark.parametrize("order", [2, 1, 0])
def test_trajectory_with_growth_and_division(crb_lib, order):
    # spring lengths close to the threshold (6 um): rods divide within the 100 steps
    w = workloads.bacterial_rods(6, 5, rods_per_tile=6, spring_length_range=(5.6, 5.99), solver_order=order,
                                 capacity_factor=3.0)
    dev, ora = make_pair(w)
    n_div = 0
    for _ in range(10):
        dev.step(10)
        ora.step(10)
        assert dev.n_agents == ora.n_agents
    pd, vd, po, vo = assert_same_population(dev, ora)
    n_div = dev.n_agents - w.n
    assert n_div > 10, "the test population must divide"
    L = max(w.cfg_kwargs["domain_max"])
    assert np.abs(pd - po).max() / L <= TRAJ_TOL
    # velocity error integrated over the run stays below the trajectory tolerance
    assert np.abs(vd - vo).max() * (100 * 0.1) <= TRAJ_TOL * L
    mpd, mpo = dev.get_params(), ora.get_params()
    np.testing.assert_array_equal(mpd[:, 5], mpo[:, 5])  # growth rates: pure function of the counts
    assert np.abs(mpd[:, 3] - mpo[:, 3]).max() <= 1e-15 * 6e-6 * 100
    hxd, hvd, hld = dev.get_aux()
    hxo, hvo, hlo = ora.get_aux()
    np.testing.assert_array_equal(hld, hlo)
    # newest position increments are velocities (dx = vel): same integrated bound as the velocities
    assert np.abs(hxd - hxo).max() * (100 * 0.1) <= TRAJ_TOL * L
    err, _ = force_check(dev, ora)
    # forces of the two evolved states (sanity: positions differ by ~1e-15 m, stiffness 5e5 / m)
    assert err <= 1e-6


def test_division_step_is_exact(crb_lib):
    """one step from identical state in which rods divide: geometry of mother and daughter, identifiers
    Division(voxel, counter), parents and the fresh AuxStorage of the daughter."""
    w = workloads.bacterial_rods(4, 4, rods_per_tile=5, spring_length_range=(5.995, 5.9999), capacity_factor=3.0)
    dev, ora = make_pair(w)
    dev.step(1)
    ora.step(1)
    assert ora.n_divisions > 0
    assert dev.n_divisions == ora.n_divisions
    pd, vd, po, vo = assert_same_population(dev, ora)
    L = max(w.cfg_kwargs["domain_max"])
    assert np.abs(pd - po).max() / L <= 1e-15
    _, _, ident, parent, _ = dev.get_agents()
    div = ident[:, 0] == 0
    assert div.sum() == 2 * dev.n_divisions
    assert np.all(parent[div, 0] == 1)  # parents are Initial(n)
    _, _, hl = dev.get_aux()
    assert set(hl[div].tolist()) == {0, 1}  # daughters start empty, mothers keep their ring buffer
    np.testing.assert_array_equal(dev.get_params()[:, 3], ora.get_params()[:, 3])


def test_bitwise_reruns(crb_lib):
    w = workloads.bacterial_rods(6, 5, rods_per_tile=6, spring_length_range=(5.6, 5.99), capacity_factor=3.0)
    outs = []
    for _ in range(3):
        d = RodSimulation(w.config())
        d.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
        d.step(60)
        outs.append(d.get_agents())
        d.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            np.testing.assert_array_equal(a, b)


def test_invariants_physical_mode(crb_lib):
    """PHYSICAL weights: interaction forces of a closed population sum to zero per axis."""
    w = workloads.bacterial_rods(4, 4, rods_per_tile=20, pair_weight_mode=_abi.CRB_WEIGHT_PHYSICAL)
    d = RodSimulation(w.config())
    d.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    f = d.debug_forces()
    scale = np.abs(f).sum()
    assert scale > 0
    assert np.abs(f.sum(axis=(0, 1))).max() <= 1e-12 * scale
    d.step(20)
    pos = d.get_agents()[0]
    for c in range(3):
        assert pos[:, :, c].min() >= 0 and pos[:, :, c].max() <= w.cfg_kwargs["domain_max"][c]


def test_vertex_counts(crb_lib):
    for V in (3, 4, 7):
        w = workloads.bacterial_rods(3, 3, rods_per_tile=12, n_vertices=V, spring_length_range=(2.5, 3.2))
        dev, ora = make_pair(w)
        err, _ = force_check(dev, ora)
        assert err <= FORCE_TOL
        dev.step(20)
        ora.step(20)
        pd, vd, po, vo = assert_same_population(dev, ora)
        assert np.abs(pd - po).max() / max(w.cfg_kwargs["domain_max"]) <= TRAJ_TOL


def test_error_paths(crb_lib):
    w = workloads.bacterial_rods(4, 4, rods_per_tile=5, spring_length_range=(5.995, 5.9999), capacity_factor=1.0)
    d = RodSimulation(w.config(capacity=w.n))
    d.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    with pytest.raises(SimulationError) as e:
        d.step(3)
    assert e.value.code == _abi.CRB_ERR_CAPACITY
    d.close()
    # a rod whose vertex mean lies at/below the domain minimum: BoundaryError at setup
    w = workloads.bacterial_rods(1, 1)
    bad = w.pos.copy()
    bad[0, :, 1] = -1e-6
    d = RodSimulation(w.config())
    with pytest.raises(SimulationError) as e:
        d.set_agents(bad, w.vel, w.mech_params, w.inter_params)
    assert e.value.code == _abi.CRB_ERR_BOUNDARY
    # non-uniform tension is rejected
    mp = w.mech_params.copy()
    mp[1, 1] *= 2
    with pytest.raises(SimulationError) as e:
        d.set_agents(w.pos, w.vel, mp, w.inter_params)
    assert e.value.code == _abi.CRB_ERR_INVALID_ARG
    d.close()
    with pytest.raises(SimulationError):
        RodSimulation(w.config(dim=2))


@pytest.mark.parametrize("n_threads", [1, 8])
def test_domain_force_gel_pressure_and_surface_friction(crb_lib, n_threads):
    """aspect DomainForce for rods (CRB_DOMAIN_FORCE_ROD_GEL = SubDomainForce of CartesianSubDomainRods,
    bacterial_rods.rs:433-477): gel pressure on every vertex, friction on moving vertices near a wall."""
    w = workloads.bacterial_rods(4, 4, rods_per_tile=24, n_threads=n_threads)
    um = workloads.MICRO_METRE
    w.vel = np.random.default_rng(8).normal(0.0, 0.3 * um, w.vel.shape)
    w.vel[::3] = 0.0  # rods at rest feel no friction
    over = dict(domain_force_kind=_abi.CRB_DOMAIN_FORCE_ROD_GEL,
                domain_force_strength=[0.02 * um, 0.4, 2.0 * um])
    dev, ora = make_pair(w, **over)
    plain = RodSimulation(w.config())
    plain.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    diff = dev.debug_forces() - plain.debug_forces()
    plain.close()
    assert np.abs(diff[..., 2]).min() > 0                       # every vertex is pushed down
    assert (np.abs(diff[..., 0]) > 0).any() and (diff[..., 0] == 0).any()  # friction only for some
    assert_same_population(dev, ora)
    err, _ = force_check(dev, ora)
    assert err <= FORCE_TOL, err
    dev.step(60)
    ora.step(60)
    pd, vd, po, vo = assert_same_population(dev, ora)
    L = max(w.cfg_kwargs["domain_max"])
    assert np.abs(pd - po).max() / L <= TRAJ_TOL
    assert np.abs(vd - vo).max() <= 1e-8 * np.abs(vo).max()
    dev.close()
    ora.close()


def test_domain_force_kind_mismatch_is_rejected(crb_lib):
    w = workloads.bacterial_rods(1, 1)
    with pytest.raises(SimulationError):
        RodSimulation(w.config(domain_force_kind=_abi.CRB_DOMAIN_FORCE_SPRING))


def _with_diffusion(w, diffusion):
    mech = w.mech_params.copy()
    mech[:, 0] = diffusion
    return mech


@pytest.mark.parametrize("order", [2, 0])
def test_random_contribution_noise_table(crb_lib, order):
    """RodMechanics::get_random_contribution (bacterial_rods.rs:215-240) with the shared noise table: device and
    oracle read the same standard normal numbers (column-major fill of the V x 3 matrix), also for daughters born
    inside the run (their table row is a hash of Division(voxel, counter))."""
    w = workloads.bacterial_rods(6, 5, rods_per_tile=6, spring_length_range=(5.6, 5.99), solver_order=order,
                                 capacity_factor=3.0)
    mech = _with_diffusion(w, 0.5e-6)
    z = np.random.default_rng(11).standard_normal((60, 257, 15))   # fewer rows than rods: rows wrap modulo n_ids
    dev = RodSimulation(w.config(noise_mode=_abi.CRB_NOISE_TABLE))
    ora = RodsOracle(w.config(noise_mode=_abi.CRB_NOISE_TABLE), n_threads=4)
    for sim in (dev, ora):
        sim.set_agents(w.pos, w.vel, mech, w.inter_params)
        sim.set_noise(z)
    # one step from identical state: the random displacement itself is exact
    dev.step(1)
    ora.step(1)
    pd, vd, po, vo = assert_same_population(dev, ora)
    L = max(w.cfg_kwargs["domain_max"])
    assert np.abs(pd - po).max() / L <= 1e-15
    for k in (10, 10, 10, 10, 10, 9):
        dev.step(k)
        ora.step(k)
        assert dev.n_agents == ora.n_agents
    pd, vd, po, vo = assert_same_population(dev, ora)
    assert dev.n_agents > w.n, "the test population must divide"
    assert np.abs(pd - po).max() / L <= TRAJ_TOL
    # table exhausted: the next step must fail loudly on both sides
    with pytest.raises(SimulationError):
        dev.step(1)


def test_random_contribution_changes_the_trajectory_and_needs_a_noise_source(crb_lib):
    w = workloads.bacterial_rods(2, 2)
    mech = _with_diffusion(w, 0.5e-6)
    plain = RodSimulation(w.config())
    with pytest.raises(SimulationError):   # diffusion_constant != 0 with noise_mode NONE is refused, not ignored
        plain.set_agents(w.pos, w.vel, mech, w.inter_params)
    plain.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    noisy = RodSimulation(w.config(noise_mode=_abi.CRB_NOISE_PHILOX, rng_seed=5))
    noisy.set_agents(w.pos, w.vel, mech, w.inter_params)
    plain.step(5)
    noisy.step(5)
    assert np.abs(plain.get_agents()[0] - noisy.get_agents()[0]).max() > 1e-9


def test_random_contribution_philox_statistics(crb_lib):
    """device generator: free rods far apart (no pair forces, relaxed springs, no growth): after one Euler step every
    vertex coordinate has moved by sqrt(dt) z / 4 * D with z standard normal: mean 0, variance 1, no correlation
    between vertices or coordinates."""
    V, D, dt = 5, 0.5e-6, 0.1
    nx, nz = 64, 32
    w = workloads.bacterial_rods(nx, nz, rods_per_tile=1, solver_order=0)
    n = w.n
    pos = np.zeros((n, V, 3))
    tile = np.array([200e-6, 15e-6, 45e-6])
    ix, iz = np.divmod(np.arange(n), nz)
    pos[:, :, 0] = (ix * tile[0] + 90e-6)[:, None] + 3e-6 * np.arange(V)[None, :]
    pos[:, :, 1] = 7.5e-6
    pos[:, :, 2] = (iz * tile[2] + 22e-6)[:, None]
    mech = _with_diffusion(w, D)
    mech[:, 3] = 3e-6
    mech[:, 5] = 0.0
    mech[:, 7] = 0.0
    dev = RodSimulation(w.config(noise_mode=_abi.CRB_NOISE_PHILOX, rng_seed=123))
    dev.set_agents(pos, np.zeros_like(pos), mech, w.inter_params)
    dev.step(1)
    p1, _, ident, _, _ = dev.get_agents()
    order = np.argsort(ident[:, 1])
    zhat = (p1[order] - pos) / (np.sqrt(dt) * 0.25 * D)          # n x V x 3 standard normal numbers
    flat = zhat.reshape(n, -1)
    m = flat.size
    assert abs(flat.mean()) < 5 / np.sqrt(m)
    assert abs(flat.var() - 1.0) < 5 * np.sqrt(2.0 / m)
    c = np.corrcoef(flat.T)
    off = c - np.eye(c.shape[0])
    assert np.abs(off).max() < 5 / np.sqrt(n)


@pytest.mark.parametrize("n_threads", [1, 4])
def test_device_equals_the_independent_transcription(crb_lib, n_threads):
    """the device against tests/ref_transcript_rods.py (a literal Python transcription of the Rust that never reads
    oracle/), without the oracle in the loop: forces <= 1e-12, neighbour counts exact, three steps with growth"""
    import ref_transcript as R
    import ref_transcript_rods as T

    V = 5
    w = workloads.bacterial_rods(2, 2, rods_per_tile=7, n_threads=n_threads, spring_length_range=(2.0, 3.5))
    kw = w.cfg_kwargs
    cub = R.CartesianCuboid.from_boundaries_and_n_voxels(list(kw["domain_min"]), list(kw["domain_max"]), [int(x) for x in kw["n_voxels"]])
    t = T.RodsTranscript(cub, V, n_threads, kw["dt"], 2)
    cells, states = [], {}
    for k in range(w.n):
        mp, ip = w.mech_params[k], w.inter_params[k]
        m = T.RodMechanics(T.flat(w.pos[k].tolist()), T.flat(w.vel[k].tolist()), mp[0], mp[1], mp[2], mp[3], mp[4], V)
        cells.append(R.Cell(k, m, T.RodInteraction(R.MorsePotential(*ip), V)))
        states[k] = T.RodCellState(mp[5], mp[6], mp[7])
    t.set_cells(cells)
    dev = RodSimulation(w.config())
    dev.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    keys = dev.get_agents()[2][:, 1].astype(int)
    fd, nbd = dev.debug_forces(), dev.debug_neighbor_counts()
    ft, nbt = t.forces(), t.neighbor_counts()
    assert nbd.sum() > 50
    for row, k in enumerate(keys):
        f_t, abs_sum = ft[k]
        f_t = np.array(f_t).reshape(V, 3)
        assert np.abs(fd[row] - f_t).max() / max(np.abs(f_t).max(), abs_sum, 1e-300) <= FORCE_TOL
        assert int(nbd[row]) == nbt[k]
    L = max(kw["domain_max"])
    for _ in range(3):
        dev.step(1)
        t.step(states=states)
    pd, vd, ident, _, vox = dev.get_agents()
    mp = dev.get_params()
    got = {c.key: (c, p) for p, c, _ in t.cells_in_order()}
    for row, k in enumerate(ident[:, 1].astype(int)):
        c, plain = got[k]
        assert np.abs(pd[row].reshape(-1) - np.array(c.mech.pos)).max() / L <= TRAJ_TOL
        assert int(vox[row]) == plain
        assert mp[row, 5] == states[k].growth_rate


def test_rods_async_pipeline_equals_synchronous_calls(crb_lib):
    """crb_set_agents_async + crb_step + crb_snapshot_async for CRB_MECH_ROD on two alternating handles (the rods e2e
    loop of bench.py: the next upload is queued before the step of the current handle; parameter rows checked on the
    device; clears and the division count as kernels) leave exactly the bits of crb_set_agents + crb_step +
    crb_get_agents — with rods that divide, so that the agent count changes between upload and snapshot."""
    import ctypes as C

    from cellular_raza_b200.device import _ptr

    w = workloads.bacterial_rods(4, 4, rods_per_tile=6, spring_length_range=(5.9, 5.99))
    n, V = w.n, 5
    n_steps = 6
    ref = RodSimulation(w.config())
    want = []
    pops = [(w.pos + k * 1e-9, w.vel) for k in range(4)]  # four populations that differ in every coordinate
    for k in range(4):
        ref.set_agents(pops[k][0], pops[k][1], w.mech_params, w.inter_params)
        ref.step(n_steps)
        pos, vel, ident, _, vox = ref.get_agents()
        want.append((pos, vel, ident[:, 1].copy(), vox))
    assert any(len(x[0]) > n for x in want), "the test needs divisions"
    ref.close()

    sims = [RodSimulation(w.config()) for _ in range(2)]
    lib = sims[0]._lib
    cap = 2 * n
    nn = C.c_uint64()
    hm = sims[0]._pinned("mp", (n, 8), np.float64)
    hi = sims[0]._pinned("ip", (n, 4), np.float64)
    hm[:], hi[:] = w.mech_params, w.inter_params
    bufs = []
    for s in sims:
        bufs.append(dict(in_p=s._pinned("in_pos", (n, V, 3), np.float64), in_v=s._pinned("in_vel", (n, V, 3), np.float64),
                         out_p=s._pinned("out_pos", (cap, V, 3), np.float64), out_v=s._pinned("out_vel", (cap, V, 3), np.float64),
                         out_i=s._pinned("out_ids", (cap,), np.uint64), out_x=s._pinned("out_vox", (cap,), np.uint64)))
    got = []
    for k in range(4):
        s, b = sims[k % 2], bufs[k % 2]
        b["in_p"][:], b["in_v"][:] = pops[k]
        s._check(lib.crb_set_agents_async(s._h, n, _ptr(b["in_p"]), _ptr(b["in_v"]), _ptr(hm), _ptr(hi), None, None))
        s._check(lib.crb_step_async(s._h, n_steps))
        s._check(lib.crb_snapshot_async(s._h, cap, _ptr(b["out_p"]), _ptr(b["out_v"]), _ptr(b["out_i"]), _ptr(b["out_x"]),
                                        None, None, None))
        s._check(lib.crb_snapshot_wait(s._h, C.byref(nn)))
        m = int(nn.value)
        assert m == s.n_agents
        got.append((b["out_p"][:m].copy(), b["out_v"][:m].copy(), b["out_i"][:m].copy(), b["out_x"][:m].copy()))
    for g, x in zip(got, want):
        for a, c in zip(g, x):
            assert np.array_equal(a, c)
    # the ring buffers of a rod are not part of a snapshot
    s = sims[0]
    assert lib.crb_snapshot_async(s._h, cap, None, None, None, None, _ptr(bufs[0]["out_p"]), None, None) != 0
    # a parameter row that differs (not spring_length / growth_rate) is found on the device and reported by the
    # next synchronising call
    hm[n // 2, 4] *= 2.0
    s._check(lib.crb_set_agents_async(s._h, n, _ptr(bufs[0]["in_p"]), _ptr(bufs[0]["in_v"]), _ptr(hm), _ptr(hi), None, None))
    with pytest.raises(Exception, match="uniform"):
        s.step(1)
    for s in sims:
        s.close()


def test_rod_wrapper_async_calls(crb_lib):
    """RodSimulation.set_agents_async / snapshot_async / snapshot_wait (the object wrappers over the calls of the test
    above) against the synchronous wrappers"""
    w = workloads.bacterial_rods(4, 4, rods_per_tile=6, spring_length_range=(5.9, 5.99))
    a, b = RodSimulation(w.config()), RodSimulation(w.config())
    a.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    b.set_agents_async(w.pos, w.vel, w.mech_params, w.inter_params)
    a.step(8)
    b.step(8)
    pos, vel, ident, _, vox = a.get_agents()
    sp, sv, si, sx = b.snapshot_wait(b.snapshot_async())
    assert len(pos) > w.n
    assert np.array_equal(sp, pos) and np.array_equal(sv, vel) and np.array_equal(si, ident[:, 1]) and np.array_equal(sx, vox)
    with pytest.raises(ValueError):
        b.snapshot_async(aux=True)
    a.close()
    b.close()
