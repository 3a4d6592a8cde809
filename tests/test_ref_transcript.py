"""CPU: the C++ oracle against the reference-derived fixtures (the independent Python transcription of the Rust,
tests/ref_transcript.py), and the transcript against its own committed fixtures (stale-fixture guard) and against
the closed forms of the reference's potentials."""
import math

import numpy as np
import pytest

import ref_derived
import ref_transcript as R
from oracle.chili_oracle import ChiliOracle


@pytest.mark.parametrize("name", ref_derived.names())
def test_oracle_equals_reference_transcript(oracle_lib, name):
    """forces (1e-12), trajectories, voxel keys, in-voxel save order, ring fill and migration sets of every
    fixture: the two restatements were written independently (the transcript never reads oracle/), so agreement
    pins pair forces, the 1 / 0.5 / 1 chili weights, the AB recurrence and the local-then-remote mover order."""
    fix = ref_derived.load(name)
    ora = ChiliOracle(ref_derived.config(fix))
    try:
        ferr, serr = ref_derived.check(ora, fix)
    finally:
        ora.close()
    assert ferr <= 1e-12 and serr <= 1e-11


def test_fixture_is_what_the_transcript_computes_now():
    """regenerates one fixture in memory: the committed JSON must be the transcript's current output"""
    import importlib.util
    import sys
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("make_ref_derived", Path(__file__).resolve().parents[1] / "tools" / "make_ref_derived.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["make_ref_derived"] = mod
    spec.loader.exec_module(mod)
    fix = ref_derived.load("cell_sorting_3d_order1_t3")
    c = fix["config"]
    dom = R.CartesianCuboid.from_boundaries_and_interaction_range(c["domain_min"], c["domain_max"], 18.0)
    cells = []
    a = fix["agents"]
    for i, (p, v, s, mp, ip) in enumerate(zip(a["pos"], a["vel"], a["species"], a["mech_params"], a["inter_params"])):
        cells.append(R.Cell(i, R.NewtonDamped(list(p), list(v), mp[0], mp[1]), R.CellSpecificInteraction(s, *ip)))
    t = R.ChiliTranscript(dom, c["n_threads"], c["dt"], c["solver_order"])
    t.set_cells(cells)
    f0 = t.forces()
    assert [f0[i][0] for i in range(len(cells))] == fix["forces_t0"]
    t.step()
    cp = fix["checkpoints"][0]
    assert cp["step"] == 1
    got = {cell.key: list(cell.mech.pos) for _, cell, _ in t.cells_in_order()}
    assert [got[i] for i in range(len(cells))] == cp["pos"]
    assert [cell.key for _, cell, _ in t.cells_in_order()] == cp["save_order"]


# ---- closed forms of the potentials (what the Rust expressions reduce to) ---------------------------------------
def test_cell_sorting_closed_form_and_branches():
    """cell_sorting/src/main.rs:43-97: beyond the clamp the force on `own` is -dir * s(sigma) with
    s = strength * (sigma^-2 - sigma^-4) for equal species, and only the repelling part (s < 0) otherwise."""
    own = R.CellSpecificInteraction(1, 6.0, 2.0, 1.5)
    for r, same in [(5.0, True), (11.0, True), (13.0, True), (17.9, True), (11.0, False), (14.0, False)]:
        f_own, f_ext = own.calculate_force_between([r, 0.0, 0.0], None, [0.0, 0.0, 0.0], None, (6.0, 1 if same else 0))
        sig = r / 12.0
        bound = 4.0 + 1.0 / sig
        s = 2.0 * min(max(sig ** -2 - sig ** -4, -bound), bound)
        expect = -s if (same or s < 0) else 0.0
        assert f_own[0] == pytest.approx(expect, rel=1e-14, abs=1e-300)
        assert f_ext[0] == pytest.approx(-expect, rel=1e-14, abs=1e-300)
        assert f_own[1] == 0.0 and f_own[2] == 0.0
    # clamp region (:51-52, 64-68): r is replaced by 0.316 R, dir is not
    rc = 6.0 * 0.3162277660168379
    f_a, _ = own.calculate_force_between([0.5, 0.0, 0.0], None, [0.0, 0.0, 0.0], None, (6.0, 1))
    f_b, _ = own.calculate_force_between([1.5, 0.0, 0.0], None, [0.0, 0.0, 0.0], None, (6.0, 1))
    assert f_a == f_b
    sig = rc / 12.0
    assert f_a[0] == pytest.approx(-2.0 * max(sig ** -2 - sig ** -4, -(4.0 + 1.0 / sig)), rel=1e-14)
    # coincident positions: zero (:60-63); cutoff: r == range * (R1 + R2) still interacts (signum(+0) = 1)
    assert own.calculate_force_between([1.0, 2.0, 3.0], None, [1.0, 2.0, 3.0], None, (6.0, 1)) == ([0.0] * 3, [0.0] * 3)
    at, _ = own.calculate_force_between([18.0, 0.0, 0.0], None, [0.0, 0.0, 0.0], None, (6.0, 1))
    out, _ = own.calculate_force_between([math.nextafter(18.0, 19.0), 0.0, 0.0], None, [0.0, 0.0, 0.0], None, (6.0, 1))
    assert at[0] != 0.0 and out[0] == 0.0
    # the benchmark copy swaps the signs (cellular_raza-benchmarks/src/cell_sorting.rs:79-86)
    bench = R.CellSpecificInteraction(1, 6.0, 2.0, 1.5, bench=True)
    fo, fe = own.calculate_force_between([9.0, 1.0, 0.0], None, [0.0, 0.0, 0.0], None, (6.0, 1))
    bo, be = bench.calculate_force_between([9.0, 1.0, 0.0], None, [0.0, 0.0, 0.0], None, (6.0, 1))
    assert bo == fe and be == fo


def test_morse_lj_mie_closed_forms():
    m = R.MorsePotential(3.0, 0.5, 12.0, 0.01)
    fo, fe = m.calculate_force_between([7.0, 0.0], None, [0.0, 0.0], None, 3.0)
    e = math.exp(-0.5 * (7.0 - 6.0))
    assert fo[0] == pytest.approx(-2.0 * 0.01 * 0.5 * e * (1.0 - e), rel=1e-15) and fe[0] == -fo[0]
    assert m.calculate_force_between([12.5, 0.0], None, [0.0, 0.0], None, 3.0)[0] == [0.0, 0.0]
    lj = R.BoundLennardJones(0.5, 3.0, 4.0, 10.0)
    fo, fe = lj.calculate_force_between([4.0, 0.0, 0.0], None, [0.0, 0.0, 0.0], None, None)
    val = 4.0 * 0.5 / 4.0 * (12.0 * 0.75 ** 11 - 6.0 * 0.75 ** 5)
    assert fo[0] == pytest.approx(-min(4.0 / 4.0, val), rel=1e-14) and fe[0] == -fo[0]
    mie = R.MiePotential(2.0, 0.3, 5.0, 10.0, 4.0, 2.0)
    with pytest.raises(R.CalcError):
        mie.calculate_force_between([1.0, 1.0, 1.0], None, [1.0, 1.0, 1.0], None, 2.0)
    fo, _ = mie.calculate_force_between([5.0, 0.0, 0.0], None, [0.0, 0.0, 0.0], None, 2.0)
    sigma = (2.0 / 4.0) ** (1.0 / 2.0) * 4.0
    c = 4.0 / 2.0 * (4.0 / 2.0) ** (2.0 / 2.0)
    assert fo[0] == pytest.approx(min(c * 0.3 * (4.0 / sigma * (sigma / 5.0) ** 5 - 2.0 / sigma * (sigma / 5.0) ** 3), 5.0), rel=1e-14)


def test_pair_weights_of_chili_as_written():
    """update_mechanics.rs:186-234: same voxel -> both halves; other voxel, same subdomain -> the return force is
    dropped (weight 1/2); other subdomain -> PosInformation / ForceInformation deliver both halves."""
    def run(n_threads, pb):
        dom = R.CartesianCuboid.from_boundaries_and_n_voxels([0.0, 0.0], [20.0, 40.0], [1, 4])
        cells = [R.Cell(0, R.NewtonDamped([5.0, 8.0], [0.0, 0.0], 0.0, 1.0), R.MorsePotential(3.0, 0.5, 10.0, 1.0)),
                 R.Cell(1, R.NewtonDamped(pb, [0.0, 0.0], 0.0, 1.0), R.MorsePotential(3.0, 0.5, 10.0, 1.0))]
        t = R.ChiliTranscript(dom, n_threads, 0.1, 2)
        t.set_cells(cells)
        return t.forces()
    full, _ = R.MorsePotential(3.0, 0.5, 10.0, 1.0).calculate_force_between([5.0, 8.0], None, [5.0, 9.5], None, 3.0)
    f = run(1, [5.0, 9.5])            # same voxel (y index 0): weight 1
    assert f[0][0] == pytest.approx(full, rel=1e-15)
    full2, _ = R.MorsePotential(3.0, 0.5, 10.0, 1.0).calculate_force_between([5.0, 8.0], None, [5.0, 12.0], None, 3.0)
    f = run(1, [5.0, 12.0])           # neighbouring voxel, one subdomain: weight 1/2
    assert f[0][0][1] == pytest.approx(0.5 * full2[1], rel=1e-15)
    f = run(4, [5.0, 12.0])           # neighbouring voxel in ANOTHER subdomain: weight 1
    assert f[0][0][1] == pytest.approx(full2[1], rel=1e-15)
    assert f[1][0][1] == pytest.approx(-full2[1], rel=1e-15)


def test_adams_bashforth_wrapper_recurrence():
    """solvers.rs:171-230 with the ring pushed before it is read (aux_storage.rs:862-868): step 0 is
    AB2 over [dx0, dx0], step k >= 1 is AB3 over [dx_k, dx_{k-1}, dx_k]"""
    dom = R.CartesianCuboid.from_boundaries_and_n_voxels([0.0], [100.0], [1])

    class NoInter:
        def get_interaction_information(self):
            return None

        def calculate_force_between(self, *a):
            return [0.0], [0.0]

    c = R.Cell(0, R.NewtonDamped([10.0], [3.0], 0.7, 1.0), NoInter())
    t = R.ChiliTranscript(dom, 1, 0.1, 2)
    t.set_cells([c])
    x, v, dt = 10.0, 3.0, 0.1
    prev = None
    for k in range(4):
        dx, dv = v, 0.0 / 1.0 - 0.7 * v
        if prev is None:
            x_new = ((dx * 1.5 + dx * -0.5) * dt + x) * 1.0 + 0.0 * dt
            v_new = ((dv * 1.5 + dv * -0.5) * dt + v) * 1.0 + 0.0 * dt
        else:
            f0, f1, f2 = 23.0 / 12.0, -16.0 / 12.0, 5.0 / 12.0
            x_new = ((dx * f0 + (prev[0] * f1 + dx * f2)) * dt + x) * 1.0 + 0.0 * dt
            v_new = ((dv * f0 + (prev[1] * f1 + dv * f2)) * dt + v) * 1.0 + 0.0 * dt
        prev = (dx, dv)
        x, v = x_new, v_new
        t.step()
        assert c.mech.pos == [x] and c.mech.vel == [v], k


def test_geometry_pins_of_the_reference():
    """cartesian_cuboid_n.rs:12-18 (the rows of that comment table that satisfy the invariants the reference's own
    test asserts, :670-699: n + m = n_regions, n avg + m (avg - 1) = n_voxel), 167-182, 380-402"""
    assert [R.get_decomp_res(*a) for a in [(10, 3), (13, 4), (225, 16), (225, 17)]] == [
        (1, 2, 4), (1, 3, 4), (1, 15, 15), (4, 13, 14)]
    for n_voxel in range(1, 300, 7):
        for n_regions in range(1, 60, 3):
            n, m, avg = R.get_decomp_res(n_voxel, n_regions)
            assert n + m == n_regions and n * avg + m * (avg - 1) == n_voxel
    d = R.CartesianCuboid.from_boundaries_and_interaction_range([2.0, 3.0, 1.0], [10.0, 10.0, 20.0], 2.0)
    assert d.n_voxels == [4, 3, 9]
    d = R.CartesianCuboid.from_boundaries_and_n_voxels([0.0] * 3, [100.0] * 3, [5, 5, 5])
    subs = d.create_subdomains(4)
    assert len(subs) == 4 and sum(len(v) for _, v in subs) == 125
    assert d.get_all_voxel_indices()[:3] == [(0, 0, 0), (0, 0, 1), (0, 0, 2)]


# ---- randomised sweep: every potential, oracle (C++) against the transcript (Python) ------------------------------
@pytest.mark.parametrize("dim", [2, 3])
def test_pair_forces_of_every_potential_random_sweep(oracle_lib, dim):
    """2,000 random pairs per potential and dimension — distances from deep inside the clamp region to beyond the
    cutoff, equal and different species, both sign conventions of the species potential — through
    ``oracle_pair_force`` and through the transcript: the two restatements agree to 1e-13 of the pair force on both
    agents (pow / exp of libm on both sides)."""
    from oracle.chili_oracle import pair_force

    rng = np.random.default_rng(20 + dim)
    worst = 0.0
    for _ in range(2000):
        d = rng.standard_normal(dim)
        d /= np.linalg.norm(d)
        own = rng.uniform(-50.0, 50.0, dim)
        cases = []
        # species potential (cell_sorting/src/main.rs:43-97 and the benchmark copy)
        rad, rad_ext = rng.uniform(3.0, 8.0, 2)
        strength, rel = rng.uniform(0.5, 3.0), rng.uniform(1.1, 2.0)
        so, se = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        r = rng.choice([rng.uniform(1e-3, 0.4 * rad), rng.uniform(0.4 * rad, rel * (rad + rad_ext)),
                        rng.uniform(rel * (rad + rad_ext), 3.0 * rel * (rad + rad_ext))])
        for kind, bench in ((0, False), (1, True)):
            cases.append((kind, r, [rad, strength, rel], so, rad_ext, se,
                          R.CellSpecificInteraction(so, rad, strength, rel, bench=bench), (rad_ext, se)))
        # Morse (interaction.rs:151-260)
        mr, ms, mc, mst = rng.uniform(1.0, 4.0), rng.uniform(0.1, 1.0), rng.uniform(8.0, 14.0), rng.uniform(0.01, 1.0)
        cases.append((2, rng.uniform(0.05, 16.0), [mr, ms, mc, mst], 0, rad_ext, 0, R.MorsePotential(mr, ms, mc, mst), rad_ext))
        # BoundLennardJones (interaction.rs:112-145)
        le, lsg, lb, lc = rng.uniform(0.1, 2.0), rng.uniform(1.0, 4.0), rng.uniform(1.0, 20.0), rng.uniform(6.0, 12.0)
        cases.append((3, rng.uniform(0.3, 14.0), [le, lsg, lb, lc], 0, 0.0, 0, R.BoundLennardJones(le, lsg, lb, lc), None))
        # Mie (interaction.rs:330-389)
        en = rng.uniform(3.0, 12.0)
        em = rng.uniform(1.0, en - 0.5)
        ir, ist, ib, ic = rng.uniform(1.0, 4.0), rng.uniform(0.05, 1.0), rng.uniform(1.0, 20.0), rng.uniform(6.0, 14.0)
        cases.append((4, rng.uniform(0.3, 16.0), [ir, ist, ib, ic, en, em], 0, rad_ext, 0,
                      R.MiePotential(ir, ist, ib, ic, en, em), rad_ext))
        for kind, dist, params, s_own, ext_radius, s_ext, model, ext_info in cases:
            ext = own - dist * d
            fo, fe = pair_force(kind, own, ext, params, s_own, ext_radius, s_ext)
            to, te = model.calculate_force_between(list(own), None, list(ext), None, ext_info)
            to, te = np.array(to), np.array(te)
            scale = max(np.abs(to).max(), np.abs(te).max(), 1e-300)
            worst = max(worst, np.abs(fo - to).max() / scale, np.abs(fe - te).max() / scale)
            assert (to == 0.0).all() == (fo == 0.0).all()  # the cutoff decides the same way on both sides
    assert worst <= 1e-13, worst


def test_geometry_random_sweep_oracle_against_transcript(oracle_lib):
    """random cuboids in 1 - 3 dimensions with 1 - 6 subdomains: voxel index of random positions (faces, the lower
    and upper corner, outside, NaN; per-step and setup variants), neighbour lists (ascending plain index) and the
    subdomain of every voxel — C++ oracle == transcript, exactly."""
    from cellular_raza_b200 import _abi
    from cellular_raza_b200.device import make_config

    rng = np.random.default_rng(77)
    ran = 0
    for trial in range(60):
        dim = int(rng.integers(1, 4))
        nv = [int(x) for x in rng.integers(1, 7, dim)]
        mn = [float(x) for x in rng.uniform(-20.0, 5.0, dim)]
        mx = [mn[i] + float(rng.uniform(3.0, 40.0)) for i in range(dim)]
        n_threads = int(rng.integers(1, 7))
        n_vox = int(np.prod(nv))
        if R.get_decomp_res(n_vox, n_threads) is None:
            continue
        dom = R.CartesianCuboid.from_boundaries_and_n_voxels(mn, mx, nv)
        cfg = make_config(dim=dim, mechanics_kind=_abi.CRB_MECH_NEWTON_DAMPED, interaction_kind=_abi.CRB_INT_MORSE,
                          domain_min=mn, domain_max=mx, n_voxels=np.array(nv, dtype=np.uint64), dx=dom.dx, dt=0.1,
                          capacity=16, ref_n_subdomains=n_threads)
        ora = ChiliOracle(cfg)
        try:
            strides = [int(np.prod(nv[i + 1:])) for i in range(dim)]
            plain = lambda idx: sum(int(a) * s for a, s in zip(idx, strides))
            # subdomain of every voxel (create_subdomains: contiguous chunks of the enumeration)
            sub_t = {}
            for k, (_key, voxels) in enumerate(dom.create_subdomains(n_threads)):
                for v in voxels:
                    sub_t[plain(v)] = k
            subs_o = [ora.sub_of(p) for p in range(n_vox)]
            # (chunk keys are consecutive from 0, so the k-th group is subdomain k)
            assert subs_o == [sub_t[p] for p in range(n_vox)]
            # neighbours
            for idx in dom.get_all_voxel_indices():
                want = sorted(plain(q) for q in dom.get_neighbor_voxel_indices(idx))
                assert list(ora.neighbors(plain(idx))) == want
            # voxel index of positions
            pts = [list(rng.uniform(np.array(mn) - 1.0, np.array(mx) + 1.0)) for _ in range(30)]
            pts += [list(mn), list(mx), [np.nextafter(a, -np.inf) for a in mn], [np.nextafter(a, np.inf) for a in mn],
                    [mn[i] + dom.dx[i] * float(rng.integers(0, nv[i] + 1)) for i in range(dim)], [float("nan")] * dim]
            for p in pts:
                for setup in (False, True):
                    try:
                        want = (dom.get_voxel_index_of_raw if setup else dom.get_index_of)(p)
                    except R.BoundaryError:
                        want = None
                    got = ora.voxel_index(p, setup_variant=setup)
                    if want is None:
                        assert got is None, (p, setup, got)
                    else:
                        assert got is not None and tuple(int(x) for x in got) == want, (p, setup, got, want)
        finally:
            ora.close()
        ran += 1
    assert ran >= 40


def _make_ref_derived():
    import importlib.util
    import sys
    from pathlib import Path

    if "make_ref_derived" in sys.modules:
        return sys.modules["make_ref_derived"]
    spec = importlib.util.spec_from_file_location("make_ref_derived", Path(__file__).resolve().parents[1] / "tools" / "make_ref_derived.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["make_ref_derived"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("seed", range(16))
def test_random_chili_steps_oracle_against_transcript(oracle_lib, seed):
    """whole steps of randomly drawn cases, built in memory by the fixture generator (nothing is written): dimension
    2 / 3, every mechanics and potential, 1 - 5 emulated subdomains, solver orders 0 / 1 / 2, heterogeneous
    parameters, noise tables — forces, trajectories, voxel keys, save order, ring fill and migration sets of the C++
    oracle against the transcript"""
    import random

    mod = _make_ref_derived()
    rng = random.Random(1000 + seed)
    dim = rng.choice([2, 3])
    inter = rng.choice([
        {"kind": "cell_sorting", "radius": 6.0, "strength": 2.0, "range": 1.5},
        {"kind": "cell_sorting_bench", "radius": 6.0, "strength": 2.0, "range": 1.5},
        {"kind": "morse", "radius": 3.0, "stiffness": 0.5, "cutoff": 12.0, "strength": 0.5},
        {"kind": "bound_lj", "epsilon": 0.5, "sigma": 3.0, "bound": 4.0, "cutoff": 10.0},
        {"kind": "mie", "radius": 2.0, "strength": 0.3, "bound": 5.0, "cutoff": 10.0, "en": 4.0, "em": 2.0}])
    edge = {"cell_sorting": 18.0, "cell_sorting_bench": 18.0, "morse": 12.0, "bound_lj": 10.0, "mie": 10.0}[inter["kind"]]
    mech = rng.choice([{"kind": "newton", "damping": 1.2, "mass": 1.1},
                       {"kind": "brownian", "diffusion": 0.3, "kt": 0.3},
                       {"kind": "langevin", "mass": 1.5, "damping": 0.7, "kt": 0.4}])
    nv = [rng.randint(2, 4) for _ in range(dim)]
    mx = [edge * n + rng.uniform(0.0, 0.9 * edge) for n in nv]  # (from_boundaries_and_interaction_range floors)
    noise = mech["kind"] != "newton"
    dt = {"cell_sorting": 0.1, "cell_sorting_bench": 0.1, "morse": 0.02, "bound_lj": 0.005, "mie": 0.005}[inter["kind"]]
    fix = mod.build(f"random_{seed}", dim=dim, mn=[0.0] * dim, mx=mx, interaction_range=edge,
                    n_threads=rng.randint(1, 5), dt=dt, order=rng.choice([0, 1, 2]), mech=mech, inter=inter,
                    n_agents=rng.randint(50, 90), seed=seed, steps=[1, 2, 4], vel_scale=rng.choice([0.0, 5.0, 20.0]),
                    noise=noise, hetero=rng.random() < 0.5, write=False)
    ora = ChiliOracle(ref_derived.config(fix))
    try:
        ferr, serr = ref_derived.check(ora, fix)
    finally:
        ora.close()
    assert ferr <= 1e-12 and serr <= 1e-11
