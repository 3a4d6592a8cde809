"""run_simulation (reference-shaped entry point) on the device: the reference's own whole-loop checks."""
import json

import numpy as np
import pytest

from cellular_raza_b200 import backend as cr

pytestmark = pytest.mark.gpu


def test_free_motion_like_chili_doctest(crb_lib):
    """chili/mod.rs:134-157: no force, pos = key + iteration * dt * vel within 1e-5, 10 cells, every
    stored iteration holds all cells with Initial identifiers."""
    vel = 0.1
    agents = [cr.Agent(cr.NewtonDamped3D([1.0 + n, 5.0, 5.0], [vel, 0.0, 0.0], 0.0, 1.0),
                       cr.MorsePotential(0.1, 1.0, 0.05, 0.0)) for n in range(10)]
    domain = cr.CartesianCuboid.from_boundaries_and_n_voxels([0.0] * 3, [20.0, 10.0, 10.0], [4, 2, 2])
    settings = cr.Settings(1, cr.FixedStepsize.from_partial_save_interval(0.0, 0.1, 20.0, 2.0),
                           cr.StorageBuilder(priority=["Memory"]))
    acc = cr.run_simulation(domain, agents, settings, mechanics_solver_order=2)
    hist = acc.cells.load_all_elements()
    assert sorted(hist) == list(range(0, 201, 20))
    for it, cells in hist.items():
        assert len(cells) == 10
        for (kind, key), c in cells.items():
            assert kind == "Initial"
            assert abs(c["pos"][0] - (1.0 + key + it * 0.1 * vel)) < 1e-5


def test_cell_sorting_example_with_json_storage(crb_lib, tmp_path):
    """cell_sorting as shipped, shortened: JSON snapshots land in
    <location>/<date>/cells/json/<iteration:020>/batch_<instance:020>_<k:020>.json with
    {"data": [{"identifier", "element": [CellBox, AuxStorage]}]} (storage/concepts.rs:156-170, 702-743)."""
    rng = np.random.default_rng(1)
    n = 400
    pos = rng.uniform(0.5, 72.0, (n, 3))
    agents = [cr.Agent(cr.NewtonDamped3D(pos[k].tolist(), [0.0] * 3, 2.0, 1.0),
                       cr.CellSpecificInteraction(int(k <= n // 2), 6.0, 2.0, 1.5)) for k in range(n)]
    domain = cr.CartesianCuboid.from_boundaries_and_interaction_range([0.0] * 3, [72.0] * 3, 18.0)
    settings = cr.Settings(4, cr.FixedStepsize.from_partial_save_steps(0.0, 0.2, 40, 20),
                           cr.StorageBuilder(location=str(tmp_path), priority=["SerdeJson", "Memory"]))
    acc = cr.run_simulation(domain, agents, settings)
    its = acc.cells.get_all_iterations()
    assert its == [0, 20, 40]
    f = acc.path / "cells" / "json" / f"{20:020d}" / f"batch_{0:020d}_{0:020d}.json"
    assert f.exists()
    body = json.loads(f.read_text())
    e = body["data"][0]
    assert set(e) == {"identifier", "element"} and "Initial" in e["identifier"]
    cbox, aux = e["element"]
    assert set(cbox) == {"identifier", "parent", "cell"} and cbox["parent"] is None
    assert set(aux["mechanics"]) == {"positions", "velocities", "current_force", "zero_force"}
    assert len(aux["mechanics"]["positions"]) == 2  # ring of N = 2 increments, oldest first
    mem = acc.cells.memory[40]
    assert len(mem) == n
    moved = np.array([np.abs(mem[("Initial", k)]["pos"] - pos[k]).max() for k in range(n)])
    assert moved.max() > 1e-3  # the cells interact
    # ---- what cellular_raza-examples/cell_sorting/src/plot.py:14-57 does with the output directory ------------
    import glob

    def get_cells_at_iteration(iteration):
        cells = []
        for batch in glob.glob(str(acc.path / "cells" / "json" / f"{iteration:020d}") + "/*"):
            cells.extend(json.loads(open(batch).read())["data"])
        return cells

    cells = get_cells_at_iteration(40)
    assert len(cells) == n
    info = [(c["element"][0]["cell"]["mechanics"]["pos"], c["element"][0]["cell"]["interaction"]["cell_radius"],
             c["element"][0]["cell"]["interaction"]["species"]) for c in cells]
    assert all(len(p) == 3 and r == 6.0 and sp in ("RedCell", "BlueCell") for p, r, sp in info)
    assert sum(sp == "BlueCell" for _, _, sp in info) == n // 2 + 1
    one = cells[0]["element"][0]["cell"]
    assert set(one["mechanics"]) == {"pos", "vel", "damping_constant", "mass"}  # serde of NewtonDamped3D, mechanics.rs:54-65
    assert set(one["interaction"]) == {"species", "cell_radius", "potential_strength", "relative_interaction_range"}
    # one batch file per storage instance = chili subdomain (Settings.n_threads = 4), cells of a file lie in its voxels
    files = sorted((acc.path / "cells" / "json" / f"{40:020d}").glob("batch_*.json"))
    assert [f.name for f in files] == [f"batch_{i:020d}_{0:020d}.json" for i in range(4)]
    subs = []
    for single in sorted(glob.glob(str(acc.path / "subdomains" / "json" / f"{40:020d}") + "/*")):
        d = json.loads(open(single).read())
        subs.append(d["element"])
    assert len(subs) == 4
    assert subs[0]["domain_min"] == [0.0] * 3 and subs[0]["domain_max"] == [72.0] * 3  # plot.py:46-50
    assert sum(len(x["voxels"]) for x in subs) == 64 and subs[0]["domain_n_voxels"] == [4, 4, 4]
    for i, f in enumerate(files):
        vox = {tuple(v) for v in subs[i]["voxels"]}
        for c in json.loads(f.read_text())["data"]:
            p = c["element"][0]["cell"]["mechanics"]["pos"]
            assert tuple(int(x // 18.0) for x in p) in vox


def test_snapshot_async_equals_blocking_download(crb_lib):
    """crb_snapshot_async / crb_snapshot_wait (side-stream D2H into pinned memory) return exactly what
    crb_get_agents / crb_get_aux return at the same iteration, although the stepping loop went on
    while the copy was in flight."""
    from cellular_raza_b200 import workloads
    from cellular_raza_b200.device import DeviceSimulation

    w = workloads.cell_sorting_scaled(60_000, 7.1)
    sim = DeviceSimulation(w.config())
    sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    sim.step(3)
    ref = (*sim.get_agents(), *sim.get_aux())
    token = sim.snapshot_async()
    sim.step(5)  # overwrites every state buffer at least twice
    snap = sim.snapshot_wait(token)
    for a, b in zip(ref, snap):
        assert np.array_equal(a, b)
    # a second snapshot while one is pending waits for the first; wait without a snapshot is an error
    t1 = sim.snapshot_async(aux=False)
    now = sim.get_agents()
    got = sim.snapshot_wait(t1)
    assert all(np.array_equal(a, b) for a, b in zip(now, got))
    with pytest.raises(cr.SimulationError):
        sim.snapshot_wait(t1)
    sim.close()


@pytest.mark.parametrize("n_steps", [1, 7])
def test_async_pipeline_equals_synchronous_calls(crb_lib, n_steps):
    """crb_set_agents_async + crb_step_async + crb_snapshot_async on two alternating handles (the e2e loop of bench.py:
    transfers first, force pass on the low-priority stream, kernels instead of copy-engine clears; n_steps = 7 also
    replays the two-step CUDA graphs) leave exactly the bits of crb_set_agents + crb_step + crb_get_agents."""
    import ctypes as C

    from cellular_raza_b200 import workloads
    from cellular_raza_b200.device import DeviceSimulation, _ptr

    w = workloads.cell_sorting_scaled(120_000, 7.1)
    ref = DeviceSimulation(w.config())
    ref.set_uniform_params(w.mech_params[0], w.inter_params[0])
    pops = [(w.pos, w.vel)]
    want = []
    for _ in range(4):  # every population is the result of the previous one
        ref.set_agents(pops[-1][0], pops[-1][1], None, None, w.species)
        ref.step(n_steps)
        pos, vel, ids, _ = ref.get_agents()
        o = np.argsort(ids, kind="stable")
        want.append((pos[o], vel[o]))
        pops.append(want[-1])
    ref.close()

    sims = [DeviceSimulation(w.config()) for _ in range(2)]
    for s in sims:
        s.set_uniform_params(w.mech_params[0], w.inter_params[0])
    lib = sims[0]._lib
    n = w.n
    nn = C.c_uint64()
    species = sims[0]._pinned("sp", (n,), np.uint8)
    species[:] = w.species
    got = []
    for k in range(4):
        s = sims[k % 2]
        hin_p, hin_v = s._pinned("in_pos", (n, 3), np.float64), s._pinned("in_vel", (n, 3), np.float64)
        hin_p[:], hin_v[:] = pops[k]
        out_p, out_v = s._pinned("out_pos", (n, 3), np.float64), s._pinned("out_vel", (n, 3), np.float64)
        out_i = s._pinned("out_ids", (n,), np.uint64)
        s._check(lib.crb_set_agents_async(s._h, n, _ptr(hin_p), _ptr(hin_v), None, None, _ptr(species), None))
        s._check(lib.crb_step_async(s._h, n_steps))
        s._check(lib.crb_snapshot_async(s._h, n, _ptr(out_p), _ptr(out_v), _ptr(out_i), None, None, None, None))
        s._check(lib.crb_snapshot_wait(s._h, C.byref(nn)))
        s.sync()  # latched errors of the asynchronous calls surface here
        assert nn.value == n
        o = np.argsort(out_i, kind="stable")
        got.append((out_p[o].copy(), out_v[o].copy()))
    for (gp, gv), (wp, wv) in zip(got, want):
        assert np.array_equal(gp, wp) and np.array_equal(gv, wv)
    for s in sims:
        s.close()


def test_domain_force_spring_like_reference_test(crb_lib):
    """cellular_raza/tests/domain_force_spring.rs:76-117 through run_simulation with aspects
    [Mechanics, DomainForce]: x(t) = x0 cos(omega t) within 1e-3 at every saved iteration."""
    k, mass, dt, x0 = 0.0032, 0.5, 0.001, 10.0
    domain = cr.CartesianCuboid.from_boundaries_and_n_voxels([-77.0] * 2, [77.0] * 2, [3, 3])
    settings = cr.Settings(1, cr.FixedStepsize.from_partial_save_interval(0.0, dt, 10.0, 0.1),
                           cr.StorageBuilder(priority=["Memory"]))
    agents = [cr.Agent(cr.NewtonDamped2D([x0, 0.0], [0.0, 0.0], 0.0, mass))]
    acc = cr.run_simulation(domain, agents, settings, aspects=["Mechanics", "DomainForce"],
                            domain_force=cr.SpringDomainForce([k, 0.0]))
    hist = acc.cells.load_all_elements()
    assert len(hist) == 101
    omega = np.sqrt(k / mass)
    for it, cells in hist.items():
        (c,) = cells.values()
        assert abs(c["pos"][0] - x0 * np.cos(omega * it * dt)) < 1e-3
    with pytest.raises(cr.SimulationError):
        cr.run_simulation(domain, agents, settings, aspects=["Mechanics", "DomainForce"])


def test_bacterial_rods_example_through_run_simulation(crb_lib, tmp_path):
    """cellular_raza-examples/bacterial_rods/src/main.rs:111-212, shortened: rod agents with the four aspects
    of the example through the reference-shaped entry point; rods grow and divide (dynamic agent count,
    Division identifiers with parents), snapshots equal the oracle's state at the same iterations."""
    from cellular_raza_b200 import workloads
    from oracle.rods_oracle import RodsOracle

    w = workloads.bacterial_rods(2, 2, rods_per_tile=6, spring_length_range=(5.5, 5.99))
    um = cr.MICRO_METRE
    agents = []
    for k in range(w.n):
        mp, ip = w.mech_params[k], w.inter_params[k]
        mech = cr.RodMechanics(w.pos[k].tolist(), w.vel[k].tolist(), mp[0], mp[1], mp[2], mp[3], mp[4])
        agents.append(cr.RodAgent(mech, cr.RodInteraction(cr.MorsePotential(ip[0], ip[1], ip[2], ip[3])),
                                  growth_rate=mp[5], spring_length_threshold=mp[6], growth_base_rate=mp[7]))
    kw = w.cfg_kwargs
    domain = cr.CartesianCuboidRods(cr.CartesianCuboid.from_boundaries_and_n_voxels(
        kw["domain_min"], kw["domain_max"], [int(x) for x in kw["n_voxels"]]))
    settings = cr.Settings(8, cr.FixedStepsize.from_partial_save_steps(0.0, kw["dt"], 120, 40),
                           cr.StorageBuilder(location=str(tmp_path), priority=["SerdeJson", "Memory"]))
    acc = cr.run_simulation(domain, agents, settings, aspects=["Mechanics", "Interaction", "Cycle", "NeighborSensing"])
    assert acc.cells.get_all_iterations() == [0, 40, 80, 120]
    ora = RodsOracle(w.config(), n_threads=4)
    ora.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    L = max(kw["domain_max"])
    for it in (40, 80, 120):
        ora.step(40)
        po, vo, ido, pao, xo = ora.get_agents()
        mem = acc.cells.memory[it]
        assert len(mem) == len(ido)
        for a in range(len(ido)):
            c = mem[cr._ident(ido[a])]
            assert np.abs(c["pos"] - po[a]).max() / L <= 1e-9
            assert c["voxel"] == int(xo[a])
            assert c["parent"] == (None if int(pao[a, 0]) > 2 else cr._ident(pao[a]))
    last = acc.cells.memory[120]
    assert len(last) > w.n and any(k[0] == "Division" for k in last)  # rods divided
    f = acc.path / "cells" / "json" / f"{120:020d}" / f"batch_{0:020d}_{0:020d}.json"
    body = json.loads(f.read_text())
    e = next(x for x in body["data"] if "Division" in x["identifier"])
    cbox, aux = e["element"]
    assert set(cbox) == {"identifier", "parent", "cell"} and cbox["parent"] is not None
    assert set(cbox["cell"]) == {"mechanics", "interaction", "growth_rate", "spring_length_threshold"}
    assert cbox["cell"]["mechanics"]["pos"][1:] == [5, None] and len(cbox["cell"]["mechanics"]["pos"][0]) == 15
    assert set(aux) == {"cycle", "mechanics", "neighbor_sensing"}
    # ---- what cellular_raza-examples/bacterial_rods/plotter.py:22-51 does with the output directory -----------
    import glob

    import pandas as pd

    results = []
    for file in glob.glob(str(acc.path / "cells/json/{:020}".format(120)) + "/*"):
        elements = json.load(open(file))["data"]
        results.extend(element["element"][0] for element in elements)
    df = pd.json_normalize(results)
    assert len(df) == len(last)
    positions = df["cell.mechanics.pos"].apply(lambda x: np.array(x[0], dtype=float).reshape(3, -1))
    radii = np.array(list(df["cell.interaction.radius"]), dtype=float)
    growth = np.array(list(df["cell.growth_rate"]), dtype=float)
    assert positions.iloc[0].shape == (3, 5) and (radii > 0).all() and growth.shape == radii.shape
    by_pos = {tuple(np.round(p.T.reshape(-1), 12)) for p in positions}
    assert tuple(np.round(np.asarray(next(iter(last.values()))["pos"]).reshape(-1), 12)) in by_pos
    sub = json.loads(next((acc.path / "subdomains" / "json" / f"{120:020d}").glob("single_*.json")).read_text())
    assert set(sub["element"]) == {"subdomain", "gel_pressure", "surface_friction", "surface_friction_distance"}
    assert set(sub["element"]["subdomain"]) == {"min", "max", "dx", "voxels", "domain_min", "domain_max", "domain_n_voxels"}


def test_run_simulation_over_slabs_equals_single_device(crb_lib):
    """``devices=[...]``: the reference-shaped entry point cut into slabs (three slabs on this one GPU)
    stores exactly the snapshots of the single-device run (slabs are bit-identical)."""
    rng = np.random.default_rng(2)
    n = 900
    pos = rng.uniform(0.5, 107.0, (n, 3))
    agents = [cr.Agent(cr.NewtonDamped3D(pos[k].tolist(), [0.0] * 3, 2.0, 1.0),
                       cr.CellSpecificInteraction(int(k <= n // 2), 6.0, 2.0, 1.5)) for k in range(n)]
    domain = cr.CartesianCuboid.from_boundaries_and_interaction_range([0.0] * 3, [108.0] * 3, 18.0)

    def run(**kw):
        settings = cr.Settings(4, cr.FixedStepsize.from_partial_save_steps(0.0, 0.2, 60, 20),
                               cr.StorageBuilder(priority=["Memory"]))
        return cr.run_simulation(domain, agents, settings, **kw).cells.load_all_elements()

    one, many = run(), run(devices=[0, 0, 0])
    assert sorted(one) == sorted(many) == [0, 20, 40, 60]
    for it in one:
        assert one[it].keys() == many[it].keys()
        for key, c in one[it].items():
            d = many[it][key]
            assert np.array_equal(c["pos"], d["pos"]) and np.array_equal(c["vel"], d["vel"])
            assert c["voxel"] == d["voxel"]
            assert np.array_equal(c["aux"]["positions"], d["aux"]["positions"])


def test_bacterial_rods_over_slabs_through_run_simulation(crb_lib):
    """run_simulation(..., devices=[0, 0, 0]) for rod agents: three slabs of voxel planes driven by one host thread
    (LocalRodSlabGroup) store the snapshots of the single-device run bit for bit — identifiers incl.
    Division(voxel, counter), parents, positions, voxels."""
    from cellular_raza_b200 import workloads

    w = workloads.bacterial_rods(3, 2, rods_per_tile=6, spring_length_range=(5.5, 5.99))
    pos = w.pos.copy()
    pos[:, :, 0] += 180e-6  # rod clusters next to the slab faces x = 200 um and 400 um (24 planes, 3 slabs)
    keep = pos[:, :, 0].max(axis=1) < w.cfg_kwargs["domain_max"][0] - 1e-6
    agents = []
    for k in np.nonzero(keep)[0]:
        mp, ip = w.mech_params[k], w.inter_params[k]
        mech = cr.RodMechanics(pos[k].tolist(), w.vel[k].tolist(), mp[0], mp[1], mp[2], mp[3], mp[4])
        agents.append(cr.RodAgent(mech, cr.RodInteraction(cr.MorsePotential(ip[0], ip[1], ip[2], ip[3])),
                                  growth_rate=mp[5], spring_length_threshold=mp[6], growth_base_rate=mp[7]))
    kw = w.cfg_kwargs
    domain = cr.CartesianCuboidRods(cr.CartesianCuboid.from_boundaries_and_n_voxels(
        kw["domain_min"], kw["domain_max"], [int(x) for x in kw["n_voxels"]]))

    def run(devices):
        settings = cr.Settings(8, cr.FixedStepsize.from_partial_save_steps(0.0, kw["dt"], 90, 30),
                               cr.StorageBuilder(priority=["Memory"]))
        return cr.run_simulation(domain, agents, settings, devices=devices,
                                 aspects=["Mechanics", "Interaction", "Cycle", "NeighborSensing"])

    one, three = run(None), run([0, 0, 0])
    assert one.cells.get_all_iterations() == three.cells.get_all_iterations() == [0, 30, 60, 90]
    for it in (30, 60, 90):
        a, b = one.cells.memory[it], three.cells.memory[it]
        assert list(a) == list(b)  # same identifiers in the same save order
        for key in a:
            assert np.array_equal(a[key]["pos"], b[key]["pos"]) and a[key]["voxel"] == b[key]["voxel"]
            assert a[key]["parent"] == b[key]["parent"]
    assert len(one.cells.memory[90]) > len(agents)  # rods divided
