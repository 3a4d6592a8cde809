"""Host-side mirror of the reference interface (cellular_raza_b200/backend.py): pure host logic on CPU."""
import numpy as np
import pytest

from cellular_raza_b200 import backend as cr


def test_domain_constructors_match_reference_tests():
    # cartesian_cuboid_n.rs:167-182
    d = cr.CartesianCuboid.from_boundaries_and_interaction_range([2.0, 3.0, 1.0], [10.0, 10.0, 20.0], 2.0)
    assert d.n_voxels.tolist() == [4, 3, 9]
    np.testing.assert_allclose(d.dx, [2.0, 7.0 / 3.0, 19.0 / 9.0])
    d = cr.CartesianCuboid.from_boundaries_and_n_voxels([0.0, 0.0], [10.0, 5.0], [5, 5])
    np.testing.assert_array_equal(d.dx, [2.0, 1.0])
    with pytest.raises(cr.SimulationError):
        cr.CartesianCuboid.from_boundaries_and_n_voxels([0.0, 6.0], [10.0, 5.0], [5, 5])


def test_fixed_stepsize_event_schedule():
    # time.rs:94-111: n_steps = 10, save_interval = 11 -> save points 0 and 11 => 11 iterations
    # (the benchmark's `from_partial_save_steps(0, dt, n, n + 1)` quirk)
    t = cr.FixedStepsize.from_partial_save_steps(0.0, 0.2, 10, 11)
    assert t.save_iterations == [0, 11] and t.maximum_iterations == 11
    t = cr.FixedStepsize.from_partial_save_steps(0.0, 0.2, 10_000, 20)
    assert t.save_iterations[:3] == [0, 20, 40] and t.maximum_iterations == 10_000
    t = cr.FixedStepsize.from_partial_save_interval(0.0, 0.1, 20.0, 2.0)
    assert t.save_iterations[1] == 20 and t.maximum_iterations == 200
    t = cr.FixedStepsize.from_partial_save_points(1.0, 0.2, [3.0, 5.0, 11.0, 20.0])
    assert t.save_iterations == [10, 20, 50, 95]
    # F::round is half away from zero (Python's round is half to even): (2.25 - 1.0) / 0.5 = 2.5 -> iteration 3
    assert cr.FixedStepsize.from_partial_save_points(1.0, 0.5, [2.25]).save_iterations == [3]
    # from_partial_save_freq as written (time.rs:135-168): events at (0..max).step_by(freq), end point exclusive
    t = cr.FixedStepsize.from_partial_save_freq(0.0, 0.1, 10.0, 25)
    assert t.maximum_iterations == 100 and t.save_iterations == [0, 25, 50, 75]
    t = cr.FixedStepsize.from_partial_save_freq(0.0, 0.1, 10.1, 25)
    assert t.maximum_iterations == 101 and t.save_iterations == [0, 25, 50, 75, 100]
    with pytest.raises(ValueError):
        cr.FixedStepsize.from_partial_save_points(1.0, 0.2, [0.5])
    with pytest.raises(ValueError):
        cr.FixedStepsize.from_partial_save_points(1.0, 0.2, [])


def test_agent_packing():
    agents = [cr.Agent(cr.NewtonDamped3D([1.0, 2.0, 3.0], [0.1, 0.0, 0.0], 2.0, 1.0),
                       cr.CellSpecificInteraction(k % 2, 6.0, 2.0, 1.5)) for k in range(4)]
    mk, ik, pos, vel, mp, ip, sp = cr._pack(agents, 3)
    assert (mk, ik) == (0, 0) and pos.shape == (4, 3) and sp.tolist() == [0, 1, 0, 1]
    assert mp[0].tolist() == [2.0, 1.0] and ip[0].tolist() == [6.0, 2.0, 1.5]
    b = cr.Agent(cr.Brownian3D([1.0, 1.0, 1.0], 0.3, 0.3), cr.MorsePotential(3.0, 0.5, 12.0, 0.01))
    mk, ik, pos, vel, mp, ip, sp = cr._pack([b], 3)
    assert (mk, ik) == (1, 2) and vel.tolist() == [[0.0, 0.0, 0.0]] and ip[0].tolist() == [3.0, 0.5, 12.0, 0.01]
    with pytest.raises(cr.SimulationError):
        cr._pack(agents + [b], 3)


def test_unsupported_aspects_are_rejected():
    d = cr.CartesianCuboid.from_boundaries_and_n_voxels([0.0] * 3, [10.0] * 3, [2, 2, 2])
    s = cr.Settings(1, cr.FixedStepsize.from_partial_save_steps(0.0, 0.1, 2, 1))
    a = [cr.Agent(cr.NewtonDamped3D([1.0, 1.0, 1.0]), cr.MorsePotential(1.0, 1.0, 2.0, 0.1))]
    with pytest.raises(cr.SimulationError):
        cr.run_simulation(d, a, s, aspects=["Mechanics", "Reactions"])


def test_rod_agent_packing_and_aspects():
    # Agent of cellular_raza-examples/bacterial_rods/src/main.rs:14-25, 120-138
    um = cr.MICRO_METRE
    mech = cr.RodMechanics(np.zeros((5, 3)).tolist(), np.zeros((5, 3)).tolist(), 0.0, 10.0, 1.0 * um, 2.0 * um, 1.5)
    a = cr.RodAgent(mech, cr.RodInteraction(cr.MorsePotential(3.0 * um, 0.5 / um, 5.0 * um, 0.1 * um ** 2)))
    assert a.mech_row() == [0.0, 10.0, 1.0 * um, 2.0 * um, 1.5, cr.GROWTH_BASE_RATE, 6.0 * um, cr.GROWTH_BASE_RATE]
    assert a.interaction.row() == [3.0 * um, 0.5 / um, 5.0 * um, 0.1 * um ** 2]
    d = cr.CartesianCuboidRods(cr.CartesianCuboid.from_boundaries_and_n_voxels([0.0] * 3, [200 * um, 15 * um, 45 * um], [8, 1, 2]))
    s = cr.Settings(8, cr.FixedStepsize.from_partial_save_steps(0.0, 0.1, 2, 1))
    with pytest.raises(cr.SimulationError):  # the rod agent needs all four aspects of main.rs:206
        cr.run_simulation(d, [a], s, aspects=["Mechanics", "Interaction"])
    assert cr._ident((1, 7, 0)) == ("Initial", 7) and cr._ident_json(("Division", 3, 2)) == {"Division": [3, 2]}
    m = cr._matrix_json(np.arange(6.0).reshape(2, 3))
    assert m == [[0.0, 3.0, 1.0, 4.0, 2.0, 5.0], 2, None]  # column-major data, nrows, Const<3> as unit


def test_snapshot_body_has_the_reference_serde_layout():
    """cell body = the building-block structs' own serde field names (mechanics.rs:54-65, 253-261, 413-426;
    interaction.rs:87-96, 215-224, 322-335; cell_sorting/src/main.rs:25-38, 99-105)"""
    names = {"mech_cls": cr.NewtonDamped3D, "inter_cls": cr.CellSpecificInteraction}
    body = cr._cell_body(names, [1.0, 2.0, 3.0], [0.1, 0.2, 0.3], [2.0, 1.5], [6.0, 2.0, 1.5], 1)
    assert body == {"interaction": {"species": "BlueCell", "cell_radius": 6.0, "potential_strength": 2.0,
                                    "relative_interaction_range": 1.5},
                    "mechanics": {"pos": [1.0, 2.0, 3.0], "vel": [0.1, 0.2, 0.3], "damping_constant": 2.0, "mass": 1.5}}
    names = {"mech_cls": cr.Brownian2D, "inter_cls": cr.MorsePotential}
    body = cr._cell_body(names, [1.0, 2.0], [0.0, 0.0], [0.3, 0.4], [3.0, 0.5, 12.0, 0.01], 0)
    assert body["mechanics"] == {"pos": [1.0, 2.0], "diffusion_constant": 0.3, "kb_temperature": 0.4}
    assert body["interaction"] == {"radius": 3.0, "potential_stiffness": 0.5, "cutoff": 12.0, "strength": 0.01}
    names = {"mech_cls": cr.Langevin3D, "inter_cls": None}
    body = cr._cell_body(names, [1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [1.5, 0.7, 0.4], [], 0)
    assert body == {"mechanics": {"pos": [1.0, 2.0, 3.0], "vel": [4.0, 5.0, 6.0], "mass": 1.5, "damping": 0.7,
                                  "kb_temperature": 0.4}}


def test_subdomain_snapshots_follow_create_subdomains():
    """CartesianSubDomain of every chili subdomain (cartesian_cuboid_n.rs:415-423, 500-583): 5^3 voxels into 4
    subdomains hold 32 + 31 + 31 + 31 voxels (the reference's own test :380-402), contiguous in the enumeration"""
    subs = cr.chili_subdomains([0.0] * 3, [100.0] * 3, [5, 5, 5], [20.0] * 3, 4)
    assert [len(s["voxels"]) for s in subs] == [32, 31, 31, 31]
    flat = [tuple(v) for s in subs for v in s["voxels"]]
    assert flat[0] == (0, 0, 0) and flat[1] == (0, 0, 1) and flat[-1] == (4, 4, 4) and len(set(flat)) == 125
    assert subs[0]["min"] == [0.0, 0.0, 0.0] and subs[0]["max"] == [40.0, 100.0, 100.0]
    assert subs[0]["domain_min"] == [0.0] * 3 and subs[0]["domain_n_voxels"] == [5, 5, 5]
    assert set(subs[0]) == {"min", "max", "dx", "voxels", "domain_min", "domain_max", "domain_n_voxels"}


@pytest.mark.parametrize("name,twin", [("NewtonDamped3DF32", "NewtonDamped3D"), ("Brownian2DF32", "Brownian2D"),
                                       ("Langevin1DF32", "Langevin1D"), ("MorsePotentialF32", "MorsePotential"),
                                       ("BoundLennardJonesF32", "BoundLennardJones"), ("MiePotentialF32", "MiePotential")])
def test_f32_building_blocks_are_refused_with_a_message(name, twin):
    """mechanics.rs:202-204, 385-387, 569-571; interaction.rs:144-145, 262-263, 405: the refusal is explicit
    (CRB_ERR_INVALID_ARG + message naming the f64 twin), never a silent widening"""
    from cellular_raza_b200 import _abi

    with pytest.raises(cr.SimulationError) as e:
        getattr(cr, name)([0.0, 0.0, 0.0], [0.0, 0.0, 0.0], 1.0, 1.0)
    assert e.value.code == _abi.CRB_ERR_INVALID_ARG
    assert twin in str(e.value) and "f64" in str(e.value)


def test_storage_read_side_like_storage_interface_load(tmp_path):
    """StorageInterfaceLoad (cellular_raza-core/src/storage/concepts.rs:863-994): single element, element history
    (one uninterrupted span, as the reference's doc example :880-898), all histories — from memory and from the
    SerdeJson layout (iteration directories of batch files), with Initial and Division identifiers"""
    import json

    states = {0: ["E1", "E2", "E3"], 1: ["E1", "E2", "E3", "E4"], 2: ["E1", "E2", "E3", "E4"], 3: ["E1", "E2", "E4"],
              4: ["E2", "E4"], 5: ["E2", "E4", "E5"], 6: ["E4", "E5"]}
    ident = {"E1": ("Initial", 1), "E2": ("Initial", 2), "E3": ("Initial", 3), "E4": ("Division", 7, 1), "E5": ("Division", 7, 2)}
    mem = cr.StorageManager(None)
    disk = cr.StorageManager(tmp_path / "cells" / "json")
    for it, names in states.items():
        mem.memory[it] = {ident[n]: {"pos": [float(it)], "name": n} for n in names}
        d = disk.json_dir / f"{it:020d}"
        d.mkdir(parents=True)
        data = [{"identifier": cr._ident_json(ident[n]),
                 "element": [{"identifier": cr._ident_json(ident[n]), "parent": None, "cell": {"name": n, "it": it}}, {"mechanics": {}}]}
                for n in names]
        (d / f"batch_{0:020d}_{0:020d}.json").write_text(json.dumps({"data": data[:2]}))
        (d / f"batch_{1:020d}_{0:020d}.json").write_text(json.dumps({"data": data[2:]}))
    for sm in (mem, disk):
        assert sm.get_all_iterations() == list(range(7))
        assert sorted(sm.load_all_elements_at_iteration(3)) == sorted(ident[n] for n in states[3])
        assert sm.load_single_element(0, ("Division", 7, 1)) is None
        assert sm.load_single_element(1, {"Division": [7, 1]}) is not None
        assert sorted(sm.load_element_history(("Initial", 1))) == [0, 1, 2, 3]
        assert sorted(sm.load_element_history({"Division": [7, 2]})) == [5, 6]
        assert sorted(sm.load_element_history(("Initial", 3))) == [0, 1, 2]
        hist = sm.load_all_element_histories()
        assert {k: sorted(v) for k, v in hist.items()} == {
            ident[n]: [it for it, names in states.items() if n in names] for n in ident}
        assert list(sm.load_all_elements()) == list(range(7))
    assert disk.load_single_element(5, ("Division", 7, 2))["cell"] == {"name": "E5", "it": 5}
    # an identifier that re-appears after a gap: the reference stops at the gap (:880-898 "invalid state")
    mem.memory[8] = {("Initial", 1): {"pos": [8.0]}}
    assert sorted(mem.load_element_history(("Initial", 1))) == [0, 1, 2, 3]
    assert cr.StorageManager(None).load_all_elements_at_iteration(0) == {}
