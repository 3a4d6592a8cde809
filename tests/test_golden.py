"""Regression fixtures under tests/golden/ (written by tools/make_golden.py FROM THE ORACLE — the Rust reference
cannot run in this image, see the script's header): the oracle must still reproduce them (CPU), and the device
must match them within the north_star tolerances without the oracle in the loop (GPU)."""
from pathlib import Path

import numpy as np
import pytest

from cellular_raza_b200 import workloads

G = Path(__file__).resolve().parent / "golden"
FORCE_TOL, TRAJ_TOL = 1e-12, 1e-9


def _c1():
    return workloads.cell_sorting(n_threads=4), np.load(G / "cell_sorting_c1_oracle.npz")


def _by_id(sim):
    p, v, ids, vox = sim.get_agents()
    a = np.argsort(ids)
    return p[a], v[a], vox[a], a


def _check_c1(sim, forces, g, tol_f, tol_x):
    w, _ = _c1()
    assert np.array_equal(w.pos, g["pos0"]) and np.array_equal(w.species, g["species"])  # same workload
    den = np.maximum(np.abs(g["force0"]).max(axis=1), g["abs_sum0"])
    assert (np.abs(forces - g["force0"]).max(axis=1) / np.where(den == 0, 1, den)).max() <= tol_f
    done = 0
    for steps in (20, 100):
        sim.step(steps - done)
        done = steps
        p, v, vox, _ = _by_id(sim)
        assert np.abs(p - g[f"pos{steps}"]).max() / 110.0 <= tol_x
        assert np.abs(v - g[f"vel{steps}"]).max() <= tol_x * 110.0
        assert (vox != g[f"voxel{steps}"]).sum() <= 1  # an agent within 1e-9 L of a voxel face may differ


def test_oracle_reproduces_cell_sorting_fixture(oracle_lib):
    w, g = _c1()
    o = oracle_lib.ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    f, _ = o.forces()
    _check_c1(o, f[np.argsort(o.get_agents()[2])], g, 1e-14, 1e-13)
    o.close()


def test_oracle_reproduces_brownian_and_rod_fixtures(oracle_lib):
    from oracle.rods_oracle import RodsOracle

    g = np.load(G / "brownian_morse_oracle.npz")
    w = workloads.brownian_morse(300, 12.0 * 5, n_threads=3)
    assert np.array_equal(w.pos, g["pos0"])
    o = oracle_lib.ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    o.set_noise(g["noise"])
    o.step(40)
    p, _, vox, _ = _by_id(o)
    assert np.abs(p - g["pos40"]).max() / 60.0 <= 1e-13 and np.array_equal(vox, g["voxel40"])
    o.close()
    g = np.load(G / "bacterial_rods_oracle.npz")
    w = workloads.bacterial_rods(3, 3, rods_per_tile=12, spring_length_range=(5.6, 5.99))
    assert np.array_equal(w.pos, g["pos0"])
    r = RodsOracle(w.config(), n_threads=2)
    r.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    r.step(60)
    p, v, ident, parent, vox = r.get_agents()
    assert np.array_equal(ident, g["ident60"]) and np.array_equal(parent, g["parent60"])
    assert np.array_equal(vox, g["voxel60"]) and len(ident) > w.n  # rods divided
    assert np.abs(p - g["pos60"]).max() / max(w.cfg_kwargs["domain_max"]) <= 1e-13
    r.close()


@pytest.mark.gpu
def test_device_matches_cell_sorting_fixture(crb_lib):
    from cellular_raza_b200.device import DeviceSimulation

    w, g = _c1()
    d = DeviceSimulation(w.config())
    d.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    f = d.debug_forces()[np.argsort(d.get_agents()[2])]
    _check_c1(d, f, g, FORCE_TOL, TRAJ_TOL)
    d.close()


@pytest.mark.gpu
def test_device_matches_brownian_and_rod_fixtures(crb_lib):
    from cellular_raza_b200.device import DeviceSimulation
    from cellular_raza_b200.rods import RodSimulation

    g = np.load(G / "brownian_morse_oracle.npz")
    w = workloads.brownian_morse(300, 12.0 * 5, n_threads=3)
    d = DeviceSimulation(w.config())
    d.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    d.set_noise(g["noise"])
    d.step(40)
    p, _, vox, _ = _by_id(d)
    assert np.abs(p - g["pos40"]).max() / 60.0 <= TRAJ_TOL and (vox != g["voxel40"]).sum() <= 1
    d.close()
    g = np.load(G / "bacterial_rods_oracle.npz")
    w = workloads.bacterial_rods(3, 3, rods_per_tile=12, spring_length_range=(5.6, 5.99))
    r = RodSimulation(w.config())
    r.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    f, nb = r.debug_forces(), r.debug_neighbor_counts()
    den = np.maximum(np.abs(g["force0"]).reshape(len(f), -1).max(axis=1), g["abs_sum0"])
    assert (np.abs(f - g["force0"]).reshape(len(f), -1).max(axis=1) / np.where(den == 0, 1, den)).max() <= FORCE_TOL
    assert np.array_equal(nb, g["neighbors0"])
    r.step(60)
    p, v, ident, parent, vox = r.get_agents()
    assert np.array_equal(ident, g["ident60"]) and np.array_equal(parent, g["parent60"])  # same divisions, same order
    assert np.array_equal(vox, g["voxel60"])
    assert np.abs(p - g["pos60"]).max() / max(w.cfg_kwargs["domain_max"]) <= TRAJ_TOL
    r.close()
