"""Shared driver of the reference-derived fixtures ``tests/golden/ref_derived_*.json``.

The fixtures are written by ``tools/make_ref_derived.py`` from ``tests/ref_transcript.py`` — an independent,
literal transcription of the reference's Rust that never reads ``oracle/``.  ``check(sim, fix)`` replays a fixture
on any object with the ``DeviceSimulation`` method surface (the CUDA path through the C ABI, or the C++ oracle)
and compares forces, trajectories, voxel keys, save order, ring-buffer fill and migration sets.

Tolerances (BASELINE.json north_star): forces 1e-12 with the metric of SURVEY.md 8d
(|df|_inf / max(|f|_inf, sum |f_ij|_inf)); positions / velocities 1e-11 of the domain extent over the <= 12 steps of
a fixture (1e-9 over 100 steps is the contract); voxel index, save order, migration sets: exact.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from cellular_raza_b200 import _abi
from cellular_raza_b200.device import make_config

GOLDEN = Path(__file__).resolve().parent / "golden"
MECH = {"newton": _abi.CRB_MECH_NEWTON_DAMPED, "brownian": _abi.CRB_MECH_BROWNIAN, "langevin": _abi.CRB_MECH_LANGEVIN}
INTER = {"cell_sorting": _abi.CRB_INT_CELL_SORTING, "cell_sorting_bench": _abi.CRB_INT_CELL_SORTING_BENCH,
         "morse": _abi.CRB_INT_MORSE, "bound_lj": _abi.CRB_INT_BOUND_LJ, "mie": _abi.CRB_INT_MIE}

FORCE_TOL = 1e-12
STATE_TOL = 1e-11


def names():
    return sorted(p.stem[len("ref_derived_"):] for p in GOLDEN.glob("ref_derived_*.json"))


def load(name: str) -> dict:
    return json.loads((GOLDEN / f"ref_derived_{name}.json").read_text())


def config(fix: dict, **overrides):
    c = fix["config"]
    kw = dict(
        dim=c["dim"], mechanics_kind=MECH[c["mechanics"]["kind"]], interaction_kind=INTER[c["interaction"]["kind"]],
        domain_min=c["domain_min"], domain_max=c["domain_max"], n_voxels=c["n_voxels"], dx=c["dx"], dt=c["dt"],
        capacity=len(fix["agents"]["pos"]), solver_order=c["solver_order"], pair_weight_mode=_abi.CRB_WEIGHT_CHILI,
        ref_n_subdomains=c["n_threads"],
        noise_mode=_abi.CRB_NOISE_TABLE if fix["noise"] is not None else _abi.CRB_NOISE_NONE)
    kw.update(overrides)
    return make_config(**kw)


def upload(sim, fix: dict):
    a = fix["agents"]
    sim.set_agents(np.array(a["pos"]), np.array(a["vel"]), np.array(a["mech_params"]), np.array(a["inter_params"]),
                   np.array(a["species"], dtype=np.uint8))
    if fix["noise"] is not None:
        sim.set_noise(np.array(fix["noise"]))


def _forces(sim):
    f = sim.debug_forces() if hasattr(sim, "debug_forces") else sim.forces()[0]
    ids = sim.get_agents()[2]
    return f[np.argsort(ids, kind="stable")]


def check(sim, fix: dict, order_exact: bool = True):
    """replays the fixture; returns the worst force / state errors seen (for reporting)"""
    c = fix["config"]
    extent = max(b - a for a, b in zip(c["domain_min"], c["domain_max"]))
    upload(sim, fix)
    f_ref = np.array(fix["forces_t0"])
    ab = np.array(fix["forces_t0_abs_sum"])
    f = _forces(sim)
    den = np.maximum(np.abs(f_ref).max(axis=1), ab)
    den = np.where(den == 0, 1.0, den)
    ferr = float((np.abs(f - f_ref).max(axis=1) / den).max())
    assert ferr <= FORCE_TOL, f"{fix['name']}: force error {ferr:.3e}"
    done, serr = 0, 0.0
    for cp in fix["checkpoints"]:
        sim.step(cp["step"] - done)
        done = cp["step"]
        pos, vel, ids, vox = sim.get_agents()
        o = np.argsort(ids, kind="stable")
        np.testing.assert_array_equal(vox[o], np.array(cp["voxel"], dtype=np.uint64),
                                      err_msg=f"{fix['name']} step {done}: voxel index")
        if order_exact:
            np.testing.assert_array_equal(ids, np.array(cp["save_order"], dtype=np.uint64),
                                          err_msg=f"{fix['name']} step {done}: in-voxel (save) order")
        pe = float(np.abs(pos[o] - np.array(cp["pos"])).max() / extent)
        ve = float(np.abs(vel[o] - np.array(cp["vel"])).max())
        vscale = max(float(np.abs(np.array(cp["vel"])).max()), 1.0)
        assert pe <= STATE_TOL, f"{fix['name']} step {done}: position error {pe:.3e}"
        assert ve <= STATE_TOL * vscale * 100, f"{fix['name']} step {done}: velocity error {ve:.3e}"
        serr = max(serr, pe)
        hl = sim.get_aux()[2]
        np.testing.assert_array_equal(hl[o], np.array(cp["hist_len"], dtype=np.uint32),
                                      err_msg=f"{fix['name']} step {done}: ring-buffer fill")
        mig = sim.get_migrations()
        got = sorted(map(list, zip(*(np.asarray(m).tolist() for m in mig))))
        assert got == cp["migrations"], f"{fix['name']} step {done}: migration set"
    return ferr, serr
