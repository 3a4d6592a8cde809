#!/usr/bin/env python
"""Writes the regression fixtures under tests/golden/.

The reference (a Rust workspace) cannot be built or run in this image, so these vectors are NOT outputs of
the reference: they are outputs of the CPU restatement of chili (oracle/), which itself is pinned on every
known-answer test the reference holds for this path (tests/test_oracle_known_answers.py,
tests/test_oracle_rods.py).  They freeze the oracle: a later change of the oracle or of its compiler flags
shows up as a diff against these files, and the device is compared against them without the oracle having to
be built on the GPU box.      python tools/make_golden.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from cellular_raza_b200 import _abi, workloads  # noqa: E402
from oracle.chili_oracle import ChiliOracle  # noqa: E402
from oracle.rods_oracle import RodsOracle  # noqa: E402

OUT = ROOT / "tests" / "golden"


def cell_sorting_c1():
    """configs[0]: cell_sorting as shipped (1,600 cells, 6^3 voxels, n_threads 4): forces at t = 0 and the
    state after 20 and 100 steps, by identifier."""
    w = workloads.cell_sorting(n_threads=4)
    o = ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    f, ab = o.forces()
    ids = o.get_agents()[2]
    a = np.argsort(ids)
    out = {"pos0": w.pos, "species": w.species, "force0": f[a], "abs_sum0": ab[a]}
    done = 0
    for steps in (20, 100):
        o.step(steps - done)
        done = steps
        p, v, ids, vox = o.get_agents()
        a = np.argsort(ids)
        out[f"pos{steps}"], out[f"vel{steps}"], out[f"voxel{steps}"] = p[a], v[a], vox[a]
    o.close()
    np.savez_compressed(OUT / "cell_sorting_c1_oracle.npz", **out)


def brownian_morse_small():
    """C3-type agents (Brownian3D + Morse) with the shared noise table, 300 agents, 40 steps."""
    w = workloads.brownian_morse(300, 12.0 * 5, n_threads=3)
    z = workloads.noise_table(40, w.n, 3, 7)
    o = ChiliOracle(w.config())
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    o.set_noise(z)
    o.step(40)
    p, v, ids, vox = o.get_agents()
    a = np.argsort(ids)
    o.close()
    np.savez_compressed(OUT / "brownian_morse_oracle.npz", pos0=w.pos, noise=z, pos40=p[a], voxel40=vox[a])


def rods_small():
    """bacterial rods: 3 x 3 tiles of 12 rods close to the division threshold, 60 steps (rods divide)."""
    w = workloads.bacterial_rods(3, 3, rods_per_tile=12, spring_length_range=(5.6, 5.99))
    o = RodsOracle(w.config(), n_threads=4)
    o.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    f, ab, nb = o.forces()
    o.step(60)
    p, v, ident, parent, vox = o.get_agents()
    o.close()
    np.savez_compressed(OUT / "bacterial_rods_oracle.npz", pos0=w.pos, mech=w.mech_params, inter=w.inter_params,
                        force0=f, abs_sum0=ab, neighbors0=nb, pos60=p, vel60=v, ident60=ident, parent60=parent,
                        voxel60=vox)


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    cell_sorting_c1()
    brownian_morse_small()
    rods_small()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")
