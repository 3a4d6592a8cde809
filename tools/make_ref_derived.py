#!/usr/bin/env python
"""Writes tests/golden/ref_derived_*.json from tests/ref_transcript.py — the independent, literal Python
transcription of the Rust sources of the chili mechanics/interaction step (it never reads oracle/).

    python tools/make_ref_derived.py

Every fixture holds the inputs (agent-major), the reference configuration (n_threads = emulated chili
subdomains, solver order, dt) and what the transcript computes: forces before the first step, and after the
listed steps the positions, velocities, plain voxel index of every agent (by input index), the save order
(agents in ascending voxel / in-voxel order) and the migration set of that step.
Cases cover what no test of the reference pins (SURVEY.md 8c "not pinned"): pair forces of every potential
incl. the clamp region r < 0.316 R, coincident positions, the cutoff edge +- 1 ulp, the species split and
the benchmark's sign convention; the 1 / 0.5 / 1 pair weights of chili over 1, 2, 3 and 4 subdomains; the
push-before-read Adams-Bashforth recurrence (steps 0, 1, 2, ...) for solver orders 0, 1, 2; reflective
boundary; migration incl. the local-then-remote append order.
"""
from __future__ import annotations

import json
import math
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
import ref_transcript as R  # noqa: E402

OUT = ROOT / "tests" / "golden"


def nextafter(x, d):
    return math.nextafter(x, d)


def crafted_pairs_3d(cutoff, r_clamp):
    """positions (inside [0, 54]^3, voxel edge 18) that hit the special branches of the potentials"""
    p = []
    # clamp region r < 0.316 R, both agents in voxel (0, 0, 0)
    p += [[5.0, 5.0, 5.0], [5.0 + 0.5 * r_clamp, 5.0, 5.0]]
    # coincident positions (zero force in cell_sorting / Morse), voxel (1, 1, 1)
    p += [[25.0, 25.0, 25.0], [25.0, 25.0, 25.0]]
    # cutoff edge across a voxel face: r == cutoff exactly, one ulp inside, one ulp outside
    p += [[10.0, 40.0, 10.0], [10.0 + cutoff, 40.0, 10.0]]
    p += [[10.0, 44.0, 30.0], [nextafter(10.0 + cutoff, 0.0), 44.0, 30.0]]
    p += [[10.0, 48.0, 50.0], [nextafter(10.0 + cutoff, 100.0), 48.0, 50.0]]
    # agent close to the walls with velocity outwards (reflection) is added by the caller through vel
    return p


def build(name, *, dim, mn, mx, interaction_range=None, n_voxels=None, n_threads, dt, order, mech, inter, n_agents,
          seed, steps, vel_scale=0.0, crafted=None, noise=False, hetero=False, crafted_vel=None, write=True):
    rng = random.Random(seed)
    if interaction_range is not None:
        dom = R.CartesianCuboid.from_boundaries_and_interaction_range(mn, mx, interaction_range)
    else:
        dom = R.CartesianCuboid.from_boundaries_and_n_voxels(mn, mx, n_voxels)
    pos = [list(map(float, q[:dim])) for q in (crafted or [])]
    while len(pos) < n_agents:
        pos.append([rng.uniform(mn[i] + 1e-3, mx[i] - 1e-3) for i in range(dim)])
    n = len(pos)
    vel = [[rng.uniform(-vel_scale, vel_scale) if vel_scale else 0.0 for _ in range(dim)] for _ in range(n)]
    for i, v in enumerate(crafted_vel or []):
        vel[i] = list(map(float, v[:dim]))
    species = [1 if i <= n // 2 else 0 for i in range(n)]  # main.rs:132: n <= N_CELLS_1 are BlueCell
    mech_rows, inter_rows, cells = [], [], []
    for i in range(n):
        scale = (0.8 + 0.4 * rng.random()) if hetero else 1.0
        if mech["kind"] == "newton":
            row = [mech["damping"], mech["mass"] * scale]
            m = R.NewtonDamped(list(pos[i]), list(vel[i]), row[0], row[1])
        elif mech["kind"] == "brownian":
            row = [mech["diffusion"], mech["kt"]]
            m = R.Brownian(list(pos[i]), row[0], row[1])
            vel[i] = [0.0] * dim
        else:
            row = [mech["mass"] * scale, mech["damping"], mech["kt"]]
            m = R.Langevin(list(pos[i]), list(vel[i]), row[0], row[1], row[2])
        if inter["kind"] in ("cell_sorting", "cell_sorting_bench"):
            irow = [inter["radius"] * scale, inter["strength"], inter["range"]]
            it = R.CellSpecificInteraction(species[i], irow[0], irow[1], irow[2], bench=inter["kind"].endswith("bench"))
        elif inter["kind"] == "morse":
            irow = [inter["radius"] * scale, inter["stiffness"], inter["cutoff"], inter["strength"]]
            it = R.MorsePotential(*irow)
        elif inter["kind"] == "bound_lj":
            irow = [inter["epsilon"], inter["sigma"], inter["bound"], inter["cutoff"]]
            it = R.BoundLennardJones(*irow)
        else:
            irow = [inter["radius"] * scale, inter["strength"], inter["bound"], inter["cutoff"], inter["en"], inter["em"]]
            it = R.MiePotential(*irow)
        mech_rows.append(row)
        inter_rows.append(irow)
        cells.append(R.Cell(i, m, it))
    t = R.ChiliTranscript(dom, n_threads, dt, order)
    t.set_cells(cells)
    n_steps = max(steps)
    z = None
    if noise:
        z = [[[rng.gauss(0.0, 1.0) for _ in range(dim)] for _ in range(n)] for _ in range(n_steps)]
    f0 = t.forces()
    out = {
        "name": name,
        "source": "tests/ref_transcript.py: literal transcription of the reference's Rust (not oracle output); "
                  "regenerate with tools/make_ref_derived.py",
        "config": {"dim": dim, "domain_min": list(map(float, mn)), "domain_max": list(map(float, mx)),
                   "n_voxels": list(dom.n_voxels), "dx": list(dom.dx), "dt": dt, "solver_order": order,
                   "n_threads": n_threads, "mechanics": mech, "interaction": inter},
        "agents": {"pos": pos, "vel": vel, "species": species, "mech_params": mech_rows, "inter_params": inter_rows},
        "noise": z,
        "forces_t0": [f0[i][0] for i in range(n)],
        "forces_t0_abs_sum": [f0[i][1] for i in range(n)],
        "checkpoints": [],
    }
    for k in range(n_steps):
        row = (lambda key, zk=z[k]: zk[key]) if z is not None else None
        t.step(row)
        if k + 1 in steps:
            order_keys, vox = [], [0] * n
            P, V = [None] * n, [None] * n
            hist_len = [0] * n
            for p, cell, aux in t.cells_in_order():
                order_keys.append(cell.key)
                vox[cell.key] = p
                P[cell.key] = list(cell.mech.pos)
                V[cell.key] = cell.mech.velocity()
                hist_len[cell.key] = aux.positions.get_size()
            out["checkpoints"].append({"step": k + 1, "pos": P, "vel": V, "voxel": vox, "save_order": order_keys,
                                       "hist_len": hist_len,
                                       "migrations": sorted(map(list, t.last_migrations))})
    if not write:  # (tests/test_ref_transcript.py builds random cases in memory)
        return out
    OUT.mkdir(exist_ok=True)
    path = OUT / f"ref_derived_{name}.json"
    path.write_text(json.dumps(out))
    nm = sum(len(c["migrations"]) for c in out["checkpoints"])
    print(f"{path.name}: {n} agents, {len(out['checkpoints'])} checkpoints, {nm} migrations at checkpoints, "
          f"{path.stat().st_size // 1024} KiB")


CS = {"kind": "cell_sorting", "radius": 6.0, "strength": 2.0, "range": 1.5}
NEWTON = {"kind": "newton", "damping": 2.0, "mass": 1.0}
MORSE = {"kind": "morse", "radius": 3.0, "stiffness": 0.5, "cutoff": 12.0, "strength": 0.01}


def main():
    crafted = crafted_pairs_3d(18.0, 6.0 * 0.3162277660168379)
    box = dict(dim=3, mn=[0.0, 0.0, 0.0], mx=[54.0, 54.0, 54.0], interaction_range=18.0)
    for nt in (1, 2, 3, 4):
        build(f"cell_sorting_3d_t{nt}", **box, n_threads=nt, dt=0.2, order=2, mech=NEWTON, inter=CS, n_agents=150,
              seed=10 + nt, steps=[1, 2, 3, 6, 12], vel_scale=12.0, crafted=crafted)
    build("cell_sorting_bench_sign_t2", **box, n_threads=2, dt=0.2, order=2, mech=NEWTON,
          inter=dict(CS, kind="cell_sorting_bench"), n_agents=120, seed=21, steps=[1, 2, 5], vel_scale=8.0, crafted=crafted)
    for order in (0, 1):
        build(f"cell_sorting_3d_order{order}_t3", **box, n_threads=3, dt=0.2, order=order, mech=NEWTON, inter=CS,
              n_agents=100, seed=30 + order, steps=[1, 2, 3, 4], vel_scale=10.0)
    build("cell_sorting_hetero_t2", **box, n_threads=2, dt=0.2, order=2, mech=NEWTON, inter=CS, n_agents=120, seed=41,
          steps=[1, 3, 6], vel_scale=6.0, hetero=True)
    # long fastest axis, 5 subdomains that cut through rows
    build("cell_sorting_slab_t5", dim=3, mn=[0.0, 0.0, 0.0], mx=[36.0, 36.0, 126.0], interaction_range=18.0, n_threads=5,
          dt=0.2, order=2, mech=NEWTON, inter=CS, n_agents=180, seed=43, steps=[1, 2, 8], vel_scale=15.0)
    # in-voxel order of movers (update_mechanics.rs:457-470, 492-511): voxel 2 = first voxel of subdomain 1 receives
    # in ONE step a mover from voxel 1 (subdomain 0: SendCell, appended in step 2) and a mover from voxel 3 (its own
    # subdomain: appended in step 1) => save order [stayer, local mover, remote mover] although the remote mover's
    # source voxel is lower; voxel 1 = last voxel of subdomain 0 gets the mirrored pair
    build("mover_order_t2", dim=3, mn=[0.0, 0.0, 0.0], mx=[18.0, 18.0, 72.0], interaction_range=18.0, n_threads=2, dt=0.2,
          order=2, mech={"kind": "newton", "damping": 0.0, "mass": 1.0}, inter=dict(CS, strength=1e-6), n_agents=40, seed=44,
          steps=[1, 2], vel_scale=1.0,
          crafted=[[9.0, 9.0, 45.0], [9.0, 5.0, 35.0], [9.0, 13.0, 55.0], [5.0, 9.0, 27.0], [4.0, 4.0, 37.0], [14.0, 14.0, 17.0]],
          crafted_vel=[[0.0, 0.0, 0.0], [0.0, 0.0, 10.0], [0.0, 0.0, -10.0], [0.0, 0.0, 0.0], [0.0, 0.0, -10.0], [0.0, 0.0, 10.0]])
    build("mover_order_t4", dim=3, mn=[0.0, 0.0, 0.0], mx=[36.0, 36.0, 72.0], interaction_range=18.0, n_threads=4, dt=0.2,
          order=2, mech={"kind": "newton", "damping": 0.0, "mass": 1.0}, inter=dict(CS, strength=1e-6), n_agents=160, seed=45,
          steps=[1, 2, 3], vel_scale=40.0)
    morse_crafted = [[5.0, 5.0, 5.0], [5.0, 5.0, 5.0], [8.0, 30.0, 8.0], [20.0, 30.0, 8.0],
                     [8.0, 33.0, 20.0], [nextafter(20.0, 100.0), 33.0, 20.0]]
    mbox = dict(dim=3, mn=[0.0, 0.0, 0.0], mx=[36.0, 36.0, 48.0], interaction_range=12.0)
    build("morse_newton_3d_t2", **mbox, n_threads=2, dt=0.05, order=2,
          mech={"kind": "newton", "damping": 0.4, "mass": 1.3}, inter=dict(MORSE, strength=0.8), n_agents=160, seed=51,
          steps=[1, 2, 3, 10], vel_scale=20.0, crafted=morse_crafted)
    build("morse_newton_2d_t3", dim=2, mn=[0.0, 0.0], mx=[60.0, 84.0], interaction_range=12.0, n_threads=3, dt=0.05,
          order=2, mech={"kind": "newton", "damping": 0.4, "mass": 1.3}, inter=dict(MORSE, strength=0.8), n_agents=140,
          seed=52, steps=[1, 2, 3, 10], vel_scale=20.0)
    build("brownian_morse_3d_t2", **mbox, n_threads=2, dt=0.01, order=2,
          mech={"kind": "brownian", "diffusion": 0.3, "kt": 0.3}, inter=MORSE, n_agents=150, seed=53,
          steps=[1, 2, 3, 8], noise=True, crafted=morse_crafted)
    build("brownian_morse_hetero_t3", **mbox, n_threads=3, dt=0.01, order=2,
          mech={"kind": "brownian", "diffusion": 0.3, "kt": 0.3}, inter=dict(MORSE, strength=0.5), n_agents=120,
          seed=57, steps=[1, 4], noise=True, hetero=True)
    build("langevin_morse_2d_t2", dim=2, mn=[0.0, 0.0], mx=[48.0, 60.0], interaction_range=12.0, n_threads=2, dt=0.02,
          order=2, mech={"kind": "langevin", "mass": 1.5, "damping": 0.7, "kt": 0.4}, inter=dict(MORSE, strength=0.3),
          n_agents=120, seed=54, steps=[1, 2, 3, 8], vel_scale=5.0, noise=True)
    build("bound_lj_3d_t2", dim=3, mn=[0.0, 0.0, 0.0], mx=[30.0, 30.0, 30.0], interaction_range=10.0, n_threads=2,
          dt=0.01, order=2, mech={"kind": "newton", "damping": 1.0, "mass": 1.0},
          inter={"kind": "bound_lj", "epsilon": 0.5, "sigma": 3.0, "bound": 4.0, "cutoff": 10.0}, n_agents=110, seed=55,
          steps=[1, 2, 5], vel_scale=10.0,
          crafted=[[4.0, 4.0, 4.0], [14.0, 4.0, 4.0], [4.0, 8.0, 14.0], [nextafter(14.0, 100.0), 8.0, 14.0]])
    build("mie_3d_t3", dim=3, mn=[0.0, 0.0, 0.0], mx=[30.0, 30.0, 40.0], interaction_range=10.0, n_threads=3, dt=0.01,
          order=2, mech={"kind": "newton", "damping": 1.0, "mass": 1.0},
          inter={"kind": "mie", "radius": 2.0, "strength": 0.3, "bound": 5.0, "cutoff": 10.0, "en": 4.0, "em": 2.0},
          n_agents=110, seed=56, steps=[1, 2, 5], vel_scale=10.0, hetero=True)


if __name__ == "__main__":
    main()
