#!/usr/bin/env python
"""Small populations through every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

cell_sorting on the 2 x 2-row tile kernel (k_force_tile2, shared-memory staging + per-lane FIFOs), Brownian + Morse on
the same kernel, bacterial rods with divisions (k_rod_force / k_rod_update / k_rod_divide), and two slabs on one device
through crb_slab_step (peer-memory exchange kernels, device-side flags).  No oracle, no torch."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("CRB_SLAB_TIMEOUT_MS", "60000")

import numpy as np  # noqa: E402

from cellular_raza_b200 import _abi, workloads  # noqa: E402
from cellular_raza_b200.device import DeviceSimulation  # noqa: E402
from cellular_raza_b200.rods import RodSimulation  # noqa: E402
from cellular_raza_b200.slabs import LocalSlabGroup  # noqa: E402

which = sys.argv[1:] or ["tile2", "morse", "rods", "slabs"]
if "tile2" in which:
    os.environ["CRB_FORCE_KERNEL"] = "tile2"
    w = workloads.cell_sorting(n_cells=6000, domain_size=170.0)
    sim = DeviceSimulation(w.config())
    sim.set_uniform_params(w.mech_params[0], w.inter_params[0])
    sim.set_agents(w.pos, w.vel, None, None, w.species)
    sim.step(3)
    print("tile2", sim.force_kernel_name, sim.n_agents, float(np.abs(sim.get_agents()[0]).sum()))
    sim.close()
if "morse" in which:
    os.environ["CRB_FORCE_KERNEL"] = "tile2"
    w = workloads.brownian_morse(6000, 12.0 * 9, noise_mode=_abi.CRB_NOISE_PHILOX)
    sim = DeviceSimulation(w.config())
    sim.set_uniform_params(w.mech_params[0], w.inter_params[0])
    sim.set_agents(w.pos, w.vel, None, None, w.species)
    sim.step(3)
    print("morse", sim.force_kernel_name, sim.n_agents, float(np.abs(sim.get_agents()[0]).sum()))
    sim.close()
os.environ.pop("CRB_FORCE_KERNEL", None)
if "rods" in which:
    w = workloads.bacterial_rods(6, 5, rods_per_tile=6, spring_length_range=(5.9, 5.999), capacity_factor=3.0)
    sim = RodSimulation(w.config())
    sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params)
    sim.step(4)
    print("rods", sim.n_agents, "from", w.n)
    sim.close()
if "slabs" in which:
    w = workloads.cell_sorting(n_cells=6000, domain_size=170.0)
    grp = LocalSlabGroup(lambda: w.config(capacity=6000), 2, transport="p2p")
    grp.set_uniform_params(w.mech_params[0], w.inter_params[0])
    grp.set_agents(w.pos, w.vel, None, None, w.species)
    grp.step(3)
    print("slabs", sum(r.n_agents for r in grp.ranks))
    grp.close()
print("sanitize_run done")
