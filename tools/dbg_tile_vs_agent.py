import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cellular_raza_b200 import workloads
from cellular_raza_b200.device import DeviceSimulation
def forces(w, kernel):
    os.environ["CRB_FORCE_KERNEL"] = kernel
    sim = DeviceSimulation(w.config()); sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    f = sim.debug_forces(); ids = sim.get_agents()[2]; vox = sim.get_agents()[3]; sim.close()
    o = np.argsort(ids); return f[o], vox[o]
for name, w in [("2d", workloads.cell_sorting(n_cells=int(sys.argv[1]) if len(sys.argv) > 1 else 3000, domain_size=90.0, dim=2)),
                ("3d dense", workloads.cell_sorting(n_cells=6000, domain_size=54.0))]:
    fa, va = forces(w, "agent"); ft, vt = forces(w, "tile")
    bad = np.nonzero(np.abs(fa - ft).max(axis=1) > 1e-12 * np.abs(fa).max())[0]
    print(name, "n bad", len(bad), "of", len(fa), "bad voxels", np.unique(va[bad])[:30], "first", bad[:10])
    if len(bad): print(fa[bad[:3]], ft[bad[:3]])
