"""GPU: the CUDA path (through the C ABI) against the reference-derived fixtures tests/golden/ref_derived_*.json.

The fixtures come from tests/ref_transcript.py, an independent literal transcription of the reference's Rust
(tools/make_ref_derived.py); neither the oracle nor the device code was consulted when it was written.  Checked per
fixture: forces before the first step (1e-12, metric of SURVEY.md 8d), positions / velocities at the checkpoints,
plain voxel index of every agent (exact), the save order = ascending voxel / chili in-voxel order incl. the
local-then-remote order of movers for n_threads > 1 (exact), ring-buffer fill, migration sets (exact)."""
import pytest

import ref_derived
from cellular_raza_b200.device import DeviceSimulation

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ref_derived.names())
def test_device_equals_reference_transcript(crb_lib, name):
    fix = ref_derived.load(name)
    with DeviceSimulation(ref_derived.config(fix)) as dev:
        ferr, serr = ref_derived.check(dev, fix)
    assert ferr <= 1e-12 and serr <= 1e-11


@pytest.mark.parametrize("kernel", ["agent", "tile", "tile2", "tile3"])
@pytest.mark.parametrize("name", ["cell_sorting_3d_t3", "cell_sorting_slab_t5", "morse_newton_2d_t3", "brownian_morse_3d_t2",
                                  "bound_lj_3d_t2", "mover_order_t2", "mover_order_t4"])
def test_every_force_kernel_equals_reference_transcript(crb_lib, monkeypatch, name, kernel):
    """the same fixtures through each force kernel (the planner would pick one by population size)"""
    monkeypatch.setenv("CRB_FORCE_KERNEL", kernel)
    fix = ref_derived.load(name)
    with DeviceSimulation(ref_derived.config(fix)) as dev:
        ref_derived.check(dev, fix)
