"""Kernel-level checks on the device: the reciprocal square root used by every potential, and the
two force kernels (per-agent over global memory / shared-memory tiled) against each other."""
import ctypes as C
import os

import numpy as np
import pytest

from cellular_raza_b200 import _abi, workloads
from cellular_raza_b200.device import DeviceSimulation

pytestmark = pytest.mark.gpu


def test_rsqrt_accuracy(crb_lib):
    rng = np.random.default_rng(3)
    x = np.concatenate([
        rng.uniform(1e-6, 4e3, 200_000),
        10.0 ** rng.uniform(-300, 300, 100_000),
        np.array([1.0, 4.0, 0.25, 324.0, 2.0, 3.0, 1e-300, 1e300]),
    ])
    y = np.empty_like(x)
    rc = crb_lib.crb_debug_rsqrt(0, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), x.size)
    assert rc == 0
    ref = 1.0 / np.sqrt(x.astype(np.longdouble))
    rel = np.abs((y.astype(np.longdouble) - ref) / ref).astype(np.float64)
    assert rel.max() <= 2.5e-16, f"max rel err {rel.max():.3e} at x={x[rel.argmax()]!r}"


def _forces(w, kernel, monkeypatch):
    monkeypatch.setenv("CRB_FORCE_KERNEL", kernel)
    sim = DeviceSimulation(w.config())
    sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    f = sim.debug_forces()
    ids = sim.get_agents()[2]
    sim.close()
    return f[np.argsort(ids)]


@pytest.mark.parametrize("make", [
    lambda: workloads.cell_sorting(),
    lambda: workloads.cell_sorting(n_threads=4),
    lambda: workloads.cell_sorting(n_cells=3000, domain_size=90.0, dim=2),
    lambda: workloads.cell_sorting_scaled(200_000, 7.1, n_threads=3),
    lambda: workloads.cell_sorting_scaled(100_000, 2.5),
    lambda: workloads.brownian_morse(50_000, 12.0 * 23, noise_mode=_abi.CRB_NOISE_NONE),
])
def test_tile_kernel_equals_agent_kernel(crb_lib, monkeypatch, make):
    """same candidates, same coef(), same ascending-index summation order => identical forces"""
    w = make()
    fa = _forces(w, "agent", monkeypatch)
    scale = np.abs(fa).max()
    for kernel in ("tile", "tile2"):
        ft = _forces(w, kernel, monkeypatch)
        assert np.abs(fa - ft).max() <= 1e-14 * scale, kernel
        assert np.array_equal(fa, ft), f"{kernel} and per-agent kernels should agree bitwise"


def test_tile_kernel_dense_voxels_fall_back(crb_lib, monkeypatch):
    """a cluster whose 27-voxel neighbourhood exceeds the staging capacity takes the per-agent path
    inside k_force_tile; results must not change."""
    w = workloads.cell_sorting(n_cells=6000, domain_size=54.0)  # 3^3 voxels, 222 agents per voxel
    fa = _forces(w, "agent", monkeypatch)
    for kernel in ("tile", "tile2"):
        assert np.array_equal(fa, _forces(w, kernel, monkeypatch)), kernel
