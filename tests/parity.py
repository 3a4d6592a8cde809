"""Helpers shared by the parity tests: drive the device (through the C ABI) and the oracle with the
same inputs and compare them with the metrics of SURVEY.md §8d."""
from __future__ import annotations

import numpy as np

from cellular_raza_b200.device import DeviceSimulation
from oracle.chili_oracle import ChiliOracle


def make_pair(w, n_threads_oracle=1, **cfg_overrides):
    cfg = w.config(**cfg_overrides)
    dev = DeviceSimulation(cfg)
    ora = ChiliOracle(w.config(**cfg_overrides), n_threads=n_threads_oracle)
    for sim in (dev, ora):
        sim.set_agents(w.pos, w.vel, w.mech_params, w.inter_params, w.species)
    return dev, ora


def by_id(ids, *arrays):
    order = np.argsort(ids, kind="stable")
    return [a[order] for a in arrays]


def assert_same_membership(dev, ora):
    """voxel keys and per-voxel membership sets are bit-exact; with a single emulated subdomain
    the in-voxel ORDER is chili's as well."""
    _, _, ids_d, vox_d = dev.get_agents()
    _, _, ids_o, vox_o = ora.get_agents()
    assert len(ids_d) == len(ids_o)
    (vd,) = by_id(ids_d, vox_d)
    (vo,) = by_id(ids_o, vox_o)
    np.testing.assert_array_equal(vd, vo)
    np.testing.assert_array_equal(dev.get_voxel_counts(), ora.get_voxel_counts())


def force_errors(dev, ora):
    """per-agent |f_dev - f_ref|_inf / max(|f_ref|_inf, sum_j |f_ij|_inf)  (SURVEY.md §8d)."""
    _, _, ids_d, _ = dev.get_agents()
    _, _, ids_o, _ = ora.get_agents()
    fd = dev.debug_forces()
    fo, ab = ora.forces()
    (fd,) = by_id(ids_d, fd)
    fo, ab = by_id(ids_o, fo, ab)
    num = np.abs(fd - fo).max(axis=1)
    den = np.maximum(np.abs(fo).max(axis=1), ab)
    den = np.where(den == 0, 1.0, den)
    return num / den, fd, fo


def state_errors(dev, ora, scale):
    pd, vd, ids_d, xd = dev.get_agents()
    po, vo, ids_o, xo = ora.get_agents()
    pd, vd, xd = by_id(ids_d, pd, vd, xd)
    po, vo, xo = by_id(ids_o, po, vo, xo)
    return np.abs(pd - po).max() / scale, np.abs(vd - vo).max(), (xd != xo)
