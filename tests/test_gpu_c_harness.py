"""The C ABI driven from a plain C host: tests/c_harness.c is compiled with gcc against include/crb.h, linked to
libcrb200.so and run; the same population goes through the ctypes twin and both hosts must print / compute the same
bits (no non-Python host had ever called the boundary)."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from cellular_raza_b200 import _abi, build
from cellular_raza_b200.device import DeviceSimulation, make_config

ROOT = Path(__file__).resolve().parents[1]


def _compile(tmp_path):
    exe = tmp_path / "c_harness"
    lib_dir = build.LIB.parent
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"),
           str(ROOT / "tests" / "c_harness.c"), "-L", str(lib_dir), "-lcrb200", f"-Wl,-rpath,{lib_dir}", "-lm",
           "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_harness_compiles_against_the_header(tmp_path):
    """(no GPU needed) the header is plain C11 and the library resolves every symbol the harness uses"""
    build.build()
    _compile(tmp_path)


def _lcg_positions(n, L):
    state = 0x9E3779B97F4A7C15
    out = np.empty(n * 3)
    for k in range(n * 3):
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out[k] = (float(state >> 11) + 0.5) * (1.0 / 9007199254740992.0) * L
    return out.reshape(n, 3)


@pytest.mark.gpu
def test_c_host_and_python_host_get_the_same_bits(crb_lib, tmp_path):
    exe = _compile(tmp_path)
    n, steps, L = 2000, 10, 72.0
    res = subprocess.run([str(exe), str(n), str(steps)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "C_HARNESS OK" in res.stdout
    m = re.search(r"n=(\d+) steps=(\d+) ring=(\d+) pairs=(\d+) .*checksum=(\S+),(\S+),(\S+)", res.stdout)
    assert m, res.stdout
    assert int(m.group(1)) == n and int(m.group(2)) == steps and int(m.group(3)) == 2
    # the same population through the ctypes twin
    pos = _lcg_positions(n, L)
    nv = np.array([4, 4, 4], dtype=np.uint64)
    cfg = make_config(dim=3, mechanics_kind=_abi.CRB_MECH_NEWTON_DAMPED, interaction_kind=_abi.CRB_INT_CELL_SORTING,
                      domain_min=[0.0] * 3, domain_max=[L] * 3, n_voxels=nv, dx=[L / 4.0] * 3, dt=0.2, capacity=n,
                      ref_n_subdomains=4)
    with DeviceSimulation(cfg) as dev:
        dev.set_agents(pos, np.zeros((n, 3)), np.tile([2.0, 1.0], (n, 1)), np.tile([6.0, 2.0, 1.5], (n, 1)),
                       (np.arange(n) <= n // 2).astype(np.uint8))
        dev.step(steps)
        p, _, ids, _ = dev.get_agents()
        pairs = dev.count_candidate_pairs()
    by_id = p[np.argsort(ids)]
    sums = [0.0, 0.0, 0.0]
    for row in by_id:  # the C harness adds in identifier order
        for d in range(3):
            sums[d] += row[d]
    assert [float.fromhex(m.group(5 + d)) for d in range(3)] == sums
    assert int(m.group(4)) == pairs


# ---- rods + the asynchronous calls from C (tests/c_harness_rods.c) ---------------------------------------------------
def _compile_rods(tmp_path):
    exe = tmp_path / "c_harness_rods"
    lib_dir = build.LIB.parent
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"),
           str(ROOT / "tests" / "c_harness_rods.c"), "-L", str(lib_dir), "-lcrb200", f"-Wl,-rpath,{lib_dir}",
           "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def rods_population_of_the_c_harness(n):
    """the rods of c_harness_rods.c, same LCG, same operations in the same order (+ and * only)"""
    um, V = 1e-6, 5
    state = [0x9E3779B97F4A7C15]

    def u():
        state[0] = (state[0] * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return (float(state[0] >> 11) + 0.5) * (1.0 / 9007199254740992.0)

    pos = np.zeros((n, V, 3))
    mp = np.zeros((n, 8))
    ip = np.zeros((n, 4))
    for i in range(n):
        x0 = (20.0 + 340.0 * u()) * um
        y0 = (5.0 + 5.0 * u()) * um
        z0 = (10.0 + 25.0 * u()) * um
        for v in range(V):
            pos[i, v] = [x0 + 3.0 * um * float(v), y0 + 0.2 * um * (u() - 0.5), z0]
        mp[i] = [0.0, 10.0, 1.0 * um, (5.9 + 0.09 * u()) * um, 1.5, 0.1 * um, 6.0 * um, 0.1 * um]
        ip[i] = [3.0 * um, 0.5 / um, 5.0 * um, 0.1 * um * um]
    mx = [400.0 * um, 15.0 * um, 45.0 * um]
    nv = np.array([16, 1, 2], dtype=np.uint64)
    cfg = make_config(dim=3, mechanics_kind=_abi.CRB_MECH_ROD, interaction_kind=_abi.CRB_INT_MORSE,
                      domain_min=[0.0] * 3, domain_max=mx, n_voxels=nv, dx=[mx[k] / float(nv[k]) for k in range(3)],
                      dt=0.1, capacity=4 * n, ref_n_subdomains=8, rod_n_vertices=V)
    return cfg, pos, mp, ip


def test_c_harness_rods_compiles_against_the_header(tmp_path):
    build.build()
    _compile_rods(tmp_path)


def test_rods_population_of_the_c_harness_divides_in_the_oracle():
    """(CPU) the population the C host uploads does what the harness expects: rods divide within its 12 steps"""
    from oracle.rods_oracle import RodsOracle

    cfg, pos, mp, ip = rods_population_of_the_c_harness(48)
    o = RodsOracle(cfg)
    o.set_agents(pos, np.zeros_like(pos), mp, ip)
    o.step(12)
    assert 48 < o.n_agents <= 4 * 48


@pytest.mark.gpu
def test_c_host_drives_rods_through_the_asynchronous_calls(crb_lib, tmp_path):
    """crb_set_agents_async -> crb_step -> crb_snapshot_async / _wait for CRB_MECH_ROD from C; the synchronous ctypes
    twin gets the same bits (device order included)"""
    from cellular_raza_b200.rods import RodSimulation

    exe = _compile_rods(tmp_path)
    n, steps = 48, 12
    res = subprocess.run([str(exe), str(n), str(steps)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "C_HARNESS_RODS OK" in res.stdout
    m = re.search(r"n0=(\d+) n1=(\d+) steps=(\d+) division_ids=(\d+) checksum=(\S+),(\S+),(\S+)", res.stdout)
    assert m, res.stdout
    cfg, pos, mp, ip = rods_population_of_the_c_harness(n)
    dev = RodSimulation(cfg)
    dev.set_agents(pos, np.zeros_like(pos), mp, ip)
    dev.step(steps)
    p, _, ident, _, _ = dev.get_agents()
    assert int(m.group(2)) == dev.n_agents > n
    assert int(m.group(4)) == int((ident[:, 0] == 0).sum()) > 0
    sums = [0.0, 0.0, 0.0]
    for row in p.reshape(-1, 3):  # the C harness adds in device order
        for d in range(3):
            sums[d] += row[d]
    assert [float.fromhex(m.group(5 + d)) for d in range(3)] == sums
    dev.close()
