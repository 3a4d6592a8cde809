"""ctypes binding of ``include/crb.h`` — the C-ABI of the B200 mechanics/interaction backend.

This is the Python twin of the ``extern "C"`` block a Rust host crate would declare (see
INTEGRATION.md; precedent in the reference: cellular_raza-core/src/backend/cara/mod.rs:24-62).
There is no CPU fallback: if ``libcrb200.so`` has not been built (``python -m
cellular_raza_b200.build`` or ``__graft_entry__.build()``) loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

CRB_ABI_VERSION = 7
CRB_MAX_DIM = 3

# crb_status
CRB_OK = 0
CRB_ERR_CALC = -1
CRB_ERR_BOUNDARY = -2
CRB_ERR_INDEX = -3
CRB_ERR_DIVISION = -4
CRB_ERR_DECOMPOSE = -5
CRB_ERR_INVALID_ARG = -16
CRB_ERR_CAPACITY = -17
CRB_ERR_CUDA = -18
CRB_ERR_STATE = -19
CRB_ERR_TIMEOUT = -20

# crb_mechanics_kind
CRB_MECH_NEWTON_DAMPED = 0
CRB_MECH_BROWNIAN = 1
CRB_MECH_LANGEVIN = 2
CRB_MECH_ROD = 3
# crb_interaction_kind
CRB_INT_CELL_SORTING = 0
CRB_INT_CELL_SORTING_BENCH = 1
CRB_INT_MORSE = 2
CRB_INT_BOUND_LJ = 3
CRB_INT_MIE = 4
CRB_INT_NONE = 5
CRB_DOMAIN_FORCE_NONE = 0
CRB_DOMAIN_FORCE_SPRING = 1
CRB_DOMAIN_FORCE_ROD_GEL = 2
# crb_pair_weight_mode
CRB_WEIGHT_CHILI = 0
CRB_WEIGHT_PHYSICAL = 1
# crb_noise_mode
CRB_NOISE_NONE = 0
CRB_NOISE_TABLE = 1
CRB_NOISE_PHILOX = 2

N_MECH_PARAMS = {CRB_MECH_NEWTON_DAMPED: 2, CRB_MECH_BROWNIAN: 2, CRB_MECH_LANGEVIN: 3, CRB_MECH_ROD: 8}
N_INTER_PARAMS = {
    CRB_INT_CELL_SORTING: 3,
    CRB_INT_CELL_SORTING_BENCH: 3,
    CRB_INT_MORSE: 4,
    CRB_INT_BOUND_LJ: 4,
    CRB_INT_MIE: 6,
    CRB_INT_NONE: 0,
}


class CrbConfig(C.Structure):
    """``struct crb_config`` (include/crb.h)."""

    _fields_ = [
        ("abi_version", C.c_uint32),
        ("dim", C.c_uint32),
        ("mechanics_kind", C.c_uint32),
        ("interaction_kind", C.c_uint32),
        ("solver_order", C.c_uint32),
        ("pair_weight_mode", C.c_uint32),
        ("ref_n_subdomains", C.c_uint32),
        ("noise_mode", C.c_uint32),
        ("rod_n_vertices", C.c_uint32),
        ("device", C.c_uint32),
        ("rank", C.c_uint32),
        ("world_size", C.c_uint32),
        ("n_voxels", C.c_uint64 * CRB_MAX_DIM),
        ("domain_min", C.c_double * CRB_MAX_DIM),
        ("domain_max", C.c_double * CRB_MAX_DIM),
        ("dx", C.c_double * CRB_MAX_DIM),
        ("dt", C.c_double),
        ("rng_seed", C.c_uint64),
        ("capacity", C.c_uint64),
        ("ghost_capacity", C.c_uint64),
        ("migration_capacity", C.c_uint64),
        ("domain_force_kind", C.c_uint32),
        ("reserved0", C.c_uint32),
        ("domain_force_strength", C.c_double * CRB_MAX_DIM),
        ("domain_force_anchor", C.c_double * CRB_MAX_DIM),
    ]


class SimulationError(RuntimeError):
    """Mirror of ``SimulationError`` (cellular_raza-core/src/backend/chili/errors.rs:73-121)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{self.variant(code)}: {message}")
        self.code = code
        self.message = message

    @staticmethod
    def variant(code: int) -> str:
        return {
            CRB_ERR_CALC: "CalcError",
            CRB_ERR_BOUNDARY: "BoundaryError",
            CRB_ERR_INDEX: "IndexError",
            CRB_ERR_DIVISION: "DivisionError",
            CRB_ERR_DECOMPOSE: "DecomposeError",
            CRB_ERR_INVALID_ARG: "InvalidArgument",
            CRB_ERR_CAPACITY: "CapacityError",
            CRB_ERR_CUDA: "CudaError",
            CRB_ERR_STATE: "StateError",
            CRB_ERR_TIMEOUT: "OtherThreadError",
        }.get(code, f"Error({code})")


_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "lib" / "libcrb200.so"

# every symbol include/crb.h declares: (name, restype, argtypes)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p
SYMBOLS = [
    ("crb_create", C.c_int, [C.POINTER(CrbConfig), C.POINTER(_vp)]),
    ("crb_destroy", None, [_vp]),
    ("crb_last_error", C.c_char_p, [_vp]),
    ("crb_abi_version", C.c_uint32, []),
    ("crb_n_mech_params", C.c_uint32, [C.c_uint32]),
    ("crb_n_inter_params", C.c_uint32, [C.c_uint32]),
    ("crb_set_agents", C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("crb_set_agents_async", C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("crb_set_uniform_params", C.c_int, [_vp, _vp, _vp]),
    ("crb_n_agents", C.c_uint64, [_vp]),
    ("crb_get_agents", C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp]),
    ("crb_get_params", C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp]),
    ("crb_get_aux", C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp]),
    ("crb_set_aux", C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp]),
    ("crb_get_identifiers", C.c_int, [_vp, C.c_uint64, _vp, _vp]),
    ("crb_n_divisions", C.c_uint64, [_vp]),
    ("crb_debug_neighbor_counts", C.c_int, [_vp, C.c_uint64, _vp]),
    ("crb_set_noise", C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64]),
    ("crb_step", C.c_int, [_vp, C.c_uint64]),
    ("crb_step_async", C.c_int, [_vp, C.c_uint64]),
    ("crb_sync", C.c_int, [_vp]),
    ("crb_iteration", C.c_uint64, [_vp]),
    ("crb_set_iteration", C.c_int, [_vp, C.c_uint64]),
    ("crb_force_kernel_name", C.c_char_p, [_vp]),
    ("crb_snapshot_async", C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("crb_snapshot_wait", C.c_int, [_vp, _u64p]),
    ("crb_host_alloc", C.c_void_p, [C.c_uint64]),
    ("crb_host_free", None, [C.c_void_p]),
    ("crb_slab_halo_bytes", C.c_uint64, [_vp]),
    ("crb_slab_migration_bytes", C.c_uint64, [_vp]),
    ("crb_slab_info", C.c_int, [_vp, _u64p]),
    ("crb_slab_halo_export", C.c_int, [_vp, _vp, _vp]),
    ("crb_slab_halo_import", C.c_int, [_vp, _vp, _vp]),
    ("crb_slab_compute", C.c_int, [_vp, C.c_int]),
    ("crb_slab_compute_interior", C.c_int, [_vp, C.c_int]),
    ("crb_slab_compute_boundary", C.c_int, [_vp]),
    ("crb_slab_migration_export", C.c_int, [_vp, _vp, _vp, _u64p]),
    ("crb_slab_migration_counts", C.c_int, [_vp, _u64p]),
    ("crb_slab_migration_import", C.c_int, [_vp, _vp, _vp, _u64p]),
    ("crb_slab_finish", C.c_int, [_vp, C.c_int]),
    ("crb_slab_arena_bytes", C.c_uint64, [_vp]),
    ("crb_slab_arena", C.c_void_p, [_vp]),
    ("crb_slab_ipc_handle", C.c_int, [_vp, _vp]),
    ("crb_slab_connect", C.c_int, [_vp, _vp, _vp]),
    ("crb_slab_connect_ipc", C.c_int, [_vp, _vp, _vp]),
    ("crb_slab_step", C.c_int, [_vp, C.c_uint64]),
    ("crb_slab_step_async", C.c_int, [_vp, C.c_uint64]),
    ("crb_debug_forces", C.c_int, [_vp, C.c_uint64, _vp]),
    ("crb_get_migrations", C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _u64p]),
    ("crb_get_voxel_counts", C.c_int, [_vp, C.c_uint64, _vp]),
    ("crb_count_candidate_pairs", C.c_int, [_vp, _u64p]),
    ("crb_kernel_launches", C.c_uint64, [_vp]),
    ("crb_enable_timers", C.c_int, [_vp, C.c_int]),
    ("crb_reset_timers", C.c_int, [_vp]),
    ("crb_stage_time_ms", C.c_int, [_vp, C.c_uint32, _f64p, _u64p]),
    ("crb_stream", C.c_void_p, [_vp]),
    ("crb_measure_fp64_peak", C.c_int, [C.c_uint32, _f64p, _f64p]),
    ("crb_debug_rsqrt", C.c_int, [C.c_uint32, _vp, _vp, C.c_uint64]),
    ("crb_debug_exp", C.c_int, [C.c_uint32, _vp, _vp, C.c_uint64]),
    ("crb_decomp_res", C.c_int, [C.c_uint64, C.c_uint64, _u64p, _u64p, _u64p]),
    ("crb_domain_from_interaction_range", C.c_int, [C.c_uint32, _f64p, _f64p, C.c_double, _u64p, _f64p]),
    ("crb_domain_from_n_voxels", C.c_int, [C.c_uint32, _f64p, _f64p, _u64p, _f64p]),
]

_lib = None


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Load ``libcrb200.so`` and declare every entry point. Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path is not None else LIB_PATH
    if not p.exists():
        raise ImportError(
            f"{p} not found: the CUDA library of cellular_raza_b200 has not been built. "
            "Run `python -m cellular_raza_b200.build` (needs nvcc). There is no CPU fallback."
        )
    lib = C.CDLL(str(p))
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.crb_abi_version() != CRB_ABI_VERSION:
        raise ImportError(f"{p}: ABI version {lib.crb_abi_version()} != {CRB_ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib
