//! `extern "C"` block mirroring include/crb.h 1:1 (ABI version 4).  Replaces the shape of
//! cellular_raza-core/src/backend/cara/mod.rs:5-62 (`#[repr(C)] Runner` + `extern "C"`).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const CRB_ABI_VERSION: u32 = 7;

pub const CRB_OK: c_int = 0;
pub const CRB_ERR_CALC: c_int = -1;
pub const CRB_ERR_BOUNDARY: c_int = -2;
pub const CRB_ERR_INDEX: c_int = -3;
pub const CRB_ERR_DIVISION: c_int = -4;
pub const CRB_ERR_DECOMPOSE: c_int = -5;
pub const CRB_ERR_INVALID_ARG: c_int = -16;
pub const CRB_ERR_CAPACITY: c_int = -17;
pub const CRB_ERR_CUDA: c_int = -18;
pub const CRB_ERR_STATE: c_int = -19;
pub const CRB_ERR_TIMEOUT: c_int = -20;

pub const CRB_MECH_NEWTON_DAMPED: u32 = 0;
pub const CRB_MECH_BROWNIAN: u32 = 1;
pub const CRB_MECH_LANGEVIN: u32 = 2;
pub const CRB_MECH_ROD: u32 = 3;

pub const CRB_INT_CELL_SORTING: u32 = 0;
pub const CRB_INT_CELL_SORTING_BENCH: u32 = 1;
pub const CRB_INT_MORSE: u32 = 2;
pub const CRB_INT_BOUND_LJ: u32 = 3;
pub const CRB_INT_MIE: u32 = 4;
pub const CRB_INT_NONE: u32 = 5;

pub const CRB_DOMAIN_FORCE_NONE: u32 = 0;
pub const CRB_DOMAIN_FORCE_SPRING: u32 = 1;
pub const CRB_DOMAIN_FORCE_ROD_GEL: u32 = 2;

pub const CRB_WEIGHT_CHILI: u32 = 0;
pub const CRB_WEIGHT_PHYSICAL: u32 = 1;

pub const CRB_NOISE_NONE: u32 = 0;
pub const CRB_NOISE_TABLE: u32 = 1;
pub const CRB_NOISE_PHILOX: u32 = 2;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct crb_config {
    pub abi_version: u32,
    pub dim: u32,
    pub mechanics_kind: u32,
    pub interaction_kind: u32,
    pub solver_order: u32,
    pub pair_weight_mode: u32,
    pub ref_n_subdomains: u32,
    pub noise_mode: u32,
    pub rod_n_vertices: u32,
    pub device: u32,
    pub rank: u32,
    pub world_size: u32,
    pub n_voxels: [u64; 3],
    pub domain_min: [f64; 3],
    pub domain_max: [f64; 3],
    pub dx: [f64; 3],
    pub dt: f64,
    pub rng_seed: u64,
    pub capacity: u64,
    pub ghost_capacity: u64,
    pub migration_capacity: u64,
    pub domain_force_kind: u32,
    pub reserved0: u32,
    pub domain_force_strength: [f64; 3],
    pub domain_force_anchor: [f64; 3],
}

#[repr(C)]
pub struct crb_sim {
    _private: [u8; 0],
}

extern "C" {
    pub fn crb_abi_version() -> u32;
    pub fn crb_n_mech_params(kind: u32) -> u32;
    pub fn crb_n_inter_params(kind: u32) -> u32;
    pub fn crb_create(cfg: *const crb_config, out: *mut *mut crb_sim) -> c_int;
    pub fn crb_destroy(sim: *mut crb_sim);
    pub fn crb_last_error(sim: *const crb_sim) -> *const c_char;
    pub fn crb_set_agents(sim: *mut crb_sim, n: u64, pos: *const f64, vel: *const f64, mech_params: *const f64,
                          inter_params: *const f64, species: *const u8, ids: *const u64) -> c_int;
    pub fn crb_set_agents_async(sim: *mut crb_sim, n: u64, pos: *const f64, vel: *const f64, mech_params: *const f64,
                                inter_params: *const f64, species: *const u8, ids: *const u64) -> c_int;
    pub fn crb_set_uniform_params(sim: *mut crb_sim, mech_row: *const f64, inter_row: *const f64) -> c_int;
    pub fn crb_n_agents(sim: *const crb_sim) -> u64;
    pub fn crb_get_agents(sim: *mut crb_sim, cap: u64, pos: *mut f64, vel: *mut f64, ids: *mut u64,
                          voxel: *mut u64) -> c_int;
    pub fn crb_get_params(sim: *mut crb_sim, cap: u64, mech_params: *mut f64, inter_params: *mut f64,
                          species: *mut u8) -> c_int;
    pub fn crb_get_aux(sim: *mut crb_sim, cap: u64, hist_dx: *mut f64, hist_dv: *mut f64, hist_len: *mut u32) -> c_int;
    pub fn crb_set_aux(sim: *mut crb_sim, n: u64, hist_dx: *const f64, hist_dv: *const f64, hist_len: *const u32) -> c_int;
    pub fn crb_set_noise(sim: *mut crb_sim, z: *const f64, n_steps: u64, n_ids: u64) -> c_int;
    pub fn crb_get_identifiers(sim: *mut crb_sim, cap: u64, ident: *mut u64, parent: *mut u64) -> c_int;
    pub fn crb_n_divisions(sim: *const crb_sim) -> u64;
    pub fn crb_step(sim: *mut crb_sim, n_steps: u64) -> c_int;
    pub fn crb_step_async(sim: *mut crb_sim, n_steps: u64) -> c_int;
    pub fn crb_sync(sim: *mut crb_sim) -> c_int;
    pub fn crb_iteration(sim: *const crb_sim) -> u64;
    pub fn crb_set_iteration(sim: *mut crb_sim, iteration: u64) -> c_int;
    // storage snapshots that do not stop the stepping loop
    pub fn crb_snapshot_async(sim: *mut crb_sim, cap: u64, pos: *mut f64, vel: *mut f64, ids: *mut u64,
                              voxel: *mut u64, hist_dx: *mut f64, hist_dv: *mut f64, hist_len: *mut u32) -> c_int;
    pub fn crb_snapshot_wait(sim: *mut crb_sim, n_agents: *mut u64) -> c_int;
    pub fn crb_host_alloc(bytes: u64) -> *mut c_void;
    pub fn crb_host_free(p: *mut c_void);
    // slab decomposition (one handle per device; the caller moves the device buffers, e.g. ncclSend/ncclRecv)
    pub fn crb_slab_halo_bytes(sim: *const crb_sim) -> u64;
    pub fn crb_slab_migration_bytes(sim: *const crb_sim) -> u64;
    pub fn crb_slab_info(sim: *const crb_sim, out: *mut u64) -> c_int;
    pub fn crb_slab_halo_export(sim: *mut crb_sim, lo_buf: *mut c_void, hi_buf: *mut c_void) -> c_int;
    pub fn crb_slab_halo_import(sim: *mut crb_sim, lo_buf: *const c_void, hi_buf: *const c_void) -> c_int;
    pub fn crb_slab_compute(sim: *mut crb_sim, save_prev: c_int) -> c_int;
    pub fn crb_slab_compute_interior(sim: *mut crb_sim, save_prev: c_int) -> c_int;
    pub fn crb_slab_compute_boundary(sim: *mut crb_sim) -> c_int;
    pub fn crb_slab_migration_export(sim: *mut crb_sim, lo_buf: *mut c_void, hi_buf: *mut c_void, n_out: *mut u64) -> c_int;
    pub fn crb_slab_migration_counts(sim: *const crb_sim, n_out: *mut u64) -> c_int;
    pub fn crb_slab_migration_import(sim: *mut crb_sim, lo_buf: *const c_void, hi_buf: *const c_void, n_in: *mut u64) -> c_int;
    pub fn crb_slab_finish(sim: *mut crb_sim, save_prev: c_int) -> c_int;
    pub fn crb_slab_arena_bytes(sim: *const crb_sim) -> u64;
    pub fn crb_slab_arena(sim: *mut crb_sim) -> *mut c_void;
    pub fn crb_slab_ipc_handle(sim: *mut crb_sim, handle64: *mut c_void) -> c_int;
    pub fn crb_slab_connect(sim: *mut crb_sim, lo_arena: *mut c_void, hi_arena: *mut c_void) -> c_int;
    pub fn crb_slab_connect_ipc(sim: *mut crb_sim, lo_handle64: *const c_void, hi_handle64: *const c_void) -> c_int;
    pub fn crb_slab_step(sim: *mut crb_sim, n_steps: u64) -> c_int;
    pub fn crb_slab_step_async(sim: *mut crb_sim, n_steps: u64) -> c_int;
    // parity taps
    pub fn crb_debug_forces(sim: *mut crb_sim, cap: u64, force: *mut f64) -> c_int;
    pub fn crb_get_voxel_counts(sim: *mut crb_sim, cap: u64, counts: *mut u32) -> c_int;
    pub fn crb_count_candidate_pairs(sim: *mut crb_sim, n_pairs: *mut u64) -> c_int;
    pub fn crb_kernel_launches(sim: *const crb_sim) -> u64;
}
