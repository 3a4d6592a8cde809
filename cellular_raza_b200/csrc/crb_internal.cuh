// crb_internal.cuh — data structures shared by the kernels and the C-ABI host code.
//
// Layout in HBM: agents are f64 structure-of-arrays, physically sorted by
// VoxelPlainIndex (ascending voxel, chili in-voxel order) so that the 3 voxels (z-1, z, z+1) of
// every (x, y) row of the 3^D neighbourhood are ONE contiguous index range.  Two copies of the
// state exist: A (sorted, read by the force pass) and B (integrated, unsorted, written by the
// force+integrate pass and consumed by the reorder pass).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crb.h"

namespace crb {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

// Voxel grid as seen by one device.  Internal axes (a0, a1, a2): a0 = slab axis (coordinate 0),
// a2 = fastest axis of the plain index (coordinate D-1), a1 = coordinate 1 for D == 3 and a dummy
// axis of extent 1 for D == 2.  plain = (i0 * n1 + i1) * n2 + i2 reproduces
// cartesian_cuboid_n.rs:260-270 / chili/datastructures.rs:113-121 for both D.
struct Grid {
    int D;
    int ng[3];      // global voxel counts per internal axis
    int L0;         // local planes held on this device (owned + ghost planes)
    int p0;         // global a0 index of local plane 0
    int own_lo;     // first owned local plane
    int own_hi;     // one past the last owned local plane
    int cax[3];     // coordinate axis of internal axis a (-1 for the dummy axis)
    double mn[3];   // CartesianCuboid::min per coordinate axis
    double mx[3];   // CartesianCuboid::max
    double dx[3];   // CartesianCuboid::dx
    // chili subdomain chunks over the GLOBAL plain index (cartesian_cuboid_n.rs:500-583)
    u64 sm_n, sm_avg, sm_switch, sm_nsub;
    __host__ __device__ u64 nvox_local() const { return (u64)L0 * ng[1] * ng[2]; }
    __host__ __device__ u64 nvox_global() const { return (u64)ng[0] * ng[1] * ng[2]; }
    __host__ __device__ u64 sub_of(u64 plain) const {
        if (plain < sm_switch) return plain / sm_avg;
        u64 a = sm_avg - 1 > 1 ? sm_avg - 1 : 1;
        return (plain - sm_switch) / a + sm_n;
    }
    __host__ __device__ u64 sub_first_voxel(u64 s) const {
        if (s <= sm_n) return s * sm_avg;
        return sm_switch + (s - sm_n) * (sm_avg - 1);
    }
};

// Structure-of-arrays agent state.  V = vertices per agent (1 for point agents); component c of
// vertex v of agent i lives at pos[c][v * cap + i].
struct State {
    double* pos[3];
    double* vel[3];
    double* h1x[3];  // newest entry of AuxStorageMechanics::positions  (increment dx_k)
    double* h1v[3];  // newest entry of AuxStorageMechanics::velocities (increment dv_k)
    double* h0x[3];  // older entry; only refreshed on the last step of a crb_step batch
    double* h0v[3];
    double* mp[CRB_MAX_MECH_PARAMS];
    double* ip[CRB_MAX_INTER_PARAMS];
    u64* id;
    u8* species;
    u8* hlen;  // entries held by the ring buffers (0..solver_order)
    // FP32 filter copy written by the reorder pass: {x - min_x, y - min_y, z - min_z (0 in 2-D),
    // bit pattern of the local voxel index}.  Only read by the conservative distance pre-test of
    // k_force_tile; never enters a result.
    float4* posf;
    u32* meta;  // species | (voxel index along the fastest axis & 0xff) << 8  (k_force_tile drain)
};

struct Uniform {
    double mp[CRB_MAX_MECH_PARAMS];
    double ip[CRB_MAX_INTER_PARAMS];
};

// sortkey layout: bit 63 = moved this step; bit 62 = moved in from ANOTHER chili subdomain; bits 20..61 =
// previous global VoxelPlainIndex; bits 0..19 = previous rank inside that voxel.  Ascending sortkey inside a
// voxel reproduces chili's order: stayers in place, then the movers of the voxel's own subdomain by (source
// voxel, in-voxel order) — appended by sort_cells_in_voxels_step_1 — then the movers received from other
// subdomains, sorted by source voxel — appended by step_2 (chili/update_mechanics.rs:427-476, 492-511).
#define CRB_SK_MOVED (1ull << 63)
#define CRB_SK_REMOTE (1ull << 62)
#define CRB_SK_VOXEL(sk) (((sk) & ~(CRB_SK_MOVED | CRB_SK_REMOTE)) >> CRB_SK_RANK_BITS)
#define CRB_SK_RANK_BITS 20
#define CRB_SK_RANK_MASK ((1ull << CRB_SK_RANK_BITS) - 1)

struct KParams {
    Grid g;
    State A, B;
    Uniform u;
    const u64* n_dev;  // != nullptr: the owned-agent count lives on the device (crb_slab_step), n bounds the grid
    u64 n;            // owned agents in A
    u64 first;        // index of the first owned agent (ghost capacity G; 0 on a single device)
    u64 cap;          // allocated agents per array (owned capacity + 2 G)
    u32* key;         // local voxel index of every agent of A (sorted ascending)
    u32* start;       // exclusive scan of per-voxel counts of A, nvox_local + 1 entries
    u32* newkey;      // local voxel index after the step (per old slot)
    u32* rank_arb;    // arrival rank in the new voxel (atomicAdd order, fixed up by the reorder)
    u64* sortkey;     // see above
    u32* count_new;   // per-voxel histogram of newkey
    double* force_out;  // debug tap: [c][cap]
    const double* noise;  // z table (device) or nullptr
    u64 noise_ids;
    u64 noise_cursor;
    u64 iteration;
    u64 seed;
    int* err;         // [0] status, [1] latch flag, [2..3] agent id
    double dt;
    int mech, inter, order, weight_mode, noise_mode, save_prev;
    // DomainForce aspect (crb_domain_force_kind): linear spring -dk[c] * (pos[c] - dx0[c])
    int dforce;
    double dk[3], dx0[3];
};

// ---- slab exchange buffers ----------------------------------------------------------
// One halo buffer carries the boundary plane of a slab: every column at a fixed offset (capacity
// gcap agents) so that the receiver needs no size from the host.
struct HaloLayout {
    u64 gcap, plane;  // agent capacity, voxels per plane
    int D, n_ip;      // n_ip == 0 when the parameters are uniform
    u64 off_rel;      // u32 rel[plane + 1]: index of the first agent of every voxel relative to the plane
    u64 off_pos[3];   // f64
    u64 off_posf;     // float4
    u64 off_meta;     // u32
    u64 off_sp;       // u8
    u64 off_ip[CRB_MAX_INTER_PARAMS];
    u64 bytes;
    u64 mode;  // parameter mode of the sender: 0 = per-agent columns, else hash of the uniform row | 1
    // header: u64 n at byte 0, u64 mode at byte 8
};
// One migration buffer carries complete agents (dynamic + static state) that changed slab.
struct MigLayout {
    u64 mcap;
    int D, n_mp, n_ip;
    u64 off_dyn[18];  // f64 columns: pos, vel, h1x, h1v, h0x, h0v (D each)
    u64 off_mp[CRB_MAX_MECH_PARAMS], off_ip[CRB_MAX_INTER_PARAMS];
    u64 off_id, off_sortkey, off_gvox;  // u64
    u64 off_hlen, off_sp;               // u8
    u64 bytes;
    u64 mode;  // as HaloLayout::mode
    // header: u64 n at byte 0, u64 mode at byte 8
};

#define CRB_KEY_GONE 0xffffffffu  // newkey of an agent that left this slab (skipped by the reorder)

}  // namespace crb
