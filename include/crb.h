/*
 * crb.h — C ABI of the B200-native mechanics/interaction backend for cellular_raza.
 *
 * This header is the drop-in boundary.  It plays the role that
 *   cellular_raza-core/src/backend/cara/mod.rs:5-62   (#[repr(C)] Runner + extern "C" block)
 * plays for the placeholder CUDA backend of the reference: a Rust host crate that sits beside
 * `chili` binds exactly these symbols (see INTEGRATION.md for the `extern "C"` block and the
 * `build.rs`), and the generated main loop of
 *   cellular_raza-core-proc-macro/src/run_sim.rs:380-403
 * is replaced by `crb_step` between save events and `crb_get_agents` at PartialSave iterations.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.
 *  - every host pointer is borrowed for the duration of the call; outputs are written into
 *    caller-allocated buffers.  All device memory is owned by the opaque handle.
 *  - return value 0 on success, a negative crb_status otherwise (mirrors the SimulationError
 *    variants used on this path, cellular_raza-core/src/backend/chili/errors.rs:73-121);
 *    `crb_last_error` returns a human readable message for the last failure on that handle.
 *  - a handle is NOT thread-safe (one host thread per handle, one handle per GPU slab).
 *  - vectors are stored agent-major: pos[i*dim + d].
 */
#ifndef CRB_H
#define CRB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRB_ABI_VERSION 7

/* ---- status codes ------------------------------------------------------------------------- */
typedef enum crb_status {
    CRB_OK = 0,
    CRB_ERR_CALC = -1,        /* CalcError     (e.g. Mie potential at r == 0, interaction.rs:363-367) */
    CRB_ERR_BOUNDARY = -2,    /* BoundaryError (agent still outside after one reflection,
                                 cartesian_cuboid_n.rs:616-625)                                   */
    CRB_ERR_INDEX = -3,       /* IndexError    (voxel lookup failed at setup, domain.rs:372-377)  */
    CRB_ERR_DIVISION = -4,    /* DivisionError (rod division failed, bacterial_rods.rs:623-626)   */
    CRB_ERR_DECOMPOSE = -5,   /* DecomposeError (get_decomp_res found nothing)                    */
    CRB_ERR_INVALID_ARG = -16,
    CRB_ERR_CAPACITY = -17,   /* more agents than `capacity` (set_agents, division, migration)    */
    CRB_ERR_CUDA = -18,       /* CUDA runtime failure; message holds cudaGetErrorString           */
    CRB_ERR_STATE = -19,      /* call sequence violation (e.g. step before set_agents)            */
    CRB_ERR_TIMEOUT = -20     /* crb_slab_step: a neighbouring slab never delivered its halo / migration
                                 buffer (the reference's counterpart: OtherThreadError at the barrier,
                                 simulation_flow.rs:397-423)                                       */
} crb_status;

/* ---- model selection ---------------------------------------------------------------------- */
/* Mechanics kinds and their per-agent parameter rows (mech_params[i*n_mech + k]).
 * reference: cellular_raza-building-blocks/src/cell_building_blocks/mechanics.rs */
typedef enum crb_mechanics_kind {
    CRB_MECH_NEWTON_DAMPED = 0, /* [damping_constant, mass]                 mechanics.rs:158-173 */
    CRB_MECH_BROWNIAN = 1,      /* [diffusion_constant, kb_temperature]     mechanics.rs:336-357 */
    CRB_MECH_LANGEVIN = 2,      /* [mass, damping, kb_temperature]          mechanics.rs:434-463 */
    CRB_MECH_ROD = 3            /* [diffusion_constant, spring_tension, rigidity, spring_length,
                                    damping, growth_rate, spring_length_threshold, growth_base_rate]
                                    RodMechanics (bacterial_rods.rs:82-107, 157-240) + the Cycle /
                                    NeighborSensing fields and GROWTH_BASE_RATE of the example agent
                                    (cellular_raza-examples/bacterial_rods/src/main.rs:12-25, 48-109).
                                    Needs interaction_kind CRB_INT_MORSE (RodInteraction<MorsePotential>),
                                    dim 3 and 2 <= rod_n_vertices <= 8; the step then runs aspects
                                    [Mechanics, Interaction, Cycle, NeighborSensing] (main.rs:202-210).
                                    spring_length and growth_rate are per-agent state, the other
                                    entries must be equal for all agents.  Slabs (world_size > 1): the six-phase
                                    protocol crb_slab_halo_export .. crb_slab_finish with crb_slab_compute (no
                                    interior / boundary split, no crb_slab_step); the uniform rows must be declared
                                    with crb_set_uniform_params on every rank. */
} crb_mechanics_kind;

/* Interaction kinds and their per-agent parameter rows (inter_params[i*n_inter + k]).
 * reference: cellular_raza-building-blocks/src/cell_building_blocks/interaction.rs,
 *            cellular_raza-examples/cell_sorting/src/main.rs:43-97 */
typedef enum crb_interaction_kind {
    CRB_INT_CELL_SORTING = 0,       /* [cell_radius, potential_strength, relative_interaction_range]
                                       + species byte; sign convention of the example            */
    CRB_INT_CELL_SORTING_BENCH = 1, /* same potential, sign convention of
                                       cellular_raza-benchmarks/src/cell_sorting.rs:79-86        */
    CRB_INT_MORSE = 2,              /* [radius, potential_stiffness, cutoff, strength]  :151-260 */
    CRB_INT_BOUND_LJ = 3,           /* [epsilon, sigma, bound, cutoff]                  :122-138 */
    CRB_INT_MIE = 4,                /* [radius, strength, bound, cutoff, en, em]        :345-389 */
    CRB_INT_NONE = 5                /* aspects without Interaction (e.g. [Mechanics, DomainForce]): no
                                       pair forces, no parameters                                 */
} crb_interaction_kind;

/* DomainForce aspect: SubDomainBox::calculate_custom_domain_force (chili/update_mechanics.rs:248-274)
 * adds SubDomainForce::calculate_custom_force(pos, vel, info) of the subdomain to every agent's force in
 * step 1 (run_sim.rs:276-278).  That function is user code in the reference; the closed device set
 * offers the linear spring of the reference's own known-answer test
 * (cellular_raza/tests/domain_force_spring.rs:56-68):
 *     force[c] = -domain_force_strength[c] * (pos[c] - domain_force_anchor[c]). */
typedef enum crb_domain_force_kind {
    CRB_DOMAIN_FORCE_NONE = 0,
    CRB_DOMAIN_FORCE_SPRING = 1,
    /* CRB_MECH_ROD only: SubDomainForce of CartesianSubDomainRods (bacterial_rods.rs:433-477): every vertex
     * feels (0, 0, -gel_pressure); a moving vertex within surface_friction_distance of any wall of the
     * DOMAIN additionally feels -vel/|vel| * gel_pressure * surface_friction.
     * domain_force_strength = {gel_pressure, surface_friction, surface_friction_distance}
     * (CartesianCuboidRods::create_subdomains passes surface_friction as the distance, :406 — the host
     * decides which value to send); domain_force_anchor is unused. */
    CRB_DOMAIN_FORCE_ROD_GEL = 2
} crb_domain_force_kind;

/* How the two half-forces of a pair are weighted (SURVEY.md finding 2).
 * CHILI reproduces cellular_raza-core/src/backend/chili/update_mechanics.rs:206-234 as written:
 * the return force of a neighbour voxel inside the same subdomain is dropped, so a pair in
 * different voxels of one subdomain acts with weight 1/2; `ref_n_subdomains` is the `n_threads`
 * of the chili run that is being matched.  PHYSICAL applies both halves everywhere. */
typedef enum crb_pair_weight_mode {
    CRB_WEIGHT_CHILI = 0,
    CRB_WEIGHT_PHYSICAL = 1
} crb_pair_weight_mode;

/* Source of the standard-normal numbers consumed by stochastic mechanics. */
typedef enum crb_noise_mode {
    CRB_NOISE_NONE = 0,   /* random contribution is zero (deterministic mechanics)              */
    CRB_NOISE_TABLE = 1,  /* z[step][agent id][d] uploaded with crb_set_noise (parity runs)     */
    CRB_NOISE_PHILOX = 2  /* counter-based generator on the device keyed by (seed, id, step)    */
} crb_noise_mode;

#define CRB_MAX_DIM 3
#define CRB_MAX_MECH_PARAMS 8
#define CRB_MAX_INTER_PARAMS 6
#define CRB_MAX_ROD_VERTICES 8

/* ---- configuration ------------------------------------------------------------------------ */
typedef struct crb_config {
    uint32_t abi_version;      /* CRB_ABI_VERSION                                                */
    uint32_t dim;              /* 1, 2 or 3 (1-D runs on the per-agent kernel; rods need 3)     */
    uint32_t mechanics_kind;   /* crb_mechanics_kind                                             */
    uint32_t interaction_kind; /* crb_interaction_kind                                           */
    uint32_t solver_order;     /* `mechanics_solver_order` kwarg of run_simulation!: N in
                                  MechanicsAdamsBashforthSolver<N>; 2 (default) | 1 | 0
                                  (run_sim.rs:131-159, solvers.rs:80-139)                        */
    uint32_t pair_weight_mode; /* crb_pair_weight_mode                                           */
    uint32_t ref_n_subdomains; /* chili n_threads to emulate in CRB_WEIGHT_CHILI (>= 1)          */
    uint32_t noise_mode;       /* crb_noise_mode                                                 */
    uint32_t rod_n_vertices;   /* CRB_MECH_ROD only: rows of the position matrix (<= 8)          */
    uint32_t device;           /* CUDA device ordinal                                            */
    /* slab decomposition along axis 0.  Single GPU: rank 0 of 1.                */
    uint32_t rank;
    uint32_t world_size;
    uint64_t n_voxels[CRB_MAX_DIM]; /* CartesianCuboid::n_voxels (cartesian_cuboid_n.rs:61-70)   */
    double domain_min[CRB_MAX_DIM]; /* CartesianCuboid::min                                      */
    double domain_max[CRB_MAX_DIM]; /* CartesianCuboid::max                                      */
    double dx[CRB_MAX_DIM];         /* CartesianCuboid::dx, computed by the host exactly as
                                       cartesian_cuboid_n.rs:183-222 / 226-256 do               */
    double dt;                      /* FixedStepsize::dt                                         */
    uint64_t rng_seed;              /* CartesianCuboid::rng_seed                                 */
    uint64_t capacity;              /* max number of OWNED agents resident on this device        */
    uint64_t ghost_capacity;        /* slabs: max agents of one boundary plane (0 = 1.5 x mean)  */
    uint64_t migration_capacity;    /* slabs: max agents crossing one slab face per step (0=auto)*/
    uint32_t domain_force_kind;     /* crb_domain_force_kind (aspect DomainForce)                */
    uint32_t reserved0;
    double domain_force_strength[CRB_MAX_DIM];
    double domain_force_anchor[CRB_MAX_DIM];
} crb_config;

typedef struct crb_sim crb_sim; /* opaque */

/* ---- lifetime ----------------------------------------------------------------------------- */
/* replaces construct_simulation_runner (chili/datastructures.rs:51-217): builds the voxel
 * grid, plain indices and the emulated chili subdomain map on the device. */
int crb_create(const crb_config* cfg, crb_sim** out);
/* replaces Drop for the runner (cara/mod.rs:16-22). */
void crb_destroy(crb_sim* sim);
/* message of the last failure on `sim` (or of the last failed crb_create when sim == NULL). */
const char* crb_last_error(const crb_sim* sim);
/* ABI version of the loaded library. */
uint32_t crb_abi_version(void);
/* number of per-agent parameters for a mechanics / interaction kind (0 if unknown). */
uint32_t crb_n_mech_params(uint32_t mechanics_kind);
uint32_t crb_n_inter_params(uint32_t interaction_kind);

/* ---- agents ------------------------------------------------------------------------------- */
/* replaces SubDomainBox::insert_initial_cells (chili/datastructures.rs:278-299): copies n agents
 * to the device, bins them with CartesianCuboid::get_voxel_index_of_raw semantics
 * (cartesian_cuboid_n.rs:320-337; pos <= min or index >= n_voxels is CRB_ERR_BOUNDARY /
 * CRB_ERR_INDEX) and sorts them by voxel keeping input order inside a voxel.  AuxStorage starts
 * empty.  pos/vel: n*dim*V doubles (V = rod_n_vertices for rods, else 1; rods are stored
 * vertex-major per agent: pos[(i*V + v)*dim + d]).  species may be NULL (all 0), ids may be
 * NULL (0..n-1). */
int crb_set_agents(crb_sim* sim, uint64_t n, const double* pos, const double* vel,
                   const double* mech_params, const double* inter_params,
                   const uint8_t* species, const uint64_t* ids);

/* Asynchronous form of crb_set_agents for a pipelined host (upload of population k + 1, step of population k and
 * download of population k - 1 overlap when two handles alternate): the copies and the ingestion kernels are only
 * queued on the handle's stream.  The host buffers must be page-locked (crb_host_alloc) and stay untouched until the
 * next synchronising call on the handle (crb_sync, crb_step, crb_get_*, crb_snapshot_wait).  Ingestion errors
 * (BoundaryError / IndexError of an agent outside the domain) are latched on the device and returned by crb_sync /
 * crb_step.  crb_set_aux cannot follow this form.
 * Rods (CRB_MECH_ROD, single-device handles): the parameter rows are uploaded as given and checked ON THE DEVICE (every
 * RodMechanics / MorsePotential field except spring_length and growth_rate must equal row 0); a row that differs is
 * latched as CRB_ERR_INVALID_ARG and returned by the next crb_step (the uploaded population is not usable after that
 * error: upload again).  From then on the handle's clears, its division
 * count and its error latch travel as kernels / stores into page-locked words, never as copy-engine operations that
 * would queue behind the transfers of the handle it alternates with. */
int crb_set_agents_async(crb_sim* sim, uint64_t n, const double* pos, const double* vel,
                         const double* mech_params, const double* inter_params,
                         const uint8_t* species, const uint64_t* ids);

/* declares that every agent shares one parameter row (the common case for the building-block
 * examples); afterwards crb_set_agents may be called with mech_params == inter_params == NULL
 * and skips the per-agent parameter columns entirely.  Passing NULL, NULL reverts. */
int crb_set_uniform_params(crb_sim* sim, const double* mech_row, const double* inter_row);

/* number of agents currently resident. */
uint64_t crb_n_agents(const crb_sim* sim);

/* replaces the read side of save_cells (chili/datastructures.rs:373-396): copies the agents in
 * device order (ascending voxel, chili in-voxel order).  Any output pointer may be NULL.
 * `voxel` receives the VoxelPlainIndex of each agent.  cap = number of agents the buffers hold. */
int crb_get_agents(crb_sim* sim, uint64_t cap, double* pos, double* vel, uint64_t* ids,
                   uint64_t* voxel);
/* mutable model state that is not pos/vel (rod spring_length, growth_rate …). */
int crb_get_params(crb_sim* sim, uint64_t cap, double* mech_params, double* inter_params,
                   uint8_t* species);

/* AuxStorageMechanics (chili/aux_storage.rs:41-130) in device order: the ring buffers of
 * previous increments oldest-first (hist_dx/hist_dv: cap*2*dim*V doubles, hist_len: entries
 * used, 0..2) and current_force (zero between steps). */
int crb_get_aux(crb_sim* sim, uint64_t cap, double* hist_dx, double* hist_dv, uint32_t* hist_len);
/* restores AuxStorage for agents given in the order of the preceding crb_set_agents call. */
int crb_set_aux(crb_sim* sim, uint64_t n, const double* hist_dx, const double* hist_dv,
                const uint32_t* hist_len);

/* z[(step*n_ids + id)*dim*V + k] standard normal numbers for the next n_steps steps, indexed by
 * agent id (ids must be < n_ids).  CRB_NOISE_TABLE only. */
int crb_set_noise(crb_sim* sim, const double* z, uint64_t n_steps, uint64_t n_ids);

/* CellIdentifier and parent of every agent in device order (cellular_raza-concepts/src/cell.rs:18-25,
 * 128-137): 3 words per agent {kind, a, b} with kind 0 = Division(voxel a, counter b), 1 = Initial(a),
 * 2 = Inserted(a, b); parent kind = UINT64_MAX when the agent has no parent.  Point agents are always
 * Initial(id).  Either pointer may be NULL. */
int crb_get_identifiers(crb_sim* sim, uint64_t cap, uint64_t* ident, uint64_t* parent);
/* rods: number of Division events carried out by the last step (update_cell_cycle_4,
 * chili/update_cycle.rs:9-69). */
uint64_t crb_n_divisions(const crb_sim* sim);

/* ---- stepping ----------------------------------------------------------------------------- */
/* replaces n iterations of the generated loop body (run_sim.rs:380-403) for
 * aspects [Mechanics, Interaction]: update_mechanics_interaction_step_1/2/3, local_mechanics_update,
 * apply_boundary, sort_cells_in_voxels_step_1/2 (chili/update_mechanics.rs). Asynchronous errors
 * raised on the device (CalcError, BoundaryError) are latched and returned here. */
int crb_step(crb_sim* sim, uint64_t n_steps);
/* queues the steps and returns without waiting for the device; errors raised on the device stay latched until
 * crb_sync / crb_step (rods always synchronise: their division counts are read back every step). */
int crb_step_async(crb_sim* sim, uint64_t n_steps);
/* blocks until all queued device work of the handle is done and returns the latched status. */
int crb_sync(crb_sim* sim);
/* number of steps taken since creation. */
uint64_t crb_iteration(const crb_sim* sim);
/* sets the step counter (the iteration number of the generated main loop, run_sim.rs:380-403): the device generator
 * of CRB_NOISE_PHILOX is keyed by (rng_seed, agent id, iteration), so a host that restarts a run from a stored
 * population — or repeats one from its initial state — puts the counter back where that population belongs.
 * crb_set_agents does NOT touch the counter (a host that re-uploads the population every step must keep drawing
 * fresh numbers). */
int crb_set_iteration(crb_sim* sim, uint64_t iteration);

/* ---- slab decomposition over several devices --------------------------------------
 * Replaces the Communicator / ChannelComm traffic of chili between subdomains
 * (cellular_raza-core/src/backend/chili/simulation_flow.rs:438-450, 689-707; message types
 * PosInformation / SendCell, chili/update_mechanics.rs:12-61) and its barrier
 * (simulation_flow.rs:350-423).  The domain is cut into world_size slabs of whole voxel planes along
 * axis 0; every rank owns one handle created with (rank, world_size) and keeps one ghost plane per
 * neighbour.  No ForceInformation message exists: each side evaluates its own half of a
 * cross-slab pair.  The caller moves the fixed-size device buffers between neighbouring devices
 * (ncclSend/ncclRecv, cudaMemcpyPeer, or a plain device copy when two handles share a device):
 *
 *     crb_slab_halo_export(lo_out, hi_out)            my boundary planes            -> neighbours
 *     crb_slab_halo_import(lo_in, hi_in)              neighbours' boundary planes   -> ghost planes
 *     crb_slab_compute(save_prev)                     forces + integrate + new voxel keys
 *     crb_slab_migration_export(lo_out, hi_out, n)    agents that left the slab     -> neighbours
 *     crb_slab_migration_import(lo_in, hi_in, n)      arrivals
 *     crb_slab_finish(save_prev)                      cell-list rebuild (sort_cells_in_voxels)
 *
 * Overlap of the halo exchange with interior force work: crb_slab_compute may be
 * replaced by the pair
 *
 *     crb_slab_compute_interior(save_prev)   planes whose 3^D neighbourhood holds no ghost plane; needs
 *                                            no halo, runs on the handle's second (low-priority) stream
 *                                            and may be called BEFORE crb_slab_halo_import
 *     crb_slab_compute_boundary()            the one or two owned planes next to a ghost plane, on the
 *                                            main stream after crb_slab_halo_import; joins both streams
 *
 * so that the sequence  compute_interior, halo_export, <send/recv>, halo_import, compute_boundary  hides
 * the transfer behind the interior pass.  Results are bit-identical to crb_slab_compute (every target
 * sums its candidates in the same order whichever launch evaluates it).
 *
 * "lo" is the neighbour with rank - 1, "hi" rank + 1; a buffer pointer may be NULL where there is no
 * neighbour.  All pointers are DEVICE pointers of at least crb_slab_halo_bytes / _migration_bytes.
 * save_prev != 0 keeps the older ring-buffer entry readable by crb_get_aux after this step. */
uint64_t crb_slab_halo_bytes(const crb_sim* sim);
uint64_t crb_slab_migration_bytes(const crb_sim* sim);
/* out[0] first owned global plane, [1] owned planes, [2] has lower neighbour, [3] has upper neighbour,
 * [4] ghost capacity, [5] migration capacity, [6] owned capacity, [7] voxels per plane. */
int crb_slab_info(const crb_sim* sim, uint64_t out[8]);
int crb_slab_halo_export(crb_sim* sim, void* lo_buf, void* hi_buf);
int crb_slab_halo_import(crb_sim* sim, const void* lo_buf, const void* hi_buf);
int crb_slab_compute(crb_sim* sim, int save_prev);
int crb_slab_compute_interior(crb_sim* sim, int save_prev);
int crb_slab_compute_boundary(crb_sim* sim);
/* n_out == NULL: the export is only enqueued; its counts are read back by crb_slab_migration_import in
 * the same host synchronisation as the arrival counts and are available from crb_slab_migration_counts
 * afterwards (the buffers have a fixed size, so the transfer does not need the counts). */
int crb_slab_migration_export(crb_sim* sim, void* lo_buf, void* hi_buf, uint64_t n_out[2]);
int crb_slab_migration_counts(const crb_sim* sim, uint64_t n_out[2]);
int crb_slab_migration_import(crb_sim* sim, const void* lo_buf, const void* hi_buf, uint64_t n_in[2]);
int crb_slab_finish(crb_sim* sim, int save_prev);

/* ---- the same step behind ONE call: exchange over peer memory, no host synchronisation ------------------------
 * Every slab owns an "exchange arena" in device memory (import buffers for both faces, double-buffered by step
 * parity, plus four step counters).  After the neighbours' arenas have been mapped (same process: the raw device
 * pointer of crb_slab_arena, peer access is enabled when the devices differ; other processes of the node: the
 * 64-byte CUDA IPC handle of crb_slab_ipc_handle, exchanged by any means — MPI, a file, torch.distributed)
 *
 *     crb_slab_step(sim, n)
 *
 * queues n complete steps: the halo and migration export kernels store straight into the neighbour's import
 * buffers over NVLink / NVSwitch and then raise a counter in the NEIGHBOUR's arena; the receiver's stream holds a
 * one-thread kernel that waits on its own, local counter.  Agent counts stay on the device (the grids are sized by
 * the capacities), the interior planes overlap the halo exchange as in crb_slab_compute_interior, and the host
 * synchronises ONCE, at the end of the batch (latched device errors are returned there).  Results are bit-identical
 * to the six-phase protocol above and to a single handle.  Every rank of the decomposition must call crb_slab_step
 * with the same n; a neighbour that never arrives turns into CRB_ERR_STATE after 20 s instead of a hang.
 * A host that drives several slabs from ONE thread must interleave the ranks step by step (n = 1, round robin):
 * a rank's queue would otherwise wait for exports that have not been queued yet. */
uint64_t crb_slab_arena_bytes(const crb_sim* sim);
void* crb_slab_arena(crb_sim* sim);
int crb_slab_ipc_handle(crb_sim* sim, void* handle64);
int crb_slab_connect(crb_sim* sim, void* lo_arena, void* hi_arena);
int crb_slab_connect_ipc(crb_sim* sim, const void* lo_handle64, const void* hi_handle64);
int crb_slab_step(crb_sim* sim, uint64_t n_steps);
/* the same without the final synchronisation (for a host thread that drives several slabs: queue one step on every
 * rank, then crb_sync every rank); the agent count and the getters are valid again after crb_sync. */
int crb_slab_step_async(crb_sim* sim, uint64_t n_steps);

/* ---- storage snapshots without stopping the stepping loop ---------------------------------------
 * Replaces the blocking save_cells of chili at PartialSave iterations (chili/datastructures.rs:373-396,
 * run_sim.rs:432-444).  crb_snapshot_async packs the CURRENT state (same layouts and order as
 * crb_get_agents / crb_get_aux; any pointer may be NULL) into a device-side snapshot buffer on the
 * handle's stream and starts the device-to-host copies on a side stream; it returns at once and
 * crb_step may be called while the copies run.  The host buffers (pinned memory from crb_host_alloc
 * for a truly asynchronous copy) belong to the library until crb_snapshot_wait returns; one snapshot
 * is in flight at a time (a second call waits for the first).
 * Rods: pos / vel are the V x 3 rows of crb_get_agents, ids the number of the identifier (crb_get_identifiers has the
 * triples), voxel as for point agents; the ring buffers are not part of a rod snapshot (hist_* must be NULL:
 * CRB_ERR_INVALID_ARG otherwise; crb_get_aux reads them). */
int crb_snapshot_async(crb_sim* sim, uint64_t cap, double* pos, double* vel, uint64_t* ids,
                       uint64_t* voxel, double* hist_dx, double* hist_dv, uint32_t* hist_len);
int crb_snapshot_wait(crb_sim* sim, uint64_t* n_agents);
/* page-locked host memory for snapshot / upload buffers. */
void* crb_host_alloc(uint64_t bytes);
void crb_host_free(void* p);

/* ---- parity taps (used by tests; not needed by a production host) ------------------------- */
/* forces that the next step would integrate (AuxStorageMechanics::current_force after
 * update_mechanics_interaction_step_3), device order, cap*dim*V doubles; does not advance. */
int crb_debug_forces(crb_sim* sim, uint64_t cap, double* force);
/* rods: the NeighborSensing accumulator of every agent as the next step would see it in
 * react_to_neighbors (examples/bacterial_rods/src/main.rs:48-78), device order; does not advance. */
int crb_debug_neighbor_counts(crb_sim* sim, uint64_t cap, uint64_t* counts);
/* migration set of the last step: (id, old VoxelPlainIndex, new VoxelPlainIndex) triples of
 * every agent that changed voxel in sort_cells_in_voxels.  Returns the count in *n_out. */
int crb_get_migrations(crb_sim* sim, uint64_t cap, uint64_t* ids, uint64_t* old_voxel,
                       uint64_t* new_voxel, uint64_t* n_out);
/* number of agents in every voxel (n_voxels_total entries, plain index order). */
int crb_get_voxel_counts(crb_sim* sim, uint64_t cap, uint32_t* counts);
/* ordered candidate pairs evaluated by one force pass for the current state. */
int crb_count_candidate_pairs(crb_sim* sim, uint64_t* n_pairs);
/* number of kernels launched by this handle so far (bench.py `gpu_launches`). */
uint64_t crb_kernel_launches(const crb_sim* sim);
/* name of the force kernel the planner chose for the current population (bench / profile labels). */
const char* crb_force_kernel_name(const crb_sim* sim);
/* device time in milliseconds spent in a named stage since the last crb_reset_timers
 * (CUDA events on the handle's stream; only recorded after crb_enable_timers(sim, 1)).
 * stage: 0 force+integrate, 1 scan, 2 reorder, 3 halo, 4 migration. */
int crb_enable_timers(crb_sim* sim, int on);
int crb_reset_timers(crb_sim* sim);
int crb_stage_time_ms(crb_sim* sim, uint32_t stage, double* ms, uint64_t* launches);

/* the cudaStream_t all work of this handle is queued on (for CUDA-event timing by the caller). */
void* crb_stream(crb_sim* sim);
/* DFMA-only micro-kernel on `device`: measured FP64 FMA throughput in TFLOP/s (2 flop per FMA);
 * the roofline denominator of the force kernel (MEASURED_PEAKS.json holds no FP64 figure). */
int crb_measure_fp64_peak(uint32_t device, double* tflops, double* ms);

/* parity tap of the device reciprocal square root used by every potential (y[i] = 1/sqrt(x[i])). */
int crb_debug_rsqrt(uint32_t device, const double* x, double* y, uint64_t n);
/* parity tap of the bounded-argument exponential of the Morse potential (y[i] = exp(x[i]), |x| < 700). */
int crb_debug_exp(uint32_t device, const double* x, double* y, uint64_t n);

/* ---- helpers that restate reference setup arithmetic on the host -------------------------- */
/* get_decomp_res (cartesian_cuboid_n.rs:19-59): returns 0 and (n, m, average_len) or
 * CRB_ERR_DECOMPOSE. */
int crb_decomp_res(uint64_t n_voxel, uint64_t n_regions, uint64_t* n, uint64_t* m,
                   uint64_t* average_len);
/* CartesianCuboid::from_boundaries_and_interaction_range (cartesian_cuboid_n.rs:183-222):
 * fills n_voxels and dx for `dim` axes. */
int crb_domain_from_interaction_range(uint32_t dim, const double* min, const double* max,
                                      double interaction_range, uint64_t* n_voxels, double* dx);
/* CartesianCuboid::from_boundaries_and_n_voxels (cartesian_cuboid_n.rs:226-256). */
int crb_domain_from_n_voxels(uint32_t dim, const double* min, const double* max,
                             const uint64_t* n_voxels, double* dx);

#ifdef __cplusplus
}
#endif
#endif /* CRB_H */
