// crb_api.cu — host side of the C ABI declared in include/crb.h.
//
// Owns all device memory behind the opaque handle, sequences the kernels of one step on the
// handle's stream and latches device-side errors (SimulationError variants of
// cellular_raza-core/src/backend/chili/errors.rs:73-121).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "crb_force_tile.cuh"
#include "crb_force_tile2.cuh"
#include "crb_force_tile3.cuh"
#include "crb_launch.cuh"
#include "crb_rods.cuh"

using namespace crb;

namespace {

thread_local std::string g_create_error;
const char* const kAsyncPendingMsg =
    "asynchronous slab steps are queued (crb_slab_step_async): call crb_sync first";
const char* const kOpenStepMsg =
    "a slab step is open (crb_slab_compute .. crb_slab_finish): the state is partly updated, read it after crb_slab_finish";

struct StageTimer {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    double ms = 0;
    u64 launches = 0;
};
constexpr int kNumStages = 5;

}  // namespace

struct crb_sim {
    crb_config cfg;
    Grid g;
    int D = 3, V = 1;
    int n_mp = 0, n_ip = 0;
    bool uniform = true;
    bool uniform_declared = false;
    Uniform uni;
    u64 cap = 0, n = 0;  // owned-agent capacity / owned agents
    u64 G = 0;           // ghost capacity per side = index of the first owned agent in every column
    u64 cap_total = 0;   // cap + 2 G
    u64 mcap = 0;        // migration capacity per side and step
    u64 nvox = 0;
    // slab decomposition
    int rank = 0, world = 1;
    bool has_lo = false, has_hi = false;  // neighbours below / above along axis 0
    HaloLayout hl;
    MigLayout ml;
    RodHaloLayout rhl;  // CRB_MECH_ROD slabs: hl.bytes / ml.bytes hold the sizes of these layouts
    RodMigLayout rml;
    bool step_open = false;  // crb_slab_compute done, crb_slab_finish pending
    bool mig_counts_pending = false;  // export enqueued without reading its counts back (n_out == NULL)
    u64 mig_out[2] = {0, 0};          // agents exported in the open step (lo, hi)
    u64 n_src = 0;           // unsorted sources of the open step (owned + arrivals)
    u64 n_gone = 0;          // agents exported in the open step
    // crb_slab_step: exchange arena (import buffers + flags) that the neighbours write into, their arenas mapped
    // into this process, device-side agent counts and the exchange step counter
    unsigned char* arena = nullptr;
    u64 arena_bytes = 0, arena_halo = 0, arena_mig = 0;  // aligned sizes of one halo / migration buffer
    unsigned char* peer_arena[2] = {nullptr, nullptr};
    bool peer_ipc[2] = {false, false};
    u64* cnt = nullptr;
    u64* cnt_host = nullptr;  // pinned staging of the owned count
    u64 slab_steps = 0;
    bool slab_async_pending = false;  // crb_slab_step_async queued: s->n is stale until crb_sync
    bool slab_preloaded = false;      // the kernels of crb_slab_step are loaded (lazy loading waits for an idle device)
    cudaStream_t stream = nullptr;
    // overlap of the halo exchange with interior force work: the interior planes run on `stream2`
    // (lowest priority) while the halo travels on `stream` (highest priority)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_interior = nullptr;
    bool kernels_only = false;  // the handle is driven through the asynchronous API: no copy-engine memsets / D2D copies
    int interior_state = 0;  // 0 none, 1 crb_slab_compute_interior done (boundary pending)
    // storage snapshots: packed on `stream` into `snap`, copied to the caller's host buffers on `stream_io`
    cudaStream_t stream_io = nullptr;
    cudaEvent_t ev_snap_packed = nullptr, ev_snap_done = nullptr;
    double* snap = nullptr;
    u64 snap_cap = 0;
    bool snap_pending = false;
    u64 snap_n = 0;
    u8* snap_hlen_host = nullptr;  // pinned
    double *snap_hx_host = nullptr, *snap_hv_host = nullptr;
    uint32_t* snap_hl_out = nullptr;
    State X, T, Y;  // X: sorted state read by the force pass, T: integrated, Y: reorder target
    u32 *key = nullptr, *key_alt = nullptr;
    u32 *start = nullptr, *start_alt = nullptr;
    u32* count_new = nullptr;
    u32 *newkey = nullptr, *rank_arb = nullptr, *tmp_idx = nullptr, *slot_of_src = nullptr;
    u64* sortkey = nullptr;
    u32* tile_sums = nullptr;
    u32 n_tiles = 0;
    double* stage = nullptr;  // cap * 8 doubles staging for agent-major host rows
    double* rod_up = nullptr;  // rods, crb_set_agents_async: pos, vel, parameter rows and ids of one upload side by side
    double* force_out = nullptr;
    double* noise = nullptr;
    u64 noise_steps = 0, noise_ids = 0, noise_cursor = 0;
    int* err = nullptr;       // device latch
    int* err_host = nullptr;  // pinned
    u64* scratch_u64 = nullptr;  // small device scratch (counters)
    u64 iteration = 0;
    u64 launches = 0;
    bool have_agents = false;
    bool aux_settable = false;
    u64 last_step_n = 0;  // agents that took part in the last step (migration tap)
    // k_force_tile geometry (chosen in crb_set_agents from the mean density)
    TileLaunch tile;
    bool tile_ok = false;
    int tile_threads = kTileThreads;
    size_t tile_smem = 0;
    // k_force_tile2 geometry (2 x 2-row tiles; preferred when it fits)
    TileLaunch tile2;
    bool tile2_ok = false;
    size_t tile2_smem = 0;
    int tile2_threads = kTileThreads;
    // k_force_tile3 geometry (2 x 2-row tiles, tensor-core distance filter)
    TileLaunch tile3;
    bool tile3_ok = false;
    size_t tile3_smem = 0;
    int tile3_threads = kTileThreads;
    int kernel_choice = 0;  // 0 auto, 1 per-agent kernel, 2 k_force_tile, 3 k_force_tile2, 4 k_force_tile3 (env CRB_FORCE_KERNEL)
    bool timers_on = false;
    StageTimer timers[kNumStages];
    // CUDA graphs of TWO consecutive steps (the X/Y, key, start buffers swap every step, so the pointer set
    // repeats with period 2), one per buffer orientation; see crb_step
    cudaGraphExec_t step_graph[2] = {nullptr, nullptr};
    const void* step_graph_key[2] = {nullptr, nullptr};
    u64 step_graph_launches = 0;  // kernels inside one graph (two steps)
    bool graphs_off = false;      // env CRB_NO_GRAPH
    // bacterial rods (CRB_MECH_ROD): own state layout, see crb_rods.cuh
    bool rods = false;
    RodState RX, RT, RY;
    u8* div_flag = nullptr;
    u32* id_counter = nullptr;
    u32* n_div_dev = nullptr;
    u32* div_list = nullptr;  // rods: slots with a pending Division event in this step
    u64* nb_out = nullptr;
    u64 last_divisions = 0;
    std::vector<void*> allocs;
    std::string error;
};

namespace {

int fail(crb_sim* s, int code, const std::string& msg) {
    if (s)
        s->error = msg;
    else
        g_create_error = msg;
    return code;
}

#define CU(s, call)                                                                        \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(s, CRB_ERR_CUDA,                                                   \
                        std::string(#call) + ": " + cudaGetErrorString(e__));              \
    } while (0)

template <class T>
int dalloc(crb_sim* s, T** p, u64 count) {
    void* q = nullptr;
    CU(s, cudaMalloc(&q, std::max<u64>(count, 1) * sizeof(T)));
    s->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

int alloc_state(crb_sim* s, State& st, bool dynamic, bool stat) {
    std::memset(&st, 0, sizeof(st));
    const u64 c = s->cap_total * s->V;
    int rc;
    if (dynamic) {
        for (int d = 0; d < s->D; ++d) {
            if ((rc = dalloc(s, &st.pos[d], c))) return rc;
            if ((rc = dalloc(s, &st.vel[d], c))) return rc;
            if ((rc = dalloc(s, &st.h1x[d], c))) return rc;
            if ((rc = dalloc(s, &st.h1v[d], c))) return rc;
            if ((rc = dalloc(s, &st.h0x[d], c))) return rc;
            if ((rc = dalloc(s, &st.h0v[d], c))) return rc;
        }
        if ((rc = dalloc(s, &st.hlen, s->cap_total))) return rc;
        if ((rc = dalloc(s, &st.posf, s->cap_total))) return rc;
        if ((rc = dalloc(s, &st.meta, s->cap_total))) return rc;
    }
    if (stat) {
        if ((rc = dalloc(s, &st.id, s->cap_total))) return rc;
        if ((rc = dalloc(s, &st.species, s->cap_total))) return rc;
        for (int k = 0; k < s->n_mp; ++k)
            if ((rc = dalloc(s, &st.mp[k], s->cap_total))) return rc;
        for (int k = 0; k < s->n_ip; ++k)
            if ((rc = dalloc(s, &st.ip[k], s->cap_total))) return rc;
    }
    return 0;
}

// parameter mode word of the exchange headers: 0 = per-agent parameter columns, otherwise an FNV-1a hash of the
// uniform rows with the lowest bit set
u64 param_mode(const crb_sim* s) {
    if (!s->uniform) return 0;
    u64 h = 1469598103934665603ull;
    auto mix = [&](const double* p, int n) {
        for (int k = 0; k < n; ++k) {
            u64 b;
            std::memcpy(&b, &p[k], 8);
            for (int i = 0; i < 8; ++i) {
                h ^= (b >> (8 * i)) & 0xff;
                h *= 1099511628211ull;
            }
        }
    };
    mix(s->uni.mp, s->n_mp);
    mix(s->uni.ip, s->n_ip);
    return h | 1ull;
}

inline u32 blocks_for(u64 n, u32 threads) { return (u32)((n + threads - 1) / threads); }

// NVTX ranges per stage (the reference's counterpart: `tracing` spans around every step function,
// chili/update_mechanics.rs:152, 284, 343, 386, 417, 484), behind the environment flag CRB_NVTX=1: host-side ranges
// around the launches of a stage, visible in Nsight Systems next to the kernels they enqueue
const char* const kStageNames[] = {"crb:force+integrate (update_mechanics_interaction_step_1-3, local_mechanics_update, apply_boundary)",
                                   "crb:scan (sort_cells_in_voxels: histogram prefix)",
                                   "crb:reorder (sort_cells_in_voxels_step_1/2)", "crb:halo (PosInformation exchange)",
                                   "crb:migration (SendCell exchange)"};
bool nvtx_on() {
    static const bool on = std::getenv("CRB_NVTX") != nullptr;
    return on;
}

void stage_begin(crb_sim* s, int stage, cudaStream_t st = nullptr) {
    if (nvtx_on()) nvtxRangePushA(kStageNames[stage]);
    if (!s->timers_on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st ? st : s->stream);
    s->timers[stage].ev.emplace_back(a, b);
}
void stage_end(crb_sim* s, int stage, int n_launches, cudaStream_t st = nullptr) {
    s->launches += n_launches;
    if (nvtx_on()) nvtxRangePop();
    if (!s->timers_on) return;
    cudaEventRecord(s->timers[stage].ev.back().second, st ? st : s->stream);
    s->timers[stage].launches += n_launches;
}
void timers_resolve(crb_sim* s) {
    for (auto& t : s->timers) {
        for (auto& e : t.ev) {
            float ms = 0;
            cudaEventSynchronize(e.second);
            cudaEventElapsedTime(&ms, e.first, e.second);
            t.ms += ms;
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
        t.ev.clear();
    }
}

KParams make_params(crb_sim* s) {
    KParams P;
    std::memset(&P, 0, sizeof(P));
    P.g = s->g;
    P.A = s->X;
    P.B = s->T;
    P.u = s->uni;
    P.n = s->n;
    P.first = s->G;
    P.cap = s->cap_total;
    P.key = s->key;
    P.start = s->start;
    P.newkey = s->newkey;
    P.rank_arb = s->rank_arb;
    P.sortkey = s->sortkey;
    P.count_new = s->count_new;
    P.force_out = s->force_out;
    P.noise = s->noise;
    P.noise_ids = s->noise_ids;
    P.noise_cursor = s->noise_cursor;
    P.iteration = s->iteration;
    P.seed = s->cfg.rng_seed;
    P.err = s->err;
    P.dt = s->cfg.dt;
    P.mech = (int)s->cfg.mechanics_kind;
    P.inter = (int)s->cfg.interaction_kind;
    P.order = (int)s->cfg.solver_order;
    P.weight_mode = (int)s->cfg.pair_weight_mode;
    P.noise_mode = (int)s->cfg.noise_mode;
    P.save_prev = 0;
    P.dforce = (int)s->cfg.domain_force_kind;
    for (int d = 0; d < 3; ++d) {
        P.dk[d] = d < s->D ? s->cfg.domain_force_strength[d] : 0.0;
        P.dx0[d] = d < s->D ? s->cfg.domain_force_anchor[d] : 0.0;
    }
    return P;
}

template <int D, int INTER, bool UNIFORM>
void launch_force_t(crb_sim* s, const KParams& P, bool force_only) {
    const u32 grid = blocks_for(P.n, kForceThreads);
    if (!grid) return;
    if (force_only)
        k_force_integrate<D, INTER, UNIFORM, true><<<grid, kForceThreads, 0, s->stream>>>(P, P.first, P.n);
    else
        k_force_integrate<D, INTER, UNIFORM, false><<<grid, kForceThreads, 0, s->stream>>>(P, P.first, P.n);
}

template <int D, int INTER>
void launch_force_u(crb_sim* s, const KParams& P, bool force_only) {
    if (s->uniform)
        launch_force_t<D, INTER, true>(s, P, force_only);
    else
        launch_force_t<D, INTER, false>(s, P, force_only);
}

template <int D>
void launch_force_d(crb_sim* s, const KParams& P, bool force_only) {
    switch (s->cfg.interaction_kind) {
        case CRB_INT_CELL_SORTING: launch_force_u<D, CRB_INT_CELL_SORTING>(s, P, force_only); break;
        case CRB_INT_CELL_SORTING_BENCH:
            launch_force_u<D, CRB_INT_CELL_SORTING_BENCH>(s, P, force_only);
            break;
        case CRB_INT_MORSE: launch_force_u<D, CRB_INT_MORSE>(s, P, force_only); break;
        case CRB_INT_BOUND_LJ: launch_force_u<D, CRB_INT_BOUND_LJ>(s, P, force_only); break;
        case CRB_INT_MIE: launch_force_u<D, CRB_INT_MIE>(s, P, force_only); break;
        case CRB_INT_NONE: launch_force_u<D, CRB_INT_NONE>(s, P, force_only); break;
    }
}

// ---- k_force_tile -------------------------------------------------------------------------------
// interaction cutoff of a uniform parameter row (beyond it every potential of the set is exactly 0)
double uniform_cutoff(const crb_sim* s) {
    const double* ip = s->uni.ip;
    switch (s->cfg.interaction_kind) {
        case CRB_INT_CELL_SORTING:
        case CRB_INT_CELL_SORTING_BENCH: return ip[2] * (ip[0] + ip[0]);  // range * (R1 + R2)
        case CRB_INT_MORSE: return ip[2];
        case CRB_INT_BOUND_LJ: return ip[3];
        case CRB_INT_MIE: return ip[3];
    }
    return INFINITY;
}

// chooses the pencil geometry from the mean density; called whenever the population changes
void plan_tile(crb_sim* s) {
    s->tile_ok = false;
    s->tile2_ok = false;
    s->tile3_ok = false;
    if (!s->uniform || s->n == 0 || s->kernel_choice == 1) return;
    if (s->D < 2) return;  // 1-D domains: 3 neighbour voxels, the per-agent kernel is the path
    if (s->cfg.interaction_kind == CRB_INT_NONE) return;  // no pair forces: the per-agent kernel integrates
    // the tile kernel evaluates the species potential without its upper clamp (acts only for R1 + R2 < 0)
    if ((s->cfg.interaction_kind == CRB_INT_CELL_SORTING || s->cfg.interaction_kind == CRB_INT_CELL_SORTING_BENCH) &&
        s->uni.ip[0] < 0.0)
        return;
    const Grid& g = s->g;
    const int n2 = g.ng[2];
    const int nrows = s->D == 3 ? 9 : 3;
    const double owned_vox = (double)(g.own_hi - g.own_lo) * g.ng[1] * n2;
    const double rho = std::max((double)s->n / std::max(owned_vox, 1.0), 1e-3);
    // 128-thread CTAs; every CTA takes ~118 targets of one row (count-balanced, snapped to voxel
    // boundaries: <= 128 with high probability, so the four warps run one full pass each); candidates
    // <= staging budget of 3 CTAs/SM
    const size_t budget = 74 * 1024;
    TileLaunch& T = s->tile;
    T.chunk_len = std::min(n2, 56);
    T.nchunks = (n2 + T.chunk_len - 1) / T.chunk_len;
    T.kmax = T.chunk_len + 2;
    const size_t fixed = (size_t)(kTileQueue + 4) * kTileThreads * sizeof(unsigned short) +
                         (size_t)kTileRows * (T.kmax + 1) * sizeof(u32) + 2048;
    const u32 cap = (u32)(((budget - fixed) / (sizeof(TileCand) + sizeof(float4))) / 32 * 32);
    // candidates of a segment of t targets: nrows * (t + 2 voxels' worth) (+ sentinels)
    // (10 % slack ~ 3 sigma of the Poisson spread; an overflowing pencil is split into z-windows)
    const double per_target = nrows * 1.1;
    double tseg = (cap - kTileRows * kTileRowPad - nrows * 2.0 * rho * 1.1) / per_target;
    tseg = std::max(8.0, std::min(tseg, 124.0));
    T.tseg = (u32)tseg;
    T.nseg = (int)std::ceil(T.chunk_len * rho * 1.3 / T.tseg) + 1;
    T.x_begin = g.own_lo;
    T.x_count = g.own_hi - g.own_lo;
    T.cap = cap;
    // conservative FP32 filter radius: |d_f32 - d| <= 2^-22.5 * extent per axis (crb_force_tile.cuh)
    double ext = 0;
    for (int d = 0; d < s->D; ++d) ext = std::max(ext, g.mx[d] - g.mn[d]);
    const double cutoff = uniform_cutoff(s);
    float cut2f = INFINITY;
    if (std::isfinite(cutoff) && cutoff > 0) {
        const double c = (cutoff + 2.0 * ext * std::ldexp(1.0, -22)) * (1.0 + std::ldexp(1.0, -19));
        const double c2 = c * c;
        cut2f = (float)c2;
        if ((double)cut2f < c2) cut2f = std::nextafterf(cut2f, INFINITY);
    }
    T.cut2f = cut2f;
    s->tile_threads = kTileThreads;
    s->tile_smem = tile_smem_bytes(T.cap, T.kmax, s->tile_threads);
    s->tile_ok = s->tile_smem <= 227 * 1024;

    // ---- k_force_tile2: 2 x 2 target rows per CTA, 16 staged rows ---------------------------------------
    s->tile2_ok = false;
    if (s->kernel_choice == 2) return;
    const int trows = s->D == 3 ? 4 : 2;   // target rows of a full tile
    const int srows = s->D == 3 ? 16 : 4;  // staged rows that hold agents
    const size_t per_cand = sizeof(TileCand) + sizeof(float4) / 2 + sizeof(float2) / 2;
    const double per_t = (double)srows / trows * 1.05;  // staged candidates per target (+5 % spread)
    const double halo = srows * 2.0 * rho * 1.05 + kT2Rows * (kT2Pad + 1);
    const double col = trows * rho;  // targets per voxel column of a tile
    // CTA shape: 256 threads x 2 CTAs/SM or 128 x 4 (16 warps per SM either way); the one that keeps
    // more targets resident per SM wins.  A CTA wants ~NT - 12 targets (every row of targets is padded to
    // a multiple of 4 slots); its chunk of the fastest axis holds m segments of that size, with 7 %
    // head room for the spread of the counts so that ceil(R / tseg) segments almost always suffice.
    double best = 0, plan_fill = 0;
    const char* forced = std::getenv("CRB_TILE2_THREADS");
    for (int nt : {256, 128}) {
        if (forced && std::atoi(forced) != nt) continue;
        const int ctas = 512 / nt;
        const size_t budget2 = (size_t)227 * 1024 / ctas - 1024;  // 1 KB per CTA is reserved by the driver
        const double lanes = nt - 12.0;
        // chunks of the fastest axis (the start[] table holds chunk_len + 3 entries for each of the 16
        // rows): the split whose ceil(R / tseg) equal segments come closest to the lane budget, with
        // 5 % head room for the spread of the counts
        for (int nchunks = 1; nchunks <= n2; ++nchunks) {
            const int chunk_len = (n2 + nchunks - 1) / nchunks;
            if (chunk_len > 61 && nchunks < n2) continue;
            if (nchunks > 1 && (n2 + nchunks - 2) / (nchunks - 1) == chunk_len) continue;  // same length, more chunks
            const int kmax = chunk_len + 2;
            const size_t fixed2 = (size_t)(kTileQueue + 4) * nt * sizeof(unsigned short) +
                                  (size_t)kT2Rows * (kmax + 1) * sizeof(u32) + 64 + 34 * per_cand + 1536 /* static */;
            if (budget2 <= fixed2) continue;
            const u32 cap2 = (u32)std::min<size_t>(((budget2 - fixed2) / per_cand) / 2 * 2, 60000);
            const double ts = std::min(((double)cap2 - halo) / per_t, lanes);
            if (ts < 8.0) continue;
            const double Rc = chunk_len * col;
            const int mseg = std::max(1, (int)std::ceil(Rc * 1.05 / ts));
            const double fill = Rc / mseg;  // expected targets per CTA
            const double score = fill * ctas;
            if (score <= best * 1.001) continue;
            best = score;
            TileLaunch& U = s->tile2;
            U = T;
            U.chunk_len = chunk_len;
            U.nchunks = (n2 + chunk_len - 1) / chunk_len;
            U.kmax = kmax;
            U.tseg = (u32)ts;
            U.cap = cap2;
            U.nseg = mseg + 1;
            s->tile2_threads = nt;
            s->tile2_smem = tile2_smem_bytes(U.cap, U.kmax, nt);
            s->tile2_ok = s->tile2_smem <= budget2 + 1024;
            plan_fill = fill;
        }
    }
    // populations whose 2 x 2-row tiles cannot even give every SM one CTA (the reference's own examples:
    // cell_sorting as shipped is 6^3 voxels = 9 tiles) run the one-row pencils of k_force_tile instead: four
    // times the CTAs at half the size; same bits (tested).  Measured at 1,600 cells: 28.7 vs 33.3 us per step;
    // equal from 20 k cells on.
    if (s->tile2_ok && s->tile_ok && s->kernel_choice == 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, (int)s->cfg.device);
        const long tiles = (long)((T.x_count + 1) / 2) * ((g.ng[1] + 1) / 2) * s->tile2.nchunks;
        if (tiles < sms) s->tile2_ok = false;
    }
    // ---- k_force_tile3: the same 2 x 2-row tiles with the tensor-core distance filter ---------------------
    // 52 B per staged candidate (exact record 32 B + five FP32 filter arrays), no sentinels, target rows padded
    // to multiples of 16 slots (a half-warp = one MMA tile of targets lies in one row)
    s->tile3_ok = false;
    // (measured round 2: within 3 % of k_force_tile2 at 8 M cells, slower at low density; the planner therefore keeps
    // k_force_tile2 unless asked: CRB_FORCE_KERNEL=tile3)
    if (s->kernel_choice == 4) {
        const size_t per_cand3 = sizeof(TileCand) + 5 * sizeof(float);
        const double halo3 = srows * 2.0 * rho * 1.05;
        double best3 = 0, fill3 = 0;
        const char* forced3 = std::getenv("CRB_TILE3_THREADS");
        for (int nt : {256, 128}) {
            if (forced3 && std::atoi(forced3) != nt) continue;
            const int ctas = 512 / nt;
            const size_t budget3 = (size_t)227 * 1024 / ctas - 1024;
            const double lanes = nt - trows * 8.0 - 4.0;  // expected padding of a row to 16 slots: 7.5
            for (int nchunks = 1; nchunks <= n2; ++nchunks) {
                const int chunk_len = (n2 + nchunks - 1) / nchunks;
                if (chunk_len > 61 && nchunks < n2) continue;
                if (nchunks > 1 && (n2 + nchunks - 2) / (nchunks - 1) == chunk_len) continue;
                const int kmax = chunk_len + 2;
                const size_t fixed3 = (size_t)(kT3Queue + 1) * nt * 6 + (size_t)kT2Rows * (kmax + 1) * sizeof(u32) + 64 +
                                      64 * per_cand3 + 128 + 1024 /* static */;
                if (budget3 <= fixed3) continue;
                const u32 cap3 = (u32)std::min<size_t>(((budget3 - fixed3) / per_cand3) / 32 * 32, 60000);
                const double ts = std::min(((double)cap3 - halo3) / per_t, lanes);
                if (ts < 8.0) continue;
                const double Rc = chunk_len * col;
                const int mseg = std::max(1, (int)std::ceil(Rc * 1.05 / ts));
                const double fill = Rc / mseg;
                const double score = fill * ctas;
                if (score <= best3 * 1.001) continue;
                best3 = score;
                TileLaunch& U = s->tile3;
                U = T;
                U.chunk_len = chunk_len;
                U.nchunks = (n2 + chunk_len - 1) / chunk_len;
                U.kmax = kmax;
                U.tseg = (u32)ts;
                U.cap = cap3;
                U.nseg = mseg + 1;
                s->tile3_threads = nt;
                s->tile3_smem = tile3_smem_bytes(U.cap, U.kmax, nt);
                s->tile3_ok = s->tile3_smem <= budget3 + 1024 && tile3_c32(U.cap) < 65536;
                fill3 = fill;
            }
        }
        if (s->tile3_ok) {
            // conservative cutoff^2 of the TF32 filter.  Filter coordinates are relative to the centre of the
            // staged window: |c0|, |c1| <= 2 voxels, |c2| <= (kmax + 1) / 2 voxels (+ 1 % for agents that sit on
            // a face).  Rounding to TF32 (cvt.rna, 10 explicit mantissa bits) moves a coordinate by at most
            // 2^-11 of its magnitude; the FP32 filter copy and the origin subtraction add <= 2 ext 2^-22 as in
            // the FP32 filter; the FP32 accumulation of the eight products and the TF32 split of |b|^2 are
            // covered by 2^-17 of the largest squared norm.
            TileLaunch& U = s->tile3;
            const int ax0 = 0, ax1 = s->D == 3 ? 1 : -1, ax2 = s->D - 1;
            double m0 = 2.02 * g.dx[ax0], m1 = ax1 >= 0 ? 2.02 * g.dx[ax1] : 0.0;
            double m2 = (0.5 * (U.kmax + 1) + 0.02) * g.dx[ax2];
            const double e0 = m0 * std::ldexp(1.0, -11), e1 = m1 * std::ldexp(1.0, -11), e2 = m2 * std::ldexp(1.0, -11);
            const double delta = std::sqrt(e0 * e0 + e1 * e1 + e2 * e2);
            const double r2max = m0 * m0 + m1 * m1 + m2 * m2;
            float cut2m = INFINITY;
            if (std::isfinite(cutoff) && cutoff > 0) {
                const double c = (cutoff + 2.0 * delta + 2.0 * ext * std::ldexp(1.0, -22)) * (1.0 + std::ldexp(1.0, -19));
                const double c2 = c * c + r2max * std::ldexp(1.0, -17);
                cut2m = (float)c2;
                if ((double)cut2m < c2) cut2m = std::nextafterf(cut2m, INFINITY);
            }
            U.cut2m = cut2m;
            if (!std::isfinite(cut2m)) s->tile3_ok = false;  // no cutoff: nothing to filter
        }
        if (s->tile3_ok && s->tile_ok && s->kernel_choice == 0) {
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, (int)s->cfg.device);
            const long tiles = (long)((T.x_count + 1) / 2) * ((g.ng[1] + 1) / 2) * s->tile3.nchunks;
            if (tiles < sms) s->tile3_ok = false;
        }
        if (s->tile3_ok && std::getenv("CRB_DEBUG_PLAN")) {
            const TileLaunch& U = s->tile3;
            std::fprintf(stderr, "[crb] tile3 plan: threads %d smem %zu cap %u tseg %u nseg %d chunk %d x %d "
                                 "rho %.2f expected targets/CTA %.0f cut2m %.4f\n",
                         s->tile3_threads, s->tile3_smem, U.cap, U.tseg, U.nseg, U.chunk_len, U.nchunks, rho, fill3,
                         (double)U.cut2m);
        }
    }
    if (s->tile2_ok && std::getenv("CRB_DEBUG_PLAN")) {
        const TileLaunch& U = s->tile2;
        std::fprintf(stderr, "[crb] tile2 plan: threads %d smem %zu cap %u tseg %u nseg %d chunk %d x %d "
                             "rho %.2f expected targets/CTA %.0f\n",
                     s->tile2_threads, s->tile2_smem, U.cap, U.tseg, U.nseg, U.chunk_len, U.nchunks, rho, plan_fill);
    }
}

// plane range of a launch: x_count < 0 => all owned planes (TileLaunch as planned)
struct PlaneRange {
    int x_begin = 0, x_count = -1;
    cudaStream_t st = nullptr;
};

// the tile kernels are compiled in their own translation units (crb_tile.cu, crb_tile2.cu, crb_tile3.cu: the library
// builds in parallel); crb_launch.cuh declares their launchers
int launch_tile_family(crb_sim* s, int which, const KParams& P, bool force_only, const PlaneRange& pr) {
    TileLaunch T = which == 3 ? s->tile3 : which == 2 ? s->tile2 : s->tile;
    if (pr.x_count >= 0) {
        T.x_begin = pr.x_begin;
        T.x_count = pr.x_count;
    }
    cudaStream_t st = pr.st ? pr.st : s->stream;
    const int inter = (int)s->cfg.interaction_kind;
    cudaError_t e;
    if (which == 1) {
        const u32 grid = (u32)T.x_count * (u32)s->g.ng[1] * (u32)T.nseg * (u32)T.nchunks;
        if (!grid) return 0;
        e = launch_k_force_tile(s->D, inter, force_only, P, T, grid, s->tile_threads, s->tile_smem, st);
    } else {
        const u32 nx2 = (u32)(T.x_count + 1) / 2, ny2 = (u32)(s->g.ng[1] + 1) / 2;
        const u32 grid = nx2 * ny2 * (u32)T.nseg * (u32)T.nchunks;
        if (!grid) return 0;
        e = which == 3 ? launch_k_force_tile3(s->D, inter, force_only, P, T, grid, s->tile3_threads, s->tile3_smem, st)
                       : launch_k_force_tile2(s->D, inter, force_only, P, T, grid, s->tile2_threads, s->tile2_smem, st);
    }
    if (e != cudaSuccess) return fail(s, CRB_ERR_CUDA, std::string("tile kernel launch: ") + cudaGetErrorString(e));
    return 0;
}

int launch_force(crb_sim* s, const KParams& P, bool force_only, const PlaneRange& pr = PlaneRange()) {
    if (s->tile3_ok) return launch_tile_family(s, 3, P, force_only, pr);
    if (s->tile2_ok) return launch_tile_family(s, 2, P, force_only, pr);
    if (s->tile_ok) return launch_tile_family(s, 1, P, force_only, pr);
    if (s->D == 3)
        launch_force_d<3>(s, P, force_only);
    else if (s->D == 2)
        launch_force_d<2>(s, P, force_only);
    else
        launch_force_d<1>(s, P, force_only);  // 1-D (Brownian1D, Langevin1D, NewtonDamped1D): per-agent kernel
    return 0;
}

// scan count_new -> start_alt, fill, reorder T(+X static) -> Y, swap.
// n_src: unsorted sources at [G, G + n_src) (owned agents of the step + arrivals, some of them marked
// CRB_KEY_GONE); n_new: agents that remain = destination slots [G, G + n_new).
// dev_counts: the numbers of sources / destinations live in s->cnt[1] / the scan total (crb_slab_step); n_src and
// n_new then only bound the grids
int sort_into_place(crb_sim* s, u64 n_src, u64 n_new, bool save_prev, bool ingest, bool dev_counts = false) {
    const u64 nv = s->nvox;
    stage_begin(s, 1);
    if (s->n_tiles == 1) {
        k_scan_single<<<1, kScanThreads, 0, s->stream>>>(s->count_new, s->start_alt, nv, (u32)s->G);
        stage_end(s, 1, 1);
    } else {
        k_scan_tiles<<<s->n_tiles, kScanThreads, 0, s->stream>>>(s->count_new, s->start_alt,
                                                                s->tile_sums, nv);
        k_scan_sums<<<1, kScanThreads, 0, s->stream>>>(s->tile_sums, s->n_tiles, s->start_alt + nv,
                                                        (u32)s->G);
        k_scan_add<<<s->n_tiles, kScanThreads, 0, s->stream>>>(s->start_alt, s->tile_sums, nv);
        stage_end(s, 1, 3);
    }
    stage_begin(s, 2);
    int nl = 0;
    if (n_src) {
        k_fill<<<blocks_for(n_src, 256 * kFillPerThread), 256, 0, s->stream>>>(s->newkey, s->rank_arb, s->start_alt,
                                                             s->tmp_idx, s->G, n_src, dev_counts ? s->cnt + 1 : nullptr);
        ++nl;
    }
    if (n_new) {
        ReorderParams R;
        std::memset(&R, 0, sizeof(R));
        R.A = s->X;
        R.B = s->T;
        R.Y = s->Y;
        R.newkey = s->newkey;
        R.start_new = s->start_alt;
        R.tmp_idx = s->tmp_idx;
        R.sortkey = s->sortkey;
        R.key_out = s->key_alt;
        R.slot_of_src = ingest ? s->slot_of_src : nullptr;
        R.n = n_new;
        R.first = s->G;
        R.D = s->D;
        R.V = s->V;
        R.n_mp = s->uniform ? 0 : s->n_mp;
        R.n_ip = s->uniform ? 0 : s->n_ip;
        R.save_prev = save_prev ? 1 : 0;
        R.has_species = 1;
        for (int d = 0; d < 3; ++d) R.mn[d] = d < s->D ? s->g.mn[d] : 0.0;
        R.n2 = (u32)s->g.ng[2];
        R.nvox_dev = dev_counts ? nv : 0;
        k_reorder<<<blocks_for(n_new, 256), 256, 0, s->stream>>>(R);
        ++nl;
    }
    stage_end(s, 2, nl);
    std::swap(s->X, s->Y);
    std::swap(s->key, s->key_alt);
    std::swap(s->start, s->start_alt);
    CU(s, cudaGetLastError());
    return 0;
}

// the captured kernel parameters (pointers, agent count, tile plan, uniform parameter rows) are baked into
// the graph nodes: any call that changes one of them drops the graphs
void graphs_invalidate(crb_sim* s) {
    for (int k = 0; k < 2; ++k) {
        if (s->step_graph[k]) cudaGraphExecDestroy(s->step_graph[k]);
        s->step_graph[k] = nullptr;
        s->step_graph_key[k] = nullptr;
    }
}

int check_latch(crb_sim* s) {
    if (s->kernels_only && s->rods) {
        // pipelined rod handle: the latch travels by a store into the page-locked words instead of a copy-engine
        // transfer (which would queue behind the snapshot download of the handle this one alternates with)
        k_rod_report<<<1, 32, 0, s->stream>>>(nullptr, s->err, nullptr, s->err_host);
    } else {
        CU(s, cudaMemcpyAsync(s->err_host, s->err, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    }
    CU(s, cudaStreamSynchronize(s->stream));
    if (s->err_host[1]) {
        const int code = s->err_host[0];
        u64 id;
        std::memcpy(&id, s->err_host + 2, sizeof(id));
        CU(s, cudaMemsetAsync(s->err, 0, 4 * sizeof(int), s->stream));
        switch (code) {
            case CRB_ERR_CALC:
                return fail(s, code,
                            "CalcError: identical position for two objects. Cannot Calculate force "
                            "in Mie potential (agent id " + std::to_string(id) + ")");
            case CRB_ERR_BOUNDARY:
                return fail(s, code, "BoundaryError: Particle is out of domain (agent id " +
                                         std::to_string(id) + ")");
            case CRB_ERR_INDEX:
                return fail(s, code, "IndexError: could not find voxel index, agent id beyond the noise table, or an agent "
                                     "crossed more than one voxel plane into a non-adjacent slab in one step (agent " +
                                         std::to_string(id) + ")");
            case CRB_ERR_TIMEOUT:
                return fail(s, code, "slab exchange: a neighbouring slab did not deliver its buffer of exchange step " +
                                         std::to_string(id) + " in time (crb_slab_step must be called with the same "
                                         "number of steps on every rank; a single host thread must interleave the ranks)");
            case CRB_ERR_STATE:
                return fail(s, code, "slab exchange: the neighbouring slab uses another parameter mode (uniform row vs "
                                     "per-agent columns, or a different uniform row): declare the same row with "
                                     "crb_set_uniform_params on every rank, or on none");
            case CRB_ERR_INVALID_ARG:
                return fail(s, code, "rods: all RodMechanics parameters except spring_length and growth_rate, and all "
                                     "interaction parameters, must be uniform (row " + std::to_string(id) +
                                         " of the rows given to crb_set_agents_async differs from row 0)");
            default: return fail(s, code, "device error " + std::to_string(code));
        }
    }
    return 0;
}

// host rows (agent-major, `width` doubles) -> device columns
int upload_rows_copy(crb_sim* s, const double* host, int width, u64 n, u64 stage_off) {
    if (!n) return 0;
    if (width > 8) return fail(s, CRB_ERR_INVALID_ARG, "row width > 8");
    CU(s, cudaMemcpyAsync(s->stage + stage_off, host, n * width * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    return 0;
}
int upload_rows_unpack(crb_sim* s, int width, double* const* cols, u64 n, u64 stage_off) {
    if (!n) return 0;
    double* c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < width; ++k) c[k] = cols[k] ? cols[k] + s->G : nullptr;  // owned agents start at G
    k_unpack_rows<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->stage + stage_off, width, c[0], c[1], c[2], c[3],
                                                            c[4], c[5], c[6], c[7], n);
    s->launches++;
    return 0;
}
int upload_rows(crb_sim* s, const double* host, int width, double* const* cols, u64 n) {
    int rc;
    if ((rc = upload_rows_copy(s, host, width, n, 0))) return rc;
    return upload_rows_unpack(s, width, cols, n, 0);
}

int download_rows(crb_sim* s, double* host, int width, const double* const* cols, u64 n) {
    if (!n) return 0;
    const double* c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < width; ++k) c[k] = cols[k] ? cols[k] + s->G : nullptr;
    k_pack_rows<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->stage, width, c[0], c[1], c[2], c[3],
                                                          c[4], c[5], c[6], c[7], n);
    s->launches++;
    CU(s, cudaMemcpyAsync(host, s->stage, n * width * sizeof(double), cudaMemcpyDeviceToHost,
                          s->stream));
    return 0;
}

}  // namespace

#include "crb_rods_host.cuh"

// loads every kernel crb_slab_step can launch on this handle (see crb_launch.cuh kPreloadOnly)
template <int D>
static void preload_agent_kernels(cudaError_t& e) {
    cudaFuncAttributes at;
    auto one = [&](auto k) {
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, k);
    };
    one(k_force_integrate<D, CRB_INT_CELL_SORTING, true, false>);
    one(k_force_integrate<D, CRB_INT_CELL_SORTING, false, false>);
    one(k_force_integrate<D, CRB_INT_CELL_SORTING_BENCH, true, false>);
    one(k_force_integrate<D, CRB_INT_CELL_SORTING_BENCH, false, false>);
    one(k_force_integrate<D, CRB_INT_MORSE, true, false>);
    one(k_force_integrate<D, CRB_INT_MORSE, false, false>);
    one(k_force_integrate<D, CRB_INT_BOUND_LJ, true, false>);
    one(k_force_integrate<D, CRB_INT_BOUND_LJ, false, false>);
    one(k_force_integrate<D, CRB_INT_MIE, true, false>);
    one(k_force_integrate<D, CRB_INT_MIE, false, false>);
    one(k_force_integrate<D, CRB_INT_NONE, true, false>);
    one(k_force_integrate<D, CRB_INT_NONE, false, false>);
}

static int slab_preload(crb_sim* s) {
    if (s->slab_preloaded) return 0;
    cudaError_t e = cudaSuccess;
    cudaFuncAttributes at;
    auto one = [&](auto k) {
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, k);
    };
    one(k_set_u64);
    one(k_zero_u32);
    one(k_halo_export);
    one(k_halo_import);
    one(k_slab_publish);
    one(k_slab_wait);
    one(k_mig_export);
    one(k_mig_import_dev);
    one(k_slab_set_count);
    one(k_scan_tiles);
    one(k_scan_single);
    one(k_scan_sums);
    one(k_scan_add);
    one(k_fill);
    one(k_reorder);
    if (s->D == 3) preload_agent_kernels<3>(e);
    else if (s->D == 2) preload_agent_kernels<2>(e);
    else preload_agent_kernels<1>(e);
    if (s->D >= 2 && s->cfg.interaction_kind != CRB_INT_NONE) {
        const int inter = (int)s->cfg.interaction_kind;
        KParams P = make_params(s);
        TileLaunch T = s->tile;
        if (e == cudaSuccess) e = launch_k_force_tile(s->D, inter, false, P, T, kPreloadOnly, kTileThreads, 0, s->stream);
        for (int nt : {128, 256}) {
            if (e == cudaSuccess) e = launch_k_force_tile2(s->D, inter, false, P, T, kPreloadOnly, nt, 0, s->stream);
            if (e == cudaSuccess) e = launch_k_force_tile3(s->D, inter, false, P, T, kPreloadOnly, nt, 0, s->stream);
        }
    }
    if (e != cudaSuccess) return fail(s, CRB_ERR_CUDA, std::string("kernel preload: ") + cudaGetErrorString(e));
    s->slab_preloaded = true;
    return 0;
}


// =================================================================================================
extern "C" {

uint32_t crb_abi_version(void) { return CRB_ABI_VERSION; }

uint32_t crb_n_mech_params(uint32_t kind) {
    switch (kind) {
        case CRB_MECH_NEWTON_DAMPED: return 2;
        case CRB_MECH_BROWNIAN: return 2;
        case CRB_MECH_LANGEVIN: return 3;
        case CRB_MECH_ROD: return 8;
    }
    return 0;
}

uint32_t crb_n_inter_params(uint32_t kind) {
    switch (kind) {
        case CRB_INT_CELL_SORTING:
        case CRB_INT_CELL_SORTING_BENCH: return 3;
        case CRB_INT_MORSE: return 4;
        case CRB_INT_BOUND_LJ: return 4;
        case CRB_INT_MIE: return 6;
        case CRB_INT_NONE: return 0;
    }
    return 0;
}

const char* crb_last_error(const crb_sim* sim) {
    return sim ? sim->error.c_str() : g_create_error.c_str();
}

// get_decomp_res (cartesian_cuboid_n.rs:19-59)
int crb_decomp_res(uint64_t n_voxel, uint64_t n_regions, uint64_t* n_out, uint64_t* m_out,
                   uint64_t* avg_out) {
    if (!n_regions) return CRB_ERR_DECOMPOSE;
    long long average_len = (long long)std::ceil((double)n_voxel / (double)n_regions);
    long long n = (long long)n_regions, m = 0;
    for (uint64_t it = 0; it < n_regions; ++it) {
        const long long residue = (long long)n_voxel - average_len * n - (average_len - 1) * m;
        if (residue == 0) {
            *n_out = (uint64_t)n;
            *m_out = (uint64_t)m;
            *avg_out = (uint64_t)average_len;
            return 0;
        }
        if (residue > 0) {
            if (n == (long long)n_regions) {
                average_len += 1;
                m = 0;
            }
        } else {
            n -= 1;
            m += 1;
        }
    }
    return CRB_ERR_DECOMPOSE;
}

// CartesianCuboid::from_boundaries_and_interaction_range (cartesian_cuboid_n.rs:183-222)
int crb_domain_from_interaction_range(uint32_t dim, const double* min, const double* max,
                                      double interaction_range, uint64_t* n_voxels, double* dx) {
    for (uint32_t i = 0; i < dim; ++i) {
        if (min[i] >= max[i]) return CRB_ERR_BOUNDARY;
        const double n = std::floor((max[i] - min[i]) / interaction_range);
        if (!(n > -1.0 && n < 18446744073709551616.0)) return CRB_ERR_BOUNDARY;
        const uint64_t m = n <= 0.0 ? 0 : (uint64_t)n;
        n_voxels[i] = m > 1 ? m : 1;
        dx[i] = (max[i] - min[i]) / n;
    }
    return 0;
}

// CartesianCuboid::from_boundaries_and_n_voxels (cartesian_cuboid_n.rs:226-256)
int crb_domain_from_n_voxels(uint32_t dim, const double* min, const double* max,
                             const uint64_t* n_voxels, double* dx) {
    for (uint32_t i = 0; i < dim; ++i) {
        if (min[i] >= max[i]) return CRB_ERR_BOUNDARY;
        dx[i] = (max[i] - min[i]) / (double)n_voxels[i];
    }
    return 0;
}

int crb_create(const crb_config* cfg, crb_sim** out) {
    if (!cfg || !out) return fail(nullptr, CRB_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (cfg->abi_version != CRB_ABI_VERSION)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "abi_version mismatch");
    if (cfg->dim < 1 || cfg->dim > 3)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "dim must be 1, 2 or 3");
    if (cfg->mechanics_kind > CRB_MECH_ROD)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "unknown mechanics_kind");
    const bool rods = cfg->mechanics_kind == CRB_MECH_ROD;
    if (rods) {
        if (cfg->dim != 3 || cfg->interaction_kind != CRB_INT_MORSE)
            return fail(nullptr, CRB_ERR_INVALID_ARG,
                        "CRB_MECH_ROD needs dim == 3 and interaction_kind == CRB_INT_MORSE (RodInteraction<MorsePotential>)");
        if (cfg->rod_n_vertices < 2 || cfg->rod_n_vertices > CRB_MAX_ROD_VERTICES)
            return fail(nullptr, CRB_ERR_INVALID_ARG, "rod_n_vertices must be in [2, 8]");
    }
    if (cfg->interaction_kind > CRB_INT_NONE)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "unknown interaction_kind");
    if (cfg->domain_force_kind > CRB_DOMAIN_FORCE_ROD_GEL)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "unknown domain_force_kind");
    if (rods && cfg->domain_force_kind == CRB_DOMAIN_FORCE_SPRING)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "CRB_MECH_ROD: DomainForce is CRB_DOMAIN_FORCE_ROD_GEL (CartesianSubDomainRods)");
    if (!rods && cfg->domain_force_kind == CRB_DOMAIN_FORCE_ROD_GEL)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "CRB_DOMAIN_FORCE_ROD_GEL needs mechanics_kind CRB_MECH_ROD");
    if (cfg->solver_order > 2)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "solver_order must be 0, 1 or 2");
    if (cfg->world_size == 0 || cfg->rank >= cfg->world_size)
        return fail(nullptr, CRB_ERR_INVALID_ARG, "rank must be < world_size");
    if (cfg->world_size > cfg->n_voxels[0])
        return fail(nullptr, CRB_ERR_DECOMPOSE, "DecomposeError: more slabs than voxel planes along axis 0");
    if (!(cfg->dt > 0.0)) return fail(nullptr, CRB_ERR_INVALID_ARG, "dt must be positive");
    const int D = (int)cfg->dim;
    u64 nv = 1;
    for (int i = 0; i < D; ++i) {
        if (cfg->n_voxels[i] == 0 || !(cfg->domain_min[i] < cfg->domain_max[i]))
            return fail(nullptr, CRB_ERR_BOUNDARY, "Min must be smaller than Max for domain boundaries");
        nv *= cfg->n_voxels[i];
    }
    if (cfg->capacity == 0 || cfg->capacity >= (1ull << 31))
        return fail(nullptr, CRB_ERR_INVALID_ARG, "capacity must be in [1, 2^31)");

    crb_sim* s = new crb_sim();
    s->cfg = *cfg;
    s->D = D;
    s->V = rods ? (int)cfg->rod_n_vertices : 1;
    s->rods = rods;
    s->n_mp = (int)crb_n_mech_params(cfg->mechanics_kind);
    s->n_ip = (int)crb_n_inter_params(cfg->interaction_kind);
    s->cap = cfg->capacity;
    s->rank = (int)cfg->rank;
    s->world = (int)cfg->world_size;
    s->has_lo = s->rank > 0;
    s->has_hi = s->rank + 1 < s->world;
    Grid& g = s->g;
    std::memset(&g, 0, sizeof(g));
    g.D = D;
    g.ng[0] = (int)cfg->n_voxels[0];
    g.ng[1] = D == 3 ? (int)cfg->n_voxels[1] : 1;
    g.ng[2] = D >= 2 ? (int)cfg->n_voxels[D - 1] : 1;  // 1-D: the only axis is the slab axis a0
    g.cax[0] = 0;
    g.cax[1] = D == 3 ? 1 : -1;
    g.cax[2] = D >= 2 ? D - 1 : -1;
    {
        // slab of whole voxel planes along axis 0: ranks below `rem` own one plane more; one ghost
        // plane towards every neighbour
        const int n0 = g.ng[0], W = s->world, r = s->rank;
        const int base = n0 / W, rem = n0 % W;
        const int nown = base + (r < rem ? 1 : 0);
        const int gp_lo = r * base + std::min(r, rem);
        g.own_lo = s->has_lo ? 1 : 0;
        g.own_hi = g.own_lo + nown;
        g.L0 = g.own_hi + (s->has_hi ? 1 : 0);
        g.p0 = gp_lo - g.own_lo;
    }
    for (int i = 0; i < D; ++i) {
        g.mn[i] = cfg->domain_min[i];
        g.mx[i] = cfg->domain_max[i];
        g.dx[i] = cfg->dx[i];
    }
    const u64 nsub = cfg->ref_n_subdomains ? cfg->ref_n_subdomains : 1;
    uint64_t sn, sm, savg;
    if (crb_decomp_res(nv, nsub, &sn, &sm, &savg) != 0) {
        delete s;
        return fail(nullptr, CRB_ERR_DECOMPOSE, "DecomposeError: Could not find a suiting decomposition");
    }
    g.sm_n = sn;
    g.sm_avg = savg;
    g.sm_switch = sn * savg;
    g.sm_nsub = nsub;
    s->nvox = g.nvox_local();
    if (s->nvox >= (1ull << 31)) {
        delete s;
        return fail(nullptr, CRB_ERR_INVALID_ARG, "more than 2^31 voxels on one device");
    }
    {
        const u64 plane = (u64)g.ng[1] * g.ng[2];
        // the same on every rank (buffer sizes must agree): floor(planes / world) owned planes
        const u64 nown = std::max<u64>((u64)g.ng[0] / (u64)s->world, 1);
        if (s->world > 1) {
            s->G = cfg->ghost_capacity ? cfg->ghost_capacity : (s->cap / nown) * 3 / 2 + 4096;
            s->mcap = cfg->migration_capacity ? cfg->migration_capacity : std::max<u64>(s->G / 4, 4096);
        }
        s->cap_total = s->cap + 2 * s->G + (s->world > 1 ? 2 * s->mcap : 0);
        if (s->cap_total >= (1ull << 32) - 2) {
            delete s;
            return fail(nullptr, CRB_ERR_INVALID_ARG, "capacity + ghost capacity exceeds 2^32");
        }
        // exchange buffer layouts
        HaloLayout& H = s->hl;
        std::memset(&H, 0, sizeof(H));
        H.gcap = s->G;
        H.plane = plane;
        H.D = D;
        H.n_ip = s->n_ip;  // columns are reserved; only copied when the parameters are not uniform
        u64 off = 64;
        auto take = [&](u64 bytes) {
            const u64 o = off;
            off += (bytes + 63) / 64 * 64;
            return o;
        };
        H.off_rel = take((plane + 1) * sizeof(u32));
        for (int c = 0; c < D; ++c) H.off_pos[c] = take(H.gcap * sizeof(double));
        H.off_posf = take(H.gcap * sizeof(float4));
        H.off_meta = take(H.gcap * sizeof(u32));
        H.off_sp = take(H.gcap);
        for (int k = 0; k < s->n_ip; ++k) H.off_ip[k] = take(H.gcap * sizeof(double));
        H.bytes = off;
        MigLayout& M = s->ml;
        std::memset(&M, 0, sizeof(M));
        M.mcap = s->mcap;
        M.D = D;
        M.n_mp = s->n_mp;
        M.n_ip = s->n_ip;
        off = 64;
        for (int c = 0; c < 6 * D; ++c) M.off_dyn[c] = take(M.mcap * sizeof(double));
        for (int k = 0; k < s->n_mp; ++k) M.off_mp[k] = take(M.mcap * sizeof(double));
        for (int k = 0; k < s->n_ip; ++k) M.off_ip[k] = take(M.mcap * sizeof(double));
        M.off_id = take(M.mcap * sizeof(u64));
        M.off_sortkey = take(M.mcap * sizeof(u64));
        M.off_gvox = take(M.mcap * sizeof(u64));
        M.off_hlen = take(M.mcap);
        M.off_sp = take(M.mcap);
        M.bytes = off;
        if (rods) {
            // rods travel with their own layouts (crb_rods.cuh); crb_slab_halo_bytes / crb_slab_migration_bytes
            // report their sizes
            const int W = 3 * (int)cfg->rod_n_vertices;
            RodHaloLayout& RH = s->rhl;
            std::memset(&RH, 0, sizeof(RH));
            RH.gcap = s->G;
            RH.plane = plane;
            RH.W = W;
            off = 64;
            RH.off_rel = take((plane + 1) * sizeof(u32));
            RH.off_pos = take(RH.gcap * W * sizeof(double));
            RH.off_cen = take(RH.gcap * 4 * sizeof(double));
            RH.bytes = off;
            RodMigLayout& RM = s->rml;
            std::memset(&RM, 0, sizeof(RM));
            RM.mcap = s->mcap;
            RM.W = W;
            off = 64;
            RM.off_dyn = take(RM.mcap * 6 * W * sizeof(double));
            RM.off_len = take(RM.mcap * sizeof(double));
            RM.off_growth = take(RM.mcap * sizeof(double));
            RM.off_cen = take(RM.mcap * 4 * sizeof(double));
            RM.off_ident = take(RM.mcap * 3 * sizeof(u64));
            RM.off_parent = take(RM.mcap * 3 * sizeof(u64));
            RM.off_sortkey = take(RM.mcap * sizeof(u64));
            RM.off_gvox = take(RM.mcap * sizeof(u64));
            RM.off_hlen = take(RM.mcap);
            RM.bytes = off;
            H.bytes = RH.bytes;
            M.bytes = RM.bytes;
        }
    }
    s->graphs_off = std::getenv("CRB_NO_GRAPH") != nullptr;
    if (const char* kc = std::getenv("CRB_FORCE_KERNEL")) {
        if (!std::strcmp(kc, "agent")) s->kernel_choice = 1;
        if (!std::strcmp(kc, "tile")) s->kernel_choice = 2;
        if (!std::strcmp(kc, "tile2")) s->kernel_choice = 3;
        if (!std::strcmp(kc, "tile3")) s->kernel_choice = 4;
    }

    auto bail = [&](int rc) {
        g_create_error = s->error;
        crb_destroy(s);
        return rc;
    };
    cudaError_t e = cudaSetDevice((int)cfg->device);
    if (e != cudaSuccess) {
        s->error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return bail(CRB_ERR_CUDA);
    }
    {
        int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        e = cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&s->stream2, cudaStreamNonBlocking, prio_lo);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_interior, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        s->error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
        return bail(CRB_ERR_CUDA);
    }
    int rc;
    if (rods) {
        if ((rc = rods_alloc(s))) return bail(rc);
    } else {
        if ((rc = alloc_state(s, s->X, true, true))) return bail(rc);
        if ((rc = alloc_state(s, s->Y, true, true))) return bail(rc);
        if ((rc = alloc_state(s, s->T, true, false))) return bail(rc);
    }
    s->n_tiles = (u32)((s->nvox + kScanTile - 1) / kScanTile);
    if ((rc = dalloc(s, &s->key, s->cap_total)) || (rc = dalloc(s, &s->key_alt, s->cap_total)) ||
        (rc = dalloc(s, &s->start, s->nvox + 1)) || (rc = dalloc(s, &s->start_alt, s->nvox + 1)) ||
        (rc = dalloc(s, &s->count_new, s->nvox + 1)) || (rc = dalloc(s, &s->newkey, s->cap_total)) ||
        (rc = dalloc(s, &s->rank_arb, s->cap_total)) || (rc = dalloc(s, &s->tmp_idx, s->cap_total)) ||
        (rc = dalloc(s, &s->slot_of_src, s->cap_total)) || (rc = dalloc(s, &s->sortkey, s->cap_total)) ||
        (rc = dalloc(s, &s->tile_sums, s->n_tiles + 1)) || (rc = dalloc(s, &s->stage, s->cap_total * (rods ? 3 * s->V : 8))) ||
        (rc = dalloc(s, &s->err, 4)) || (rc = dalloc(s, &s->scratch_u64, 16)))
        return bail(rc);
    if ((rc = dalloc(s, &s->cnt, 8))) return bail(rc);
    if (s->world > 1 && !rods) {
        s->arena_halo = (s->hl.bytes + 255) / 256 * 256;
        s->arena_mig = (s->ml.bytes + 255) / 256 * 256;
        s->arena_bytes = 256 + 4 * s->arena_halo + 4 * s->arena_mig;
        // a plain cudaMalloc allocation of its own: cudaIpcGetMemHandle exports whole allocations
        e = cudaMalloc((void**)&s->arena, s->arena_bytes);
        if (e == cudaSuccess) e = cudaMemset(s->arena, 0, s->arena_bytes);
        if (e != cudaSuccess) {
            s->error = std::string("exchange arena: ") + cudaGetErrorString(e);
            return bail(CRB_ERR_CUDA);
        }
    }
    e = cudaMallocHost((void**)&s->err_host, 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&s->cnt_host, 8 * sizeof(u64));
    if (e != cudaSuccess) {
        s->error = std::string("cudaMallocHost: ") + cudaGetErrorString(e);
        return bail(CRB_ERR_CUDA);
    }
    cudaMemsetAsync(s->err, 0, 4 * sizeof(int), s->stream);
    cudaMemsetAsync(s->start, 0, (s->nvox + 1) * sizeof(u32), s->stream);
    e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        s->error = std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e);
        return bail(CRB_ERR_CUDA);
    }
    *out = s;
    return 0;
}

void crb_destroy(crb_sim* s) {
    if (!s) return;
    cudaSetDevice((int)s->cfg.device);
    if (s->stream2) cudaStreamSynchronize(s->stream2);
    if (s->stream) cudaStreamSynchronize(s->stream);
    timers_resolve(s);
    graphs_invalidate(s);
    for (int k = 0; k < 2; ++k)
        if (s->peer_arena[k] && s->peer_ipc[k]) cudaIpcCloseMemHandle(s->peer_arena[k]);
    if (s->arena) cudaFree(s->arena);
    for (void* p : s->allocs) cudaFree(p);
    if (s->err_host) cudaFreeHost(s->err_host);
    if (s->cnt_host) cudaFreeHost(s->cnt_host);
    if (s->snap_pending) cudaEventSynchronize(s->ev_snap_done);
    if (s->stream_io) cudaStreamSynchronize(s->stream_io);
    if (s->ev_snap_packed) cudaEventDestroy(s->ev_snap_packed);
    if (s->ev_snap_done) cudaEventDestroy(s->ev_snap_done);
    if (s->stream_io) cudaStreamDestroy(s->stream_io);
    if (s->snap) cudaFree(s->snap);
    if (s->snap_hlen_host) cudaFreeHost(s->snap_hlen_host);
    if (s->ev_ready) cudaEventDestroy(s->ev_ready);
    if (s->ev_interior) cudaEventDestroy(s->ev_interior);
    if (s->stream2) cudaStreamDestroy(s->stream2);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

uint64_t crb_n_agents(const crb_sim* s) { return s ? s->n : 0; }
uint64_t crb_iteration(const crb_sim* s) { return s ? s->iteration : 0; }
int crb_set_iteration(crb_sim* s, uint64_t iteration) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, "crb_set_iteration inside an open slab step");
    s->iteration = iteration;
    return 0;
}
uint64_t crb_kernel_launches(const crb_sim* s) { return s ? s->launches : 0; }
const char* crb_force_kernel_name(const crb_sim* s) {
    if (!s) return "";
    if (s->rods) return "k_rod_force";
    if (s->tile3_ok) return s->tile3_threads == 256 ? "k_force_tile3<256 threads, 2 CTAs/SM, tf32 mma filter>" : "k_force_tile3<128 threads, 4 CTAs/SM, tf32 mma filter>";
    if (s->tile2_ok) return s->tile2_threads == 256 ? "k_force_tile2<256 threads, 2 CTAs/SM>" : "k_force_tile2<128 threads, 4 CTAs/SM>";
    if (s->tile_ok) return "k_force_tile";
    return "k_force_integrate";
}

// do all n rows of `width` doubles equal `row`?  One pass on all host cores (8 M agents are 320 MB of parameter rows;
// the single-threaded scan was a visible part of crb_set_agents for a host that does not declare the uniform row)
static bool rows_all_equal(const double* rows, u64 n, int width, const double* row) {
    if (!n || !width) return true;
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const unsigned nth = n < 65536 ? 1u : hw;
    std::vector<int> bad(nth, 0);
    auto scan = [&](unsigned t) {
        const u64 a = n * t / nth, b = n * (t + 1) / nth;
        int acc = 0;
        for (u64 i = a; i < b && !acc; i += 4096) {  // early exit per block of rows, branch-free inside
            const u64 e = std::min(b, i + 4096);
            for (u64 j = i; j < e; ++j)
                for (int k = 0; k < width; ++k) acc |= rows[j * width + k] != row[k];
        }
        bad[t] = acc;
    };
    if (nth == 1) {
        scan(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nth; ++t) th.emplace_back(scan, t);
        for (auto& x : th) x.join();
    }
    for (int b : bad)
        if (b) return false;
    return true;
}

static int set_agents_impl(crb_sim* s, uint64_t n, const double* pos, const double* vel,
                           const double* mech_params, const double* inter_params, const uint8_t* species,
                           const uint64_t* ids, bool async) {
    if (!s) return CRB_ERR_INVALID_ARG;
    graphs_invalidate(s);
    if (n > s->cap) return fail(s, CRB_ERR_CAPACITY, "more agents than capacity");
    if (n && !pos) return fail(s, CRB_ERR_INVALID_ARG, "pos is required");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods)
        return async ? rods_set_agents_async(s, n, pos, vel, mech_params, inter_params, (const u64*)ids)
                     : rods_set_agents(s, n, pos, vel, mech_params, inter_params, (const u64*)ids);
    const int D = s->D;
    // uniform-parameter fast path: all agents share one parameter row
    if (s->uniform_declared) {
        s->uniform = true;
    } else if (s->world > 1) {
        // slabs: whether ALL agents share one parameter row is a property of the whole population, not of the
        // subset this rank was handed (a slab that starts empty, or one whose own agents happen to agree, must not
        // take the uniform path while a neighbour ships per-agent columns).  Without a declaration
        // (crb_set_uniform_params on EVERY rank) a slab keeps per-agent parameter columns.
        if (n && (!mech_params || !inter_params))
            return fail(s, CRB_ERR_INVALID_ARG,
                        "slab handle: mech_params and inter_params are required unless crb_set_uniform_params "
                        "declared one row on every rank");
        s->uniform = false;
        for (int k = 0; k < s->n_mp; ++k) s->uni.mp[k] = 0.0;
        for (int k = 0; k < s->n_ip; ++k) s->uni.ip[k] = 0.0;
    } else {
        if (n && (!mech_params || !inter_params))
            return fail(s, CRB_ERR_INVALID_ARG, "mech_params and inter_params are required");
        s->uniform = true;
        for (int k = 0; k < s->n_mp; ++k) s->uni.mp[k] = n ? mech_params[k] : 0.0;
        for (int k = 0; k < s->n_ip; ++k) s->uni.ip[k] = n ? inter_params[k] : 0.0;
        s->uniform = rows_all_equal(mech_params, n, s->n_mp, s->uni.mp) && rows_all_equal(inter_params, n, s->n_ip, s->uni.ip);
    }
    // the uniform-parameter kernels evaluate the Morse exponential without overflow handling (exp_bounded): its
    // largest argument, potential_stiffness * (R1 + R2) at distance 0, must stay below 700; such a parameter row
    // (physically meaningless: a force of 1e304) takes the per-agent-parameter kernels with CUDA's exp() instead
    if (s->uniform && s->cfg.interaction_kind == CRB_INT_MORSE && !(s->uni.ip[1] * (s->uni.ip[0] + s->uni.ip[0]) < 700.0)) {
        if (s->uniform_declared || !inter_params)
            return fail(s, CRB_ERR_CALC, "CalcError: MorsePotential potential_stiffness * 2 radius >= 700 overflows exp in f64");
        s->uniform = false;
    }
    int rc;
    const bool has_vel = vel && s->cfg.mechanics_kind != CRB_MECH_BROWNIAN;
    if (async) {
        // both host transfers first, then the kernels: a transfer queued behind an unpack kernel would wait for that
        // kernel to find free SMs next to the force pass of the handle this one alternates with, and the copy
        // engine would idle meanwhile (the staging buffer holds 8 doubles per agent: pos and vel rows side by side)
        if ((rc = upload_rows_copy(s, pos, D, n, 0))) return rc;
        if (has_vel && (rc = upload_rows_copy(s, vel, D, n, n * D))) return rc;
        if (species) CU(s, cudaMemcpyAsync(s->X.species + s->G, species, n, cudaMemcpyHostToDevice, s->stream));
        if (ids) CU(s, cudaMemcpyAsync(s->X.id + s->G, ids, n * sizeof(u64), cudaMemcpyHostToDevice, s->stream));
        if ((rc = upload_rows_unpack(s, D, s->T.pos, n, 0))) return rc;
        if (has_vel && (rc = upload_rows_unpack(s, D, s->T.vel, n, n * D))) return rc;
    } else {
        if ((rc = upload_rows(s, pos, D, s->T.pos, n))) return rc;
        if (has_vel && (rc = upload_rows(s, vel, D, s->T.vel, n))) return rc;
    }
    if (async) {
        // kernels only (see k_zero_cols)
        s->kernels_only = true;
        ZeroCols Z;
        std::memset(&Z, 0, sizeof(Z));
        for (int d = 0; d < D; ++d) {
            if (!has_vel) Z.col[Z.n_cols++] = s->T.vel[d] + s->G;
            Z.col[Z.n_cols++] = s->T.h1x[d] + s->G;
            Z.col[Z.n_cols++] = s->T.h1v[d] + s->G;
            Z.col[Z.n_cols++] = s->T.h0x[d] + s->G;
            Z.col[Z.n_cols++] = s->T.h0v[d] + s->G;
        }
        Z.bytes_a = s->T.hlen + s->G;
        Z.bytes_b = species ? nullptr : s->X.species + s->G;
        if (n) {
            k_zero_cols<<<blocks_for(n, 256), 256, 0, s->stream>>>(Z, n);
            s->launches++;
        }
    } else {
        if (!has_vel)
            for (int d = 0; d < D; ++d)
                CU(s, cudaMemsetAsync(s->T.vel[d] + s->G, 0, n * sizeof(double), s->stream));
        for (int d = 0; d < D; ++d) {
            CU(s, cudaMemsetAsync(s->T.h1x[d] + s->G, 0, n * sizeof(double), s->stream));
            CU(s, cudaMemsetAsync(s->T.h1v[d] + s->G, 0, n * sizeof(double), s->stream));
            CU(s, cudaMemsetAsync(s->T.h0x[d] + s->G, 0, n * sizeof(double), s->stream));
            CU(s, cudaMemsetAsync(s->T.h0v[d] + s->G, 0, n * sizeof(double), s->stream));
        }
        CU(s, cudaMemsetAsync(s->T.hlen + s->G, 0, n, s->stream));
    }
    if (!s->uniform) {
        if ((rc = upload_rows(s, mech_params, s->n_mp, s->X.mp, n))) return rc;
        if ((rc = upload_rows(s, inter_params, s->n_ip, s->X.ip, n))) return rc;
    }
    if (!async) {
        if (species)
            CU(s, cudaMemcpyAsync(s->X.species + s->G, species, n, cudaMemcpyHostToDevice, s->stream));
        else
            CU(s, cudaMemsetAsync(s->X.species + s->G, 0, n, s->stream));
    }
    if (ids) {
        if (!async) CU(s, cudaMemcpyAsync(s->X.id + s->G, ids, n * sizeof(u64), cudaMemcpyHostToDevice, s->stream));
    } else if (n) {
        k_iota_u64<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->X.id + s->G, n);
        s->launches++;
    }
    if (async) {
        k_zero_u32<<<blocks_for(s->nvox + 1, 256), 256, 0, s->stream>>>(s->count_new, s->nvox + 1, nullptr, 0, nullptr, 0);
        s->launches++;
    } else {
        CU(s, cudaMemsetAsync(s->count_new, 0, (s->nvox + 1) * sizeof(u32), s->stream));
    }
    s->n = n;
    s->slab_async_pending = false;  // the host-side count is authoritative again
    KParams P = make_params(s);
    if (n) {
        if (D == 3)
            k_ingest_keys<3><<<blocks_for(n, 256), 256, 0, s->stream>>>(P);
        else if (D == 2)
            k_ingest_keys<2><<<blocks_for(n, 256), 256, 0, s->stream>>>(P);
        else
            k_ingest_keys<1><<<blocks_for(n, 256), 256, 0, s->stream>>>(P);
        s->launches++;
    }
    // asynchronous form: ingestion errors (BoundaryError / IndexError of an agent outside the domain) stay latched
    // on the device and are returned by the next synchronising call
    if (!async && (rc = check_latch(s))) {
        s->n = 0;
        s->have_agents = false;
        return rc;
    }
    if ((rc = sort_into_place(s, n, n, true, true))) return rc;
    if (!async) CU(s, cudaStreamSynchronize(s->stream));
    plan_tile(s);
    s->have_agents = true;
    s->aux_settable = !async;
    s->last_step_n = 0;
    s->noise_cursor = 0;
    return 0;
}

int crb_set_agents(crb_sim* s, uint64_t n, const double* pos, const double* vel,
                   const double* mech_params, const double* inter_params, const uint8_t* species,
                   const uint64_t* ids) {
    return set_agents_impl(s, n, pos, vel, mech_params, inter_params, species, ids, false);
}

int crb_set_agents_async(crb_sim* s, uint64_t n, const double* pos, const double* vel,
                         const double* mech_params, const double* inter_params, const uint8_t* species,
                         const uint64_t* ids) {
    return set_agents_impl(s, n, pos, vel, mech_params, inter_params, species, ids, true);
}

int crb_set_uniform_params(crb_sim* s, const double* mech_row, const double* inter_row) {
    if (!s) return CRB_ERR_INVALID_ARG;
    graphs_invalidate(s);
    if (!mech_row && !inter_row) {
        s->uniform_declared = false;
        return 0;
    }
    if (!mech_row || !inter_row) return fail(s, CRB_ERR_INVALID_ARG, "both rows are required");
    for (int k = 0; k < s->n_mp; ++k) s->uni.mp[k] = mech_row[k];
    for (int k = 0; k < s->n_ip; ++k) s->uni.ip[k] = inter_row[k];
    s->uniform_declared = true;
    return 0;
}

void* crb_stream(crb_sim* s) { return s ? (void*)s->stream : nullptr; }

int crb_measure_fp64_peak(uint32_t device, double* tflops, double* ms_out) {
    if (!tflops) return CRB_ERR_INVALID_ARG;
    if (cudaSetDevice((int)device) != cudaSuccess) return CRB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, (int)device) != cudaSuccess) return CRB_ERR_CUDA;
    double* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return CRB_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k_dfma_peak<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return CRB_ERR_CUDA;
    const double flops = 2.0 * 64.0 * (double)iters * threads * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}

int crb_debug_rsqrt(uint32_t device, const double* x, double* y, uint64_t n) {
    if (!x || !y) return CRB_ERR_INVALID_ARG;
    if (cudaSetDevice((int)device) != cudaSuccess) return CRB_ERR_CUDA;
    double *dx = nullptr, *dy = nullptr;
    if (cudaMalloc(&dx, n * sizeof(double)) != cudaSuccess) return CRB_ERR_CUDA;
    if (cudaMalloc(&dy, n * sizeof(double)) != cudaSuccess) {
        cudaFree(dx);
        return CRB_ERR_CUDA;
    }
    cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice);
    k_rsqrt_test<<<blocks_for(n, 256), 256>>>(dx, dy, n);
    cudaMemcpy(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(dy);
    return cudaGetLastError() == cudaSuccess ? 0 : CRB_ERR_CUDA;
}

int crb_debug_exp(uint32_t device, const double* x, double* y, uint64_t n) {
    if (!x || !y) return CRB_ERR_INVALID_ARG;
    if (cudaSetDevice((int)device) != cudaSuccess) return CRB_ERR_CUDA;
    double *dx = nullptr, *dy = nullptr;
    if (cudaMalloc(&dx, n * sizeof(double)) != cudaSuccess) return CRB_ERR_CUDA;
    if (cudaMalloc(&dy, n * sizeof(double)) != cudaSuccess) {
        cudaFree(dx);
        return CRB_ERR_CUDA;
    }
    cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice);
    k_exp_test<<<blocks_for(n, 256), 256>>>(dx, dy, n);
    cudaMemcpy(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(dy);
    return cudaGetLastError() == cudaSuccess ? 0 : CRB_ERR_CUDA;
}

int crb_get_agents(crb_sim* s, uint64_t cap, double* pos, double* vel, uint64_t* ids,
                   uint64_t* voxel) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, kOpenStepMsg);
    if (s->slab_async_pending) return fail(s, CRB_ERR_STATE, kAsyncPendingMsg);
    if (cap < s->n) return fail(s, CRB_ERR_CAPACITY, "output buffers too small");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods) return rods_get_agents(s, cap, pos, vel, (u64*)ids, (u64*)voxel);
    int rc;
    const u64 n = s->n;
    if (pos) {
        if ((rc = download_rows(s, pos, s->D, s->X.pos, n))) return rc;
        CU(s, cudaStreamSynchronize(s->stream));
    }
    if (vel) {
        if (s->cfg.mechanics_kind == CRB_MECH_BROWNIAN) {
            std::memset(vel, 0, n * s->D * sizeof(double));
        } else {
            if ((rc = download_rows(s, vel, s->D, s->X.vel, n))) return rc;
            CU(s, cudaStreamSynchronize(s->stream));
        }
    }
    if (ids) CU(s, cudaMemcpyAsync(ids, s->X.id + s->G, n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    if (voxel && n) {
        const u64 base = (u64)s->g.p0 * s->g.ng[1] * s->g.ng[2];
        k_global_keys<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->key + s->G, (u64*)s->stage, base, n);
        s->launches++;
        CU(s, cudaMemcpyAsync(voxel, s->stage, n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    }
    CU(s, cudaStreamSynchronize(s->stream));
    return 0;
}

int crb_get_params(crb_sim* s, uint64_t cap, double* mech_params, double* inter_params,
                   uint8_t* species) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, kOpenStepMsg);
    if (s->slab_async_pending) return fail(s, CRB_ERR_STATE, kAsyncPendingMsg);
    if (cap < s->n) return fail(s, CRB_ERR_CAPACITY, "output buffers too small");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods) return rods_get_params(s, cap, mech_params, inter_params, species);
    const u64 n = s->n;
    int rc;
    if (mech_params) {
        if (s->uniform) {
            for (u64 i = 0; i < n; ++i)
                for (int k = 0; k < s->n_mp; ++k) mech_params[i * s->n_mp + k] = s->uni.mp[k];
        } else {
            if ((rc = download_rows(s, mech_params, s->n_mp, s->X.mp, n))) return rc;
            CU(s, cudaStreamSynchronize(s->stream));
        }
    }
    if (inter_params) {
        if (s->uniform) {
            for (u64 i = 0; i < n; ++i)
                for (int k = 0; k < s->n_ip; ++k) inter_params[i * s->n_ip + k] = s->uni.ip[k];
        } else {
            if ((rc = download_rows(s, inter_params, s->n_ip, s->X.ip, n))) return rc;
            CU(s, cudaStreamSynchronize(s->stream));
        }
    }
    if (species)
        CU(s, cudaMemcpyAsync(species, s->X.species + s->G, n, cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    return 0;
}

int crb_get_aux(crb_sim* s, uint64_t cap, double* hist_dx, double* hist_dv, uint32_t* hist_len) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, kOpenStepMsg);
    if (s->slab_async_pending) return fail(s, CRB_ERR_STATE, kAsyncPendingMsg);
    if (cap < s->n) return fail(s, CRB_ERR_CAPACITY, "output buffers too small");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods) return rods_get_aux(s, cap, hist_dx, hist_dv, hist_len);
    const u64 n = s->n;
    const int D = s->D;
    std::vector<u8> hl(n);
    CU(s, cudaMemcpyAsync(hl.data(), s->X.hlen + s->G, n, cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    int rc;
    // rows of [h0 (older), h1 (newest)], re-packed oldest-first per agent below
    std::vector<double> tmp(n * 2 * D);
    for (int which = 0; which < 2; ++which) {
        double* outp = which == 0 ? hist_dx : hist_dv;
        if (!outp) continue;
        const double* cols[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int d = 0; d < D; ++d) {
            cols[d] = which == 0 ? s->X.h0x[d] : s->X.h0v[d];
            cols[D + d] = which == 0 ? s->X.h1x[d] : s->X.h1v[d];
        }
        if ((rc = download_rows(s, tmp.data(), 2 * D, cols, n))) return rc;
        CU(s, cudaStreamSynchronize(s->stream));
        for (u64 i = 0; i < n; ++i) {
            double* o = outp + i * 2 * D;
            const double* t = tmp.data() + i * 2 * D;
            for (int k = 0; k < 2 * D; ++k) o[k] = 0.0;
            if (hl[i] == 2) {
                for (int k = 0; k < 2 * D; ++k) o[k] = t[k];
            } else if (hl[i] == 1) {
                for (int d = 0; d < D; ++d) o[d] = t[D + d];
            }
        }
    }
    if (hist_len)
        for (u64 i = 0; i < n; ++i) hist_len[i] = hl[i];
    return 0;
}

namespace {
__global__ void k_scatter_aux(State X, State T, const u32* __restrict__ slot, int D, u64 first, u64 n) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u64 i = first + t;
    const u32 d = slot[i];
    for (int c = 0; c < D; ++c) {
        X.h0x[c][d] = T.h0x[c][i];
        X.h1x[c][d] = T.h1x[c][i];
        X.h0v[c][d] = T.h0v[c][i];
        X.h1v[c][d] = T.h1v[c][i];
    }
    X.hlen[d] = T.hlen[i];
}
}  // namespace

int crb_set_aux(crb_sim* s, uint64_t n, const double* hist_dx, const double* hist_dv,
                const uint32_t* hist_len) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (s->rods) return fail(s, CRB_ERR_STATE, "crb_set_aux is not available for CRB_MECH_ROD");
    if (!s->aux_settable || n != s->n)
        return fail(s, CRB_ERR_STATE, "crb_set_aux must directly follow crb_set_agents with the same n");
    if (!hist_dx || !hist_dv || !hist_len) return fail(s, CRB_ERR_INVALID_ARG, "null argument");
    CU(s, cudaSetDevice((int)s->cfg.device));
    const int D = s->D;
    const u32 N = s->cfg.solver_order;
    // host rows are oldest-first; device keeps h1 = newest, h0 = older
    std::vector<double> rx(n * 2 * D), rv(n * 2 * D);
    std::vector<u8> hl(n);
    for (u64 i = 0; i < n; ++i) {
        const u32 len = std::min<u32>(hist_len[i], N);
        hl[i] = (u8)len;
        for (int d = 0; d < D; ++d) {
            double o_x = 0, n_x = 0, o_v = 0, n_v = 0;
            if (len == 2) {
                o_x = hist_dx[(i * 2 + 0) * D + d];
                n_x = hist_dx[(i * 2 + 1) * D + d];
                o_v = hist_dv[(i * 2 + 0) * D + d];
                n_v = hist_dv[(i * 2 + 1) * D + d];
            } else if (len == 1) {
                n_x = hist_dx[(i * 2 + 0) * D + d];
                n_v = hist_dv[(i * 2 + 0) * D + d];
            }
            rx[i * 2 * D + d] = o_x;
            rx[i * 2 * D + D + d] = n_x;
            rv[i * 2 * D + d] = o_v;
            rv[i * 2 * D + D + d] = n_v;
        }
    }
    double* cx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double* cv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int d = 0; d < D; ++d) {
        cx[d] = s->T.h0x[d];
        cx[D + d] = s->T.h1x[d];
        cv[d] = s->T.h0v[d];
        cv[D + d] = s->T.h1v[d];
    }
    int rc;
    if ((rc = upload_rows(s, rx.data(), 2 * D, cx, n))) return rc;
    CU(s, cudaStreamSynchronize(s->stream));
    if ((rc = upload_rows(s, rv.data(), 2 * D, cv, n))) return rc;
    CU(s, cudaMemcpyAsync(s->T.hlen + s->G, hl.data(), n, cudaMemcpyHostToDevice, s->stream));
    if (n) {
        k_scatter_aux<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->X, s->T, s->slot_of_src, D, s->G, n);
        s->launches++;
    }
    CU(s, cudaStreamSynchronize(s->stream));
    CU(s, cudaGetLastError());
    return 0;
}

int crb_set_noise(crb_sim* s, const double* z, uint64_t n_steps, uint64_t n_ids) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (s->cfg.noise_mode != CRB_NOISE_TABLE)
        return fail(s, CRB_ERR_STATE, "crb_set_noise needs noise_mode == CRB_NOISE_TABLE");
    CU(s, cudaSetDevice((int)s->cfg.device));
    const u64 count = n_steps * n_ids * s->D * s->V;  // rods: one number per vertex and coordinate
    if (s->noise) {
        cudaFree(s->noise);
        s->allocs.erase(std::find(s->allocs.begin(), s->allocs.end(), (void*)s->noise));
        s->noise = nullptr;
    }
    int rc;
    if ((rc = dalloc(s, &s->noise, count))) return rc;
    CU(s, cudaMemcpyAsync(s->noise, z, count * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    s->noise_steps = n_steps;
    s->noise_ids = n_ids;
    s->noise_cursor = 0;
    return 0;
}

namespace {
// stage A of one step: pair forces over the 3^D neighbourhood + fused integrate/boundary/new key for
// the owned agents (ghost planes must hold the neighbours' boundary agents on a slab)
int step_prepare(crb_sim* s, bool kernels_only = false) {
    const bool stochastic = s->cfg.noise_mode == CRB_NOISE_TABLE &&
                            s->cfg.mechanics_kind != CRB_MECH_NEWTON_DAMPED;
    if (stochastic && s->noise_cursor >= s->noise_steps)
        return fail(s, CRB_ERR_STATE, "noise table exhausted: call crb_set_noise");
    if (kernels_only) {
        k_zero_u32<<<blocks_for(s->nvox + 1, 256), 256, 0, s->stream>>>(s->count_new, s->nvox + 1, nullptr, 0, nullptr, 0);
        s->launches++;
    } else {
        CU(s, cudaMemsetAsync(s->count_new, 0, (s->nvox + 1) * sizeof(u32), s->stream));
    }
    return 0;
}
// force + integrate for the targets of a plane range (default: all owned planes) on pr.st
int step_launch(crb_sim* s, bool save_prev, const PlaneRange& pr) {
    KParams P = make_params(s);
    P.save_prev = save_prev ? 1 : 0;
    stage_begin(s, 0, pr.st);
    int rc;
    if ((rc = launch_force(s, P, false, pr))) return rc;
    stage_end(s, 0, s->n ? 1 : 0, pr.st);
    CU(s, cudaGetLastError());
    return 0;
}
// the whole force pass inside crb_slab_step: tile kernels take their targets from start[]; the per-agent kernel
// (non-uniform parameters) gets the device-side count
int step_launch_dev(crb_sim* s, bool save_prev) {
    if (s->tile_ok || s->tile2_ok || s->tile3_ok) return step_launch(s, save_prev, PlaneRange());
    KParams P = make_params(s);
    P.save_prev = save_prev ? 1 : 0;
    P.n = s->cap;
    P.n_dev = s->cnt;
    stage_begin(s, 0);
    int rc;
    if ((rc = launch_force(s, P, false, PlaneRange()))) return rc;
    stage_end(s, 0, 1);
    CU(s, cudaGetLastError());
    return 0;
}
void step_mark(crb_sim* s) {
    s->last_step_n = s->n;
    s->n_src = s->n;
    s->n_gone = 0;
}
int step_compute(crb_sim* s, bool save_prev) {
    int rc;
    if ((rc = step_prepare(s, s->kernels_only))) return rc;
    if (s->kernels_only && s->stream2 && s->n && (s->tile_ok || s->tile2_ok || s->tile3_ok) && !s->timers_on) {
        // pipelined host (two handles alternate on one device): the force pass runs on the handle's LOW-priority
        // stream, so that the short unpack / ingest / pack kernels of the other handle — queued on its high-priority
        // stream between two host transfers — are scheduled as soon as an SM has room instead of behind the ~40,000
        // CTAs of this pass; a late pack kernel would leave the copy engines idle
        CU(s, cudaEventRecord(s->ev_ready, s->stream));
        CU(s, cudaStreamWaitEvent(s->stream2, s->ev_ready, 0));
        PlaneRange pr;
        pr.st = s->stream2;
        if ((rc = step_launch(s, save_prev, pr))) return rc;
        CU(s, cudaEventRecord(s->ev_interior, s->stream2));
        CU(s, cudaStreamWaitEvent(s->stream, s->ev_interior, 0));
    } else if ((rc = step_launch(s, save_prev, PlaneRange()))) {
        return rc;
    }
    step_mark(s);
    return 0;
}
// stage B: cell-list rebuild (scan + stable counting-sort scatter of the whole state)
int step_finish(crb_sim* s, bool save_prev) {
    const u64 n_new = s->n_src - s->n_gone;
    if (n_new > s->cap) return fail(s, CRB_ERR_CAPACITY, "more agents than capacity after migration");
    int rc;
    if ((rc = sort_into_place(s, s->n_src, n_new, save_prev, false))) return rc;
    s->n = n_new;
    s->iteration++;
    s->noise_cursor++;
    return 0;
}
}  // namespace

// Launch-bound populations (the reference's own examples hold 10^3..10^5 agents: one step is seven launches
// of a few microseconds each) run batches of steps as CUDA graphs.  One graph = TWO consecutive steps: the
// sorted/target state buffers swap after every step, so the pointer set repeats with period two.  Eligible:
// one device, point agents, no per-step host input (the noise-table cursor and the Philox iteration counter
// travel by value in the kernel parameters => deterministic mechanics or CRB_NOISE_NONE only), timers off.
// The last step of a batch always runs outside the graph (it alone writes the older ring-buffer entry,
// save_prev).  Same kernels in the same order => bit-identical to the plain loop (tested).
static bool graph_eligible(const crb_sim* s) {
    if (s->graphs_off || s->timers_on || s->world > 1 || s->rods || !s->n) return false;
    const bool stochastic = s->cfg.noise_mode != CRB_NOISE_NONE && s->cfg.mechanics_kind != CRB_MECH_NEWTON_DAMPED;
    return !stochastic;
}

static int plain_step(crb_sim* s, bool last) {
    int rc;
    if ((rc = step_compute(s, last))) return rc;
    return step_finish(s, last);
}

static int graph_for_orientation(crb_sim* s, cudaGraphExec_t* out) {
    const void* key = (const void*)s->X.pos[0];
    for (int k = 0; k < 2; ++k)
        if (s->step_graph[k] && s->step_graph_key[k] == key) {
            *out = s->step_graph[k];
            return 0;
        }
    const int slot = s->step_graph[0] ? 1 : 0;
    if (s->step_graph[slot]) {
        cudaGraphExecDestroy(s->step_graph[slot]);
        s->step_graph[slot] = nullptr;
    }
    const u64 it0 = s->iteration, nc0 = s->noise_cursor, l0 = s->launches;
    CU(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed));
    int rc = plain_step(s, false);
    if (!rc) rc = plain_step(s, false);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s->stream, &g);
    // nothing ran: the host-side bookkeeping of the two captured steps is rolled back (the buffer
    // orientation is back where it was after two swaps)
    s->step_graph_launches = s->launches - l0;
    s->iteration = it0;
    s->noise_cursor = nc0;
    s->launches = l0;
    if (rc || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return rc ? rc : fail(s, CRB_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
    }
    cudaGraphExec_t ge = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) return fail(s, CRB_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ie));
    s->step_graph[slot] = ge;
    s->step_graph_key[slot] = key;
    *out = ge;
    return 0;
}

static int step_impl(crb_sim* s, uint64_t n_steps, bool async);

int crb_step(crb_sim* s, uint64_t n_steps) { return step_impl(s, n_steps, false); }
int crb_step_async(crb_sim* s, uint64_t n_steps) { return step_impl(s, n_steps, true); }

static int step_impl(crb_sim* s, uint64_t n_steps, bool async) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (async && !s->kernels_only) {
        s->kernels_only = true;
        graphs_invalidate(s);  // (a captured step holds a memset node)
    }
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "crb_step before crb_set_agents");
    if (s->world > 1)
        return fail(s, CRB_ERR_STATE, "slab handle: drive the step with crb_slab_* (halo/migration exchange)");
    CU(s, cudaSetDevice((int)s->cfg.device));
    s->aux_settable = false;
    if (s->rods) return rods_step(s, n_steps);  // (rods read their division counts back: always synchronous)
    u64 it = 0;
    int rc;
    if (n_steps >= 4 && graph_eligible(s)) {
        const u64 pairs = (n_steps - 1) / 2;  // the last step stays outside
        cudaGraphExec_t ge = nullptr;
        if ((rc = graph_for_orientation(s, &ge))) return rc;
        for (u64 k = 0; k < pairs; ++k) CU(s, cudaGraphLaunch(ge, s->stream));
        s->iteration += 2 * pairs;
        s->noise_cursor += 2 * pairs;
        s->launches += pairs * s->step_graph_launches;
        step_mark(s);
        it = 2 * pairs;
    }
    for (; it < n_steps; ++it)
        if ((rc = plain_step(s, it + 1 == n_steps))) return rc;
    return async ? 0 : check_latch(s);
}

// ---- slab protocol (one handle per device; the caller moves the buffers between devices) -----------
uint64_t crb_slab_halo_bytes(const crb_sim* s) { return s ? s->hl.bytes : 0; }
uint64_t crb_slab_migration_bytes(const crb_sim* s) { return s ? s->ml.bytes : 0; }

int crb_slab_info(const crb_sim* s, uint64_t out[8]) {
    if (!s || !out) return CRB_ERR_INVALID_ARG;
    out[0] = (u64)s->g.p0 + s->g.own_lo;              // first owned global plane
    out[1] = (u64)(s->g.own_hi - s->g.own_lo);        // owned planes
    out[2] = s->has_lo;
    out[3] = s->has_hi;
    out[4] = s->G;
    out[5] = s->mcap;
    out[6] = s->cap;
    out[7] = (u64)s->g.ng[1] * s->g.ng[2];
    return 0;
}

int crb_slab_halo_export(crb_sim* s, void* lo_buf, void* hi_buf) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "no agents");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods) return rods_slab_halo_export(s, lo_buf, hi_buf);
    const u64 plane = s->hl.plane;
    const u32 threads = 256;
    const u32 grid = blocks_for(std::max<u64>(s->hl.gcap, plane + 1), threads);
    HaloLayout L = s->hl;
    if (s->uniform) L.n_ip = 0;
    L.mode = param_mode(s);
    stage_begin(s, 3);
    int nl = 0;
    if (s->has_lo && lo_buf) {
        k_halo_export<<<grid, threads, 0, s->stream>>>(s->X, s->start, (u32)(s->g.own_lo * plane), L,
                                                      (unsigned char*)lo_buf, s->err);
        ++nl;
    }
    if (s->has_hi && hi_buf) {
        k_halo_export<<<grid, threads, 0, s->stream>>>(s->X, s->start, (u32)((s->g.own_hi - 1) * plane), L,
                                                      (unsigned char*)hi_buf, s->err);
        ++nl;
    }
    stage_end(s, 3, nl);
    CU(s, cudaGetLastError());
    return 0;
}

int crb_slab_halo_import(crb_sim* s, const void* lo_buf, const void* hi_buf) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "no agents");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods) return rods_slab_halo_import(s, lo_buf, hi_buf);
    const u64 plane = s->hl.plane;
    const u32 threads = 256;
    const u32 grid = blocks_for(std::max<u64>(s->hl.gcap, plane + 1), threads);
    HaloLayout L = s->hl;
    if (s->uniform) L.n_ip = 0;
    L.mode = param_mode(s);
    stage_begin(s, 3);
    int nl = 0;
    if (s->has_lo) {
        if (!lo_buf) return fail(s, CRB_ERR_INVALID_ARG, "lower halo buffer missing");
        k_halo_import<<<grid, threads, 0, s->stream>>>(s->X, s->start, 0u, s->G, 1, L,
                                                      (const unsigned char*)lo_buf, s->err);
        ++nl;
    }
    if (s->has_hi) {
        if (!hi_buf) return fail(s, CRB_ERR_INVALID_ARG, "upper halo buffer missing");
        k_halo_import<<<grid, threads, 0, s->stream>>>(s->X, s->start, (u32)((s->g.L0 - 1) * plane),
                                                      s->G + s->n, 0, L, (const unsigned char*)hi_buf,
                                                      s->err);
        ++nl;
    }
    stage_end(s, 3, nl);
    CU(s, cudaGetLastError());
    return 0;
}

int crb_slab_compute(crb_sim* s, int save_prev) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "no agents");
    if (s->step_open) return fail(s, CRB_ERR_STATE, "crb_slab_finish pending");
    if (s->interior_state) return fail(s, CRB_ERR_STATE, "crb_slab_compute_boundary pending");
    CU(s, cudaSetDevice((int)s->cfg.device));
    s->aux_settable = false;
    int rc;
    if (s->rods) {
        if ((rc = rods_compute(s, save_prev != 0))) return rc;
        s->step_open = true;
        return 0;
    }
    if ((rc = step_compute(s, save_prev != 0))) return rc;
    s->step_open = true;
    return 0;
}

// planes whose 3^D neighbourhood touches no ghost plane: [own_lo + has_lo, own_hi - has_hi)
static void interior_planes(const crb_sim* s, int* b, int* e) {
    *b = s->g.own_lo + (s->has_lo ? 1 : 0);
    *e = s->g.own_hi - (s->has_hi ? 1 : 0);
    if (*e < *b) *e = *b;
}

int crb_slab_compute_interior(crb_sim* s, int save_prev) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "no agents");
    if (s->step_open) return fail(s, CRB_ERR_STATE, "crb_slab_finish pending");
    if (s->interior_state) return fail(s, CRB_ERR_STATE, "crb_slab_compute_boundary pending");
    if (s->rods)
        return fail(s, CRB_ERR_STATE, "rods: the force pass is not split into interior / boundary planes; use crb_slab_compute");
    CU(s, cudaSetDevice((int)s->cfg.device));
    s->aux_settable = false;
    int rc;
    if ((rc = step_prepare(s))) return rc;
    s->interior_state = save_prev ? 3 : 1;
    int b, e;
    interior_planes(s, &b, &e);
    // the per-agent kernel (non-uniform parameters) is not split: everything runs in the boundary call
    if (!(s->tile_ok || s->tile2_ok || s->tile3_ok) || e <= b || !s->n) return 0;
    // the interior pass starts once the sorted state X of the previous step and the cleared histogram exist
    CU(s, cudaEventRecord(s->ev_ready, s->stream));
    CU(s, cudaStreamWaitEvent(s->stream2, s->ev_ready, 0));
    PlaneRange pr;
    pr.x_begin = b;
    pr.x_count = e - b;
    pr.st = s->stream2;
    if ((rc = step_launch(s, save_prev != 0, pr))) return rc;
    CU(s, cudaEventRecord(s->ev_interior, s->stream2));
    s->interior_state |= 4;  // something is running on stream2
    return 0;
}

int crb_slab_compute_boundary(crb_sim* s) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->interior_state) return fail(s, CRB_ERR_STATE, "crb_slab_compute_interior first");
    CU(s, cudaSetDevice((int)s->cfg.device));
    const bool save_prev = (s->interior_state & 2) != 0;
    const bool split = (s->interior_state & 4) != 0;
    s->interior_state = 0;
    int rc;
    if (!split) {
        if ((rc = step_launch(s, save_prev, PlaneRange()))) return rc;
    } else {
        int b, e;
        interior_planes(s, &b, &e);
        PlaneRange pr;
        pr.st = s->stream;
        if (b > s->g.own_lo) {
            pr.x_begin = s->g.own_lo;
            pr.x_count = b - s->g.own_lo;
            if ((rc = step_launch(s, save_prev, pr))) return rc;
        }
        if (s->g.own_hi > e) {
            pr.x_begin = e;
            pr.x_count = s->g.own_hi - e;
            if ((rc = step_launch(s, save_prev, pr))) return rc;
        }
        CU(s, cudaStreamWaitEvent(s->stream, s->ev_interior, 0));
    }
    step_mark(s);
    s->step_open = true;
    return 0;
}

int crb_slab_migration_export(crb_sim* s, void* lo_buf, void* hi_buf, uint64_t n_out[2]) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->step_open) return fail(s, CRB_ERR_STATE, "crb_slab_compute first");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (n_out) n_out[0] = n_out[1] = 0;
    s->mig_out[0] = s->mig_out[1] = 0;
    s->mig_counts_pending = false;
    if (s->world == 1) return 0;
    if (s->rods) return rods_slab_migration_export(s, lo_buf, hi_buf, n_out);
    const u64 plane = s->hl.plane;
    CU(s, cudaMemsetAsync(s->scratch_u64, 0, 2 * sizeof(u64), s->stream));
    stage_begin(s, 4);
    int nl = 0;
    if (s->n) {
        MigExportParams M;
        std::memset(&M, 0, sizeof(M));
        M.A = s->X;
        M.B = s->T;
        M.newkey = s->newkey;
        M.sortkey = s->sortkey;
        M.first = s->G;
        M.n = s->n;
        M.lo_end = s->has_lo ? (u32)plane : 0u;
        M.hi_begin = s->has_hi ? (u32)((s->g.L0 - 1) * plane) : 0xffffffffu;
        M.vox_base = (u64)s->g.p0 * plane;
        M.L = s->ml;
        M.L.mode = param_mode(s);
        M.buf[0] = (unsigned char*)lo_buf;
        M.buf[1] = (unsigned char*)hi_buf;
        M.counter = s->scratch_u64;
        M.err = s->err;
        M.uniform = s->uniform ? 1 : 0;
        if ((s->has_lo && !lo_buf) || (s->has_hi && !hi_buf))
            return fail(s, CRB_ERR_INVALID_ARG, "migration buffer missing");
        k_mig_export<<<blocks_for(s->n, 256), 256, 0, s->stream>>>(M);
        ++nl;
    }
    // the leavers were counted in the histogram of the ghost planes: they take no slot here
    if (s->has_lo) CU(s, cudaMemsetAsync(s->count_new, 0, plane * sizeof(u32), s->stream));
    if (s->has_hi)
        CU(s, cudaMemsetAsync(s->count_new + (s->g.L0 - 1) * plane, 0, plane * sizeof(u32), s->stream));
    // headers
    if (s->has_lo) CU(s, cudaMemcpyAsync(lo_buf, s->scratch_u64, sizeof(u64), cudaMemcpyDeviceToDevice, s->stream));
    if (s->has_hi) CU(s, cudaMemcpyAsync(hi_buf, s->scratch_u64 + 1, sizeof(u64), cudaMemcpyDeviceToDevice, s->stream));
    stage_end(s, 4, nl);
    CU(s, cudaGetLastError());
    if (!n_out) {
        // asynchronous form: the counts are read back by crb_slab_migration_import together with the
        // arrival counts (ONE host synchronisation per step instead of two)
        s->mig_counts_pending = true;
        return 0;
    }
    u64 cnt[2] = {0, 0};
    CU(s, cudaMemcpyAsync(cnt, s->scratch_u64, 2 * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    if (cnt[0] > s->mcap || cnt[1] > s->mcap)
        return fail(s, CRB_ERR_CAPACITY, "more migrating agents than migration_capacity");
    n_out[0] = cnt[0];
    n_out[1] = cnt[1];
    s->mig_out[0] = cnt[0];
    s->mig_out[1] = cnt[1];
    s->n_gone = cnt[0] + cnt[1];
    return 0;
}

int crb_slab_migration_counts(const crb_sim* s, uint64_t n_out[2]) {
    if (!s || !n_out) return CRB_ERR_INVALID_ARG;
    n_out[0] = s->mig_out[0];
    n_out[1] = s->mig_out[1];
    return 0;
}

int crb_slab_migration_import(crb_sim* s, const void* lo_buf, const void* hi_buf, uint64_t n_in[2]) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->step_open) return fail(s, CRB_ERR_STATE, "crb_slab_compute first");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods) return rods_slab_migration_import(s, lo_buf, hi_buf, n_in);
    u64 cnt[2] = {0, 0};
    u64 outc[2] = {0, 0};
    if (s->has_lo) {
        if (!lo_buf) return fail(s, CRB_ERR_INVALID_ARG, "lower migration buffer missing");
        CU(s, cudaMemcpyAsync(&cnt[0], lo_buf, sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    }
    if (s->has_hi) {
        if (!hi_buf) return fail(s, CRB_ERR_INVALID_ARG, "upper migration buffer missing");
        CU(s, cudaMemcpyAsync(&cnt[1], hi_buf, sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    }
    if (s->mig_counts_pending)
        CU(s, cudaMemcpyAsync(outc, s->scratch_u64, 2 * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    if (s->mig_counts_pending) {
        s->mig_counts_pending = false;
        if (outc[0] > s->mcap || outc[1] > s->mcap)
            return fail(s, CRB_ERR_CAPACITY, "more migrating agents than migration_capacity");
        s->mig_out[0] = outc[0];
        s->mig_out[1] = outc[1];
        s->n_gone = outc[0] + outc[1];
    }
    if (cnt[0] > s->mcap || cnt[1] > s->mcap)
        return fail(s, CRB_ERR_CAPACITY, "more arriving agents than migration_capacity");
    if (s->n_src + cnt[0] + cnt[1] > s->cap + 2 * s->mcap)
        return fail(s, CRB_ERR_CAPACITY, "more agents than capacity");
    const u64 plane = s->hl.plane;
    stage_begin(s, 4);
    int nl = 0;
    for (int side = 0; side < 2; ++side) {
        if (!cnt[side]) continue;
        MigImportParams M;
        std::memset(&M, 0, sizeof(M));
        M.A = s->X;
        M.B = s->T;
        M.newkey = s->newkey;
        M.rank_arb = s->rank_arb;
        M.count_new = s->count_new;
        M.sortkey = s->sortkey;
        M.dst = s->G + s->n_src;
        M.n = cnt[side];
        M.vox_base = (u64)s->g.p0 * plane;
        M.nvox = s->nvox;
        M.L = s->ml;
        M.L.mode = param_mode(s);
        M.buf = (const unsigned char*)(side == 0 ? lo_buf : hi_buf);
        M.err = s->err;
        M.uniform = s->uniform ? 1 : 0;
        k_mig_import<<<blocks_for(cnt[side], 256), 256, 0, s->stream>>>(M);
        ++nl;
        s->n_src += cnt[side];
    }
    stage_end(s, 4, nl);
    CU(s, cudaGetLastError());
    if (n_in) {
        n_in[0] = cnt[0];
        n_in[1] = cnt[1];
    }
    return 0;
}

int crb_slab_finish(crb_sim* s, int save_prev) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->step_open) return fail(s, CRB_ERR_STATE, "crb_slab_compute first");
    CU(s, cudaSetDevice((int)s->cfg.device));
    s->step_open = false;
    int rc;
    if (s->rods) return rods_finish(s, save_prev != 0);
    if ((rc = step_finish(s, save_prev != 0))) return rc;
    plan_tile(s);
    return 0;
}

// ---- crb_slab_step: the whole slab step behind ONE call, exchange over peer memory, no host synchronisation ------
// arena layout (the same on every rank: the buffer sizes agree): 4 flags at byte 0 (halo from lo, halo from hi,
// migration from lo, migration from hi), then halo_in[side][parity], then mig_in[side][parity]
namespace {
unsigned char* arena_halo_in(const crb_sim* s, unsigned char* base, int side, int parity) {
    return base + 256 + (u64)(side * 2 + parity) * s->arena_halo;
}
unsigned char* arena_mig_in(const crb_sim* s, unsigned char* base, int side, int parity) {
    return base + 256 + 4 * s->arena_halo + (u64)(side * 2 + parity) * s->arena_mig;
}
u64* arena_flag(unsigned char* base, int which) { return reinterpret_cast<u64*>(base) + which; }
// a dead neighbour: 20 s (CRB_SLAB_TIMEOUT_MS), then CRB_ERR_TIMEOUT
unsigned long long slab_timeout_ns() {
    static const unsigned long long ns = [] {
        const char* e = std::getenv("CRB_SLAB_TIMEOUT_MS");
        return (e ? std::strtoull(e, nullptr, 10) : 20000ull) * 1000000ull;
    }();
    return ns;
}
}  // namespace

uint64_t crb_slab_arena_bytes(const crb_sim* s) { return s ? s->arena_bytes : 0; }
void* crb_slab_arena(crb_sim* s) { return s ? (void*)s->arena : nullptr; }

int crb_slab_ipc_handle(crb_sim* s, void* handle64) {
    if (!s || !handle64) return CRB_ERR_INVALID_ARG;
    if (!s->arena) return fail(s, CRB_ERR_STATE, "no exchange arena (world_size == 1 or rods)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU(s, cudaSetDevice((int)s->cfg.device));
    cudaIpcMemHandle_t h;
    CU(s, cudaIpcGetMemHandle(&h, s->arena));
    std::memcpy(handle64, &h, 64);
    return 0;
}

static int connect_common(crb_sim* s) {
    if (!s->arena) return fail(s, CRB_ERR_STATE, "no exchange arena (world_size == 1 or rods)");
    CU(s, cudaSetDevice((int)s->cfg.device));
    for (int k = 0; k < 2; ++k) {
        if (s->peer_arena[k] && s->peer_ipc[k]) cudaIpcCloseMemHandle(s->peer_arena[k]);
        s->peer_arena[k] = nullptr;
        s->peer_ipc[k] = false;
    }
    // flags restart with the connection (every rank connects before anyone steps: host-side barrier of the caller)
    CU(s, cudaMemsetAsync(s->arena, 0, 256, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    s->slab_steps = 0;
    return 0;
}

int crb_slab_connect(crb_sim* s, void* lo_arena, void* hi_arena) {
    if (!s) return CRB_ERR_INVALID_ARG;
    int rc;
    if ((rc = connect_common(s))) return rc;
    if ((s->has_lo && !lo_arena) || (s->has_hi && !hi_arena))
        return fail(s, CRB_ERR_INVALID_ARG, "crb_slab_connect: a neighbour's arena is missing");
    void* in[2] = {s->has_lo ? lo_arena : nullptr, s->has_hi ? hi_arena : nullptr};
    for (int k = 0; k < 2; ++k) {
        if (!in[k]) continue;
        cudaPointerAttributes at;
        CU(s, cudaPointerGetAttributes(&at, in[k]));
        if (at.type != cudaMemoryTypeDevice) return fail(s, CRB_ERR_INVALID_ARG, "crb_slab_connect: not a device pointer");
        if (at.device != (int)s->cfg.device) {
            int can = 0;
            CU(s, cudaDeviceCanAccessPeer(&can, (int)s->cfg.device, at.device));
            if (!can) return fail(s, CRB_ERR_CUDA, "crb_slab_connect: no peer access to the neighbour's device");
            const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(s, CRB_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        s->peer_arena[k] = (unsigned char*)in[k];
    }
    return 0;
}

int crb_slab_connect_ipc(crb_sim* s, const void* lo_handle64, const void* hi_handle64) {
    if (!s) return CRB_ERR_INVALID_ARG;
    int rc;
    if ((rc = connect_common(s))) return rc;
    if ((s->has_lo && !lo_handle64) || (s->has_hi && !hi_handle64))
        return fail(s, CRB_ERR_INVALID_ARG, "crb_slab_connect_ipc: a neighbour's handle is missing");
    const void* in[2] = {s->has_lo ? lo_handle64 : nullptr, s->has_hi ? hi_handle64 : nullptr};
    for (int k = 0; k < 2; ++k) {
        if (!in[k]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, in[k], 64);
        void* p = nullptr;
        CU(s, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_arena[k] = (unsigned char*)p;
        s->peer_ipc[k] = true;
    }
    return 0;
}

static int slab_step_impl(crb_sim* s, uint64_t n_steps, bool async);
// the host-side copy of the owned-agent count is refreshed from the device after asynchronous slab steps
static int slab_finalize(crb_sim* s) {
    if (!s->slab_async_pending) return 0;
    u64 n_now = 0;
    CU(s, cudaMemcpyAsync(&n_now, s->cnt, sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    s->slab_async_pending = false;
    s->last_step_n = 0;
    s->n = std::min<u64>(n_now, s->cap);
    s->n_src = s->n;
    s->n_gone = 0;
    plan_tile(s);
    return 0;
}

int crb_slab_step(crb_sim* s, uint64_t n_steps) { return slab_step_impl(s, n_steps, false); }
int crb_slab_step_async(crb_sim* s, uint64_t n_steps) { return slab_step_impl(s, n_steps, true); }

static int slab_step_impl(crb_sim* s, uint64_t n_steps, bool async) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "no agents");
    if (s->rods)
        return fail(s, CRB_ERR_STATE, "rods: crb_slab_step is not available (division counts are read back every step); "
                                      "use the six-phase protocol crb_slab_halo_export .. crb_slab_finish");
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, "a crb_slab_compute .. crb_slab_finish step is open");
    if (s->world > 1 && ((s->has_lo && !s->peer_arena[0]) || (s->has_hi && !s->peer_arena[1])))
        return fail(s, CRB_ERR_STATE, "crb_slab_step: connect the neighbours first (crb_slab_connect / crb_slab_connect_ipc)");
    if (!n_steps) return 0;
    CU(s, cudaSetDevice((int)s->cfg.device));
    s->aux_settable = false;
    const u64 plane = s->hl.plane;
    const u32 threads = 256;
    const u32 hgrid = blocks_for(std::max<u64>(s->hl.gcap, plane + 1), threads);
    const u64 src_cap = s->cap + 2 * s->mcap;
    int rc;
    if ((rc = slab_preload(s))) return rc;
    // the host knows the owned count here unless asynchronous steps are still queued; from now on it lives on the
    // device
    if (!s->slab_async_pending) {
        k_set_u64<<<1, 32, 0, s->stream>>>(s->cnt, s->n);
        s->launches++;
    }
    for (u64 it = 0; it < n_steps; ++it) {
        const bool last = it + 1 == n_steps;
        const u64 step_no = s->slab_steps;
        const int par = (int)(step_no & 1);
        if ((rc = step_prepare(s, true))) return rc;
        // ---- interior planes on the second stream (needs no halo) ---------------------------------------------
        int ib, ie;
        interior_planes(s, &ib, &ie);
        const bool split = (s->tile_ok || s->tile2_ok || s->tile3_ok) && ie > ib;
        if (split) {
            CU(s, cudaEventRecord(s->ev_ready, s->stream));
            CU(s, cudaStreamWaitEvent(s->stream2, s->ev_ready, 0));
            PlaneRange pr;
            pr.x_begin = ib;
            pr.x_count = ie - ib;
            pr.st = s->stream2;
            if ((rc = step_launch(s, last, pr))) return rc;
            CU(s, cudaEventRecord(s->ev_interior, s->stream2));
        }
        // ---- halo: boundary planes straight into the neighbours' import buffers, then their flags ----------------
        HaloLayout L = s->hl;
        if (s->uniform) L.n_ip = 0;
        L.mode = param_mode(s);
        stage_begin(s, 3);
        int nl = 0;
        if (s->has_lo) {  // my lowest owned plane is the neighbour's upper ghost plane: its side "hi"
            k_halo_export<<<hgrid, threads, 0, s->stream>>>(s->X, s->start, (u32)(s->g.own_lo * plane), L,
                                                           arena_halo_in(s, s->peer_arena[0], 1, par), s->err);
            ++nl;
        }
        if (s->has_hi) {
            k_halo_export<<<hgrid, threads, 0, s->stream>>>(s->X, s->start, (u32)((s->g.own_hi - 1) * plane), L,
                                                           arena_halo_in(s, s->peer_arena[1], 0, par), s->err);
            ++nl;
        }
        if (s->has_lo || s->has_hi) {
            k_slab_publish<<<1, 32, 0, s->stream>>>(nullptr, nullptr, nullptr,
                                                   s->has_lo ? arena_flag(s->peer_arena[0], 1) : nullptr,
                                                   s->has_hi ? arena_flag(s->peer_arena[1], 0) : nullptr, step_no + 1, nullptr);
            k_slab_wait<<<1, 32, 0, s->stream>>>(s->has_lo ? arena_flag(s->arena, 0) : nullptr,
                                                s->has_hi ? arena_flag(s->arena, 1) : nullptr, step_no + 1,
                                                slab_timeout_ns(), s->err);
            nl += 2;
        }
        if (s->has_lo) {
            k_halo_import<<<hgrid, threads, 0, s->stream>>>(s->X, s->start, 0u, s->G, 1, L,
                                                           arena_halo_in(s, s->arena, 0, par), s->err, nullptr);
            ++nl;
        }
        if (s->has_hi) {
            k_halo_import<<<hgrid, threads, 0, s->stream>>>(s->X, s->start, (u32)((s->g.L0 - 1) * plane), s->G, 0, L,
                                                           arena_halo_in(s, s->arena, 1, par), s->err, s->cnt);
            ++nl;
        }
        stage_end(s, 3, nl);
        // ---- boundary planes (or everything when the force pass is not split), then join the interior pass --------
        if (!split) {
            KParams P = make_params(s);
            (void)P;
            if ((rc = step_launch_dev(s, last))) return rc;
        } else {
            PlaneRange pr;
            pr.st = s->stream;
            if (ib > s->g.own_lo) {
                pr.x_begin = s->g.own_lo;
                pr.x_count = ib - s->g.own_lo;
                if ((rc = step_launch(s, last, pr))) return rc;
            }
            if (s->g.own_hi > ie) {
                pr.x_begin = ie;
                pr.x_count = s->g.own_hi - ie;
                if ((rc = step_launch(s, last, pr))) return rc;
            }
            CU(s, cudaStreamWaitEvent(s->stream, s->ev_interior, 0));
        }
        // ---- migration: leavers straight into the neighbours' buffers, counts + flags, arrivals ---------------------
        stage_begin(s, 4);
        nl = 0;
        if (s->has_lo || s->has_hi) {
            k_zero_u32<<<1, 32, 0, s->stream>>>(nullptr, 0, nullptr, 0, s->scratch_u64, 2);
            MigExportParams M;
            std::memset(&M, 0, sizeof(M));
            M.A = s->X;
            M.B = s->T;
            M.newkey = s->newkey;
            M.sortkey = s->sortkey;
            M.first = s->G;
            M.n = s->cap;
            M.n_dev = s->cnt;
            M.lo_end = s->has_lo ? (u32)plane : 0u;
            M.hi_begin = s->has_hi ? (u32)((s->g.L0 - 1) * plane) : 0xffffffffu;
            M.vox_base = (u64)s->g.p0 * plane;
            M.L = s->ml;
            M.L.mode = param_mode(s);
            M.buf[0] = s->has_lo ? arena_mig_in(s, s->peer_arena[0], 1, par) : nullptr;
            M.buf[1] = s->has_hi ? arena_mig_in(s, s->peer_arena[1], 0, par) : nullptr;
            M.counter = s->scratch_u64;
            M.err = s->err;
            M.uniform = s->uniform ? 1 : 0;
            k_mig_export<<<blocks_for(s->cap, 256), 256, 0, s->stream>>>(M);
            // the leavers were counted in the histogram of the ghost planes: they take no slot here
            k_zero_u32<<<blocks_for(plane, 256), 256, 0, s->stream>>>(
                s->has_lo ? s->count_new : nullptr, s->has_lo ? plane : 0,
                s->has_hi ? s->count_new + (s->g.L0 - 1) * plane : nullptr, s->has_hi ? plane : 0, nullptr, 0);
            k_slab_publish<<<1, 32, 0, s->stream>>>(s->scratch_u64, M.buf[0], M.buf[1],
                                                   s->has_lo ? arena_flag(s->peer_arena[0], 3) : nullptr,
                                                   s->has_hi ? arena_flag(s->peer_arena[1], 2) : nullptr, step_no + 1,
                                                   s->cnt + 4);
            k_slab_wait<<<1, 32, 0, s->stream>>>(s->has_lo ? arena_flag(s->arena, 2) : nullptr,
                                                s->has_hi ? arena_flag(s->arena, 3) : nullptr, step_no + 1,
                                                slab_timeout_ns(), s->err);
            nl += 5;
        }
        {
            MigImportDevParams Q;
            std::memset(&Q, 0, sizeof(Q));
            Q.M.A = s->X;
            Q.M.B = s->T;
            Q.M.newkey = s->newkey;
            Q.M.rank_arb = s->rank_arb;
            Q.M.count_new = s->count_new;
            Q.M.sortkey = s->sortkey;
            Q.M.vox_base = (u64)s->g.p0 * plane;
            Q.M.nvox = s->nvox;
            Q.M.L = s->ml;
            Q.M.L.mode = param_mode(s);
            Q.M.err = s->err;
            Q.M.uniform = s->uniform ? 1 : 0;
            Q.buf[0] = s->has_lo ? arena_mig_in(s, s->arena, 0, par) : nullptr;
            Q.buf[1] = s->has_hi ? arena_mig_in(s, s->arena, 1, par) : nullptr;
            Q.first = s->G;
            Q.src_cap = src_cap;
            Q.cnt = s->cnt;
            const dim3 grid(blocks_for(std::max<u64>(s->mcap, 1), 256), 2);
            k_mig_import_dev<<<grid, 256, 0, s->stream>>>(Q);
            ++nl;
        }
        stage_end(s, 4, nl);
        CU(s, cudaGetLastError());
        // ---- cell-list rebuild with device-side counts -----------------------------------------------------------------
        if ((rc = sort_into_place(s, src_cap, s->cap, last, false, true))) return rc;
        // (sort_into_place swapped start / start_alt: the totals of this step's scan are in s->start now)
        k_slab_set_count<<<1, 32, 0, s->stream>>>(s->start, s->nvox, s->G, s->cap, s->cnt, s->err);
        s->launches++;
        s->iteration++;
        s->noise_cursor++;
        s->slab_steps++;
    }
    s->slab_async_pending = true;
    if (async) return 0;
    // ---- end of the batch: the only synchronisation -------------------------------------------------------------------
    if ((rc = slab_finalize(s))) return rc;
    return check_latch(s);
}

// ---- storage snapshots (save_cells at PartialSave iterations, chili/datastructures.rs:373-396) -----
void* crb_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void crb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int crb_snapshot_wait(crb_sim* s, uint64_t* n_agents) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->snap_pending) return fail(s, CRB_ERR_STATE, "no snapshot in flight");
    CU(s, cudaSetDevice((int)s->cfg.device));
    CU(s, cudaEventSynchronize(s->ev_snap_done));
    s->snap_pending = false;
    const u64 n = s->snap_n;
    const int D = s->D;
    // ring buffers oldest-first, unused entries zero (as crb_get_aux)
    for (double* o : {s->snap_hx_host, s->snap_hv_host}) {
        if (!o) continue;
        for (u64 i = 0; i < n; ++i) {
            double* r = o + i * 2 * D;
            const u8 hl = s->snap_hlen_host[i];
            if (hl == 1) {
                for (int d = 0; d < D; ++d) {
                    r[d] = r[D + d];
                    r[D + d] = 0.0;
                }
            } else if (hl == 0) {
                for (int k = 0; k < 2 * D; ++k) r[k] = 0.0;
            }
        }
    }
    if (s->snap_hl_out)
        for (u64 i = 0; i < n; ++i) s->snap_hl_out[i] = s->snap_hlen_host[i];
    if (n_agents) *n_agents = n;
    return 0;
}

int crb_snapshot_async(crb_sim* s, uint64_t cap, double* pos, double* vel, uint64_t* ids, uint64_t* voxel,
                       double* hist_dx, double* hist_dv, uint32_t* hist_len) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "no agents");
    if (s->rods && (hist_dx || hist_dv || hist_len))
        return fail(s, CRB_ERR_INVALID_ARG, "rods: the ring buffers are read with crb_get_aux");
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, "snapshot inside an open slab step");
    if (cap < s->n) return fail(s, CRB_ERR_CAPACITY, "output buffers too small");
    CU(s, cudaSetDevice((int)s->cfg.device));
    int rc;
    if (s->snap_pending && (rc = crb_snapshot_wait(s, nullptr))) return rc;
    if (!s->stream_io) {
        CU(s, cudaStreamCreateWithFlags(&s->stream_io, cudaStreamNonBlocking));
        CU(s, cudaEventCreateWithFlags(&s->ev_snap_packed, cudaEventDisableTiming));
        CU(s, cudaEventCreateWithFlags(&s->ev_snap_done, cudaEventDisableTiming));
    }
    const u64 n = s->n;
    const int D = s->rods ? 3 * s->V : s->D;  // doubles per position row (rods: V vertices x 3)
    // doubles per agent: pos, vel, id, voxel, hx, hv, hlen (padded); a rod snapshot carries no ring buffers
    const u64 per = s->rods ? (u64)(2 * D + 2) : (u64)(2 * D + 2 + 4 * D + 1);
    if (s->snap_cap < s->cap) {
        if (s->snap) cudaFree(s->snap);
        if (s->snap_hlen_host) cudaFreeHost(s->snap_hlen_host);
        s->snap = nullptr;
        s->snap_hlen_host = nullptr;
        s->snap_cap = 0;
        CU(s, cudaMalloc((void**)&s->snap, s->cap * per * sizeof(double)));
        CU(s, cudaMallocHost((void**)&s->snap_hlen_host, s->cap));
        s->snap_cap = s->cap;
    }
    double* b_pos = s->snap;
    double* b_vel = b_pos + n * D;
    u64* b_id = (u64*)(b_vel + n * D);
    u64* b_vox = b_id + n;
    double* b_hx = (double*)(b_vox + n);
    double* b_hv = b_hx + n * 2 * D;
    u8* b_hl = (u8*)(b_hv + n * 2 * D);
    auto pack = [&](double* dst, int width, const double* const* cols) {
        const double* c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < width; ++k) c[k] = cols[k] ? cols[k] + s->G : nullptr;
        k_pack_rows<<<blocks_for(n, 256), 256, 0, s->stream>>>(dst, width, c[0], c[1], c[2], c[3], c[4], c[5],
                                                              c[6], c[7], n);
        s->launches++;
    };
    if (s->rods) {
        if ((rc = rods_snapshot_pack(s, b_pos, b_vel, b_id, b_vox, pos, vel, ids, voxel))) return rc;
    } else if (n) {
        if (pos) pack(b_pos, D, s->X.pos);
        if (vel) {
            if (s->cfg.mechanics_kind == CRB_MECH_BROWNIAN)
                CU(s, cudaMemsetAsync(b_vel, 0, n * D * sizeof(double), s->stream));
            else
                pack(b_vel, D, s->X.vel);
        }
        if (ids) {
            k_copy_u64<<<blocks_for(n, 256), 256, 0, s->stream>>>(b_id, s->X.id + s->G, n);
            s->launches++;
        }
        if (voxel) {
            const u64 base = (u64)s->g.p0 * s->g.ng[1] * s->g.ng[2];
            k_global_keys<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->key + s->G, b_vox, base, n);
            s->launches++;
        }
        for (int which = 0; which < 2; ++which) {
            if (!(which == 0 ? hist_dx : hist_dv)) continue;
            const double* cols[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int d = 0; d < D; ++d) {
                cols[d] = which == 0 ? s->X.h0x[d] : s->X.h0v[d];
                cols[D + d] = which == 0 ? s->X.h1x[d] : s->X.h1v[d];
            }
            pack(which == 0 ? b_hx : b_hv, 2 * D, cols);
        }
        if (hist_len || hist_dx || hist_dv) {  // (crb_snapshot_wait orders the ring entries by the fill count)
            k_copy_u8<<<blocks_for(n, 256), 256, 0, s->stream>>>(b_hl, s->X.hlen + s->G, n);
            s->launches++;
        }
    }
    CU(s, cudaGetLastError());
    CU(s, cudaEventRecord(s->ev_snap_packed, s->stream));
    CU(s, cudaStreamWaitEvent(s->stream_io, s->ev_snap_packed, 0));
    if (n) {
        if (pos) CU(s, cudaMemcpyAsync(pos, b_pos, n * D * sizeof(double), cudaMemcpyDeviceToHost, s->stream_io));
        if (vel) CU(s, cudaMemcpyAsync(vel, b_vel, n * D * sizeof(double), cudaMemcpyDeviceToHost, s->stream_io));
        if (ids) CU(s, cudaMemcpyAsync(ids, b_id, n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream_io));
        if (voxel) CU(s, cudaMemcpyAsync(voxel, b_vox, n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream_io));
        if (hist_dx) CU(s, cudaMemcpyAsync(hist_dx, b_hx, n * 2 * D * sizeof(double), cudaMemcpyDeviceToHost, s->stream_io));
        if (hist_dv) CU(s, cudaMemcpyAsync(hist_dv, b_hv, n * 2 * D * sizeof(double), cudaMemcpyDeviceToHost, s->stream_io));
        if (hist_len || hist_dx || hist_dv)
            CU(s, cudaMemcpyAsync(s->snap_hlen_host, b_hl, n, cudaMemcpyDeviceToHost, s->stream_io));
    }
    CU(s, cudaEventRecord(s->ev_snap_done, s->stream_io));
    s->snap_pending = true;
    s->snap_n = n;
    s->snap_hx_host = hist_dx;
    s->snap_hv_host = hist_dv;
    s->snap_hl_out = hist_len;
    return 0;
}

int crb_sync(crb_sim* s) {
    if (!s) return CRB_ERR_INVALID_ARG;
    CU(s, cudaSetDevice((int)s->cfg.device));
    int rc;
    if ((rc = slab_finalize(s))) return rc;
    return check_latch(s);
}

int crb_debug_forces(crb_sim* s, uint64_t cap, double* force) {
    if (!s || !force) return CRB_ERR_INVALID_ARG;
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, kOpenStepMsg);
    if (s->slab_async_pending) return fail(s, CRB_ERR_STATE, kAsyncPendingMsg);
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "no agents");
    if (cap < s->n) return fail(s, CRB_ERR_CAPACITY, "output buffer too small");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods) return rods_debug_forces(s, force, nullptr);
    int rc;
    if (!s->force_out)
        if ((rc = dalloc(s, &s->force_out, s->cap_total * 3))) return rc;
    KParams P = make_params(s);
    if ((rc = launch_force(s, P, true))) return rc;
    s->launches += s->n ? 1 : 0;
    CU(s, cudaGetLastError());
    const double* cols[8] = {s->force_out, s->force_out + s->cap_total, s->force_out + 2 * s->cap_total, 0, 0, 0, 0, 0};
    if ((rc = download_rows(s, force, s->D, cols, s->n))) return rc;
    return check_latch(s);
}

int crb_get_identifiers(crb_sim* s, uint64_t cap, uint64_t* ident, uint64_t* parent) {
    if (!s) return CRB_ERR_INVALID_ARG;
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, kOpenStepMsg);
    if (s->slab_async_pending) return fail(s, CRB_ERR_STATE, kAsyncPendingMsg);
    if (cap < s->n) return fail(s, CRB_ERR_CAPACITY, "output buffers too small");
    CU(s, cudaSetDevice((int)s->cfg.device));
    if (s->rods) return rods_get_identifiers(s, cap, (u64*)ident, (u64*)parent);
    const u64 n = s->n;
    std::vector<u64> ids(n);
    CU(s, cudaMemcpyAsync(ids.data(), s->X.id + s->G, n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    for (u64 i = 0; i < n; ++i) {
        if (ident) {
            ident[i * 3 + 0] = 1;  // Initial(id)
            ident[i * 3 + 1] = ids[i];
            ident[i * 3 + 2] = 0;
        }
        if (parent) {
            parent[i * 3 + 0] = ~0ull;
            parent[i * 3 + 1] = 0;
            parent[i * 3 + 2] = 0;
        }
    }
    return 0;
}

uint64_t crb_n_divisions(const crb_sim* s) { return s ? s->last_divisions : 0; }

int crb_debug_neighbor_counts(crb_sim* s, uint64_t cap, uint64_t* counts) {
    if (!s || !counts) return CRB_ERR_INVALID_ARG;
    if (!s->rods) return fail(s, CRB_ERR_STATE, "neighbour sensing is only part of the CRB_MECH_ROD step");
    if (!s->have_agents) return fail(s, CRB_ERR_STATE, "no agents");
    if (cap < s->n) return fail(s, CRB_ERR_CAPACITY, "output buffer too small");
    CU(s, cudaSetDevice((int)s->cfg.device));
    return rods_debug_forces(s, nullptr, (u64*)counts);
}

int crb_get_migrations(crb_sim* s, uint64_t cap, uint64_t* ids, uint64_t* old_voxel,
                       uint64_t* new_voxel, uint64_t* n_out) {
    if (!s || !n_out) return CRB_ERR_INVALID_ARG;
    if (s->rods) return fail(s, CRB_ERR_STATE, "crb_get_migrations is not available for CRB_MECH_ROD");
    CU(s, cudaSetDevice((int)s->cfg.device));
    *n_out = 0;
    const u64 n = s->last_step_n;
    if (!n) return 0;
    // after the step the pre-step static fields (ids by old slot) live in Y
    u64 *d_ids, *d_old, *d_new;
    CU(s, cudaMalloc((void**)&d_ids, std::max<u64>(cap, 1) * 3 * sizeof(u64)));
    d_old = d_ids + std::max<u64>(cap, 1);
    d_new = d_old + std::max<u64>(cap, 1);
    cudaMemsetAsync(s->scratch_u64, 0, sizeof(u64), s->stream);
    k_collect_migrations<<<blocks_for(n, 256), 256, 0, s->stream>>>(
        s->sortkey, s->newkey, s->Y.id, s->g, s->G, n, cap, d_ids, d_old, d_new, s->scratch_u64);
    s->launches++;
    u64 cnt = 0;
    cudaMemcpyAsync(&cnt, s->scratch_u64, sizeof(u64), cudaMemcpyDeviceToHost, s->stream);
    cudaStreamSynchronize(s->stream);
    const u64 m = std::min<u64>(cnt, cap);
    if (ids) cudaMemcpy(ids, d_ids, m * sizeof(u64), cudaMemcpyDeviceToHost);
    if (old_voxel) cudaMemcpy(old_voxel, d_old, m * sizeof(u64), cudaMemcpyDeviceToHost);
    if (new_voxel) cudaMemcpy(new_voxel, d_new, m * sizeof(u64), cudaMemcpyDeviceToHost);
    cudaFree(d_ids);
    *n_out = cnt;
    CU(s, cudaGetLastError());
    if (cnt > cap) return fail(s, CRB_ERR_CAPACITY, "migration buffers too small");
    return 0;
}

int crb_get_voxel_counts(crb_sim* s, uint64_t cap, uint32_t* counts) {
    if (!s || !counts) return CRB_ERR_INVALID_ARG;
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, kOpenStepMsg);
    if (s->slab_async_pending) return fail(s, CRB_ERR_STATE, kAsyncPendingMsg);
    const u64 plane = (u64)s->g.ng[1] * s->g.ng[2];
    const u64 first = (u64)s->g.own_lo * plane, last = (u64)s->g.own_hi * plane;
    if (cap < last - first) return fail(s, CRB_ERR_CAPACITY, "output buffer too small");
    CU(s, cudaSetDevice((int)s->cfg.device));
    std::vector<u32> st(last - first + 1);
    CU(s, cudaMemcpyAsync(st.data(), s->start + first, st.size() * sizeof(u32),
                          cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    for (u64 v = 0; v + 1 < st.size(); ++v) counts[v] = st[v + 1] - st[v];
    return 0;
}

int crb_count_candidate_pairs(crb_sim* s, uint64_t* n_pairs) {
    if (!s || !n_pairs) return CRB_ERR_INVALID_ARG;
    if (s->step_open || s->interior_state) return fail(s, CRB_ERR_STATE, kOpenStepMsg);
    if (s->slab_async_pending) return fail(s, CRB_ERR_STATE, kAsyncPendingMsg);
    CU(s, cudaSetDevice((int)s->cfg.device));
    *n_pairs = 0;
    if (!s->n) return 0;
    CU(s, cudaMemsetAsync(s->scratch_u64, 0, sizeof(u64), s->stream));
    KParams P = make_params(s);
    k_count_pairs<<<blocks_for(s->n, 256), 256, 0, s->stream>>>(P, s->scratch_u64);
    s->launches++;
    CU(s, cudaMemcpyAsync(n_pairs, s->scratch_u64, sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    CU(s, cudaGetLastError());
    return 0;
}

int crb_enable_timers(crb_sim* s, int on) {
    if (!s) return CRB_ERR_INVALID_ARG;
    s->timers_on = on != 0;
    return 0;
}

int crb_reset_timers(crb_sim* s) {
    if (!s) return CRB_ERR_INVALID_ARG;
    timers_resolve(s);
    for (auto& t : s->timers) {
        t.ms = 0;
        t.launches = 0;
    }
    return 0;
}

int crb_stage_time_ms(crb_sim* s, uint32_t stage, double* ms, uint64_t* launches) {
    if (!s || stage >= (uint32_t)kNumStages) return CRB_ERR_INVALID_ARG;
    CU(s, cudaSetDevice((int)s->cfg.device));
    timers_resolve(s);
    if (ms) *ms = s->timers[stage].ms;
    if (launches) *launches = s->timers[stage].launches;
    return 0;
}

}  // extern "C"
