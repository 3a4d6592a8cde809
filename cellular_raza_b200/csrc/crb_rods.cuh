// crb_rods.cuh — device code of the bacterial-rods step (sm_100a): RodMechanics +
// RodInteraction<MorsePotential> with aspects [Mechanics, Interaction, Cycle, NeighborSensing]
// (BASELINE.json configs[3]).
//
// One step = k_rod_force  (candidate filter over the 3^D voxel neighbourhood, vertex-segment Morse
//                          forces of both half-calls, neighbour counting, fused spring/curvature
//                          increment + Adams-Bashforth update + growth + boundary + new voxel key)
//          -> k_rod_divide (Division events: mother/daughter geometry, identifiers, fresh AuxStorage)
//          -> k_scan_* / k_fill (shared with the point-agent path) -> k_rod_reorder.
//
// Layout: V vertices x 3 coordinates per rod as 3V structure-of-arrays planes:
// component c of vertex v of rod i lives at pos[(v * 3 + c) * cap + i].
//
// Arithmetic contract: everything that is ill-conditioned or discontinuous in the
// reference is evaluated with the reference's own operation order and no fma contraction
// (__dmul_rn/__dadd_rn/__ddiv_rn/__dsqrt_rn): the nearest point a*(1-t) + b*t in absolute
// coordinates, the spring / curvature forces (acos of a clamped cosine), the neighbour-count
// predicate, growth, division geometry, the Adams-Bashforth helpers, reflection and the voxel key.
// The Morse coefficient itself (rsqrt + exp) is evaluated to ~1e-15 relative.
#pragma once
#include "crb_kernels.cuh"

namespace crb {

constexpr int kRodThreads = 128;
constexpr int kRodList = 8;  // accepted-candidate slots per thread between two drains

struct RodState {
    double *pos, *vel;  // [(v*3+c)*cap + i]
    double *h1x, *h1v;  // newest AuxStorageMechanics entry (increments dx_k, dv_k)
    double *h0x, *h0v;  // older entry (only refreshed when save_prev)
    double *len;        // RodMechanics::spring_length
    double *growth;     // Agent::growth_rate
    double *cen;        // [4][cap]: mean of the vertices (x, y, z) and the bounding radius around it
    u64 *ident;         // [3][cap]: CellIdentifier {kind, a, b}
    u64 *parent;        // [3][cap]: parent identifier (kind == ~0: None)
    u8 *hlen;
};

struct RodParams {
    Grid g;
    RodState A, B;
    u64 n;    // owned rods in A: slots [first, first + n)
    u64 first;      // index of the first owned rod (ghost capacity G of a slab; 0 on a single device)
    u64 src_limit;  // one past the last slot a source (owned rod, daughter, arrival) may take
    u64 cap;  // plane stride
    u32 *key, *start, *newkey, *rank_arb, *count_new;
    u64 *sortkey;
    u8 *div_flag;     // [cap] Division event pending for source slot i
    u32 *id_counter;  // Voxel::id_counter per local voxel
    u32 *n_div;       // device counters: [0] daughters created this step, [1] entries of div_list
    u32 *div_list;    // slots of the rods with a pending Division event (written by the epilogue, any order)
    double *force_out;
    u64 *nb_out;
    int *err;
    double dt;
    double tension, rigidity, damping, threshold, growth_base;  // uniform RodMechanics / Agent fields
    double radius, stiffness, cutoff, strength;                 // uniform MorsePotential
    double reject;      // conservative centre-distance bound: max(cutoff, 2 radius) with slack
    double nb_thresh2;  // smallest s with sqrt_rn(s) >= 2 radius: (norm < 2 radius) <=> (s < nb_thresh2)
    int order, weight_mode, save_prev;
    // RodMechanics::get_random_contribution (bacterial_rods.rs:215-240): dpos = N(0, sqrt(dt)) * 2^-2 * D / dt
    int noise_mode;
    const double* noise;  // CRB_NOISE_TABLE: z[(step * n_ids + row) * 3V + c * V + v] (column-major fill of the matrix)
    u64 noise_ids, noise_cursor, iteration, seed;
    double diffusion;
    // aspect DomainForce (CRB_DOMAIN_FORCE_ROD_GEL): CartesianSubDomainRods (bacterial_rods.rs:433-477)
    int dforce;
    double gel_pressure, surface_friction, surface_friction_distance;
};

// row of the noise table / Philox key of a rod: Initial(k) rods use k, rods born by division fold their
// Division(voxel, counter) identifier (the oracle restates the same rule: both sides read the same numbers)
__host__ __device__ inline u64 rod_noise_key(u64 kind, u64 a, u64 b) {
    return kind == 1 ? a : ((a * 0x9E3779B97F4A7C15ull + b) * 0xBF58476D1CE4E5B9ull) >> 11;
}

// ---- exact-order helpers (nalgebra dot / norm of a 3-vector: (x*x + y*y) + z*z, no fma) -----------
__device__ __forceinline__ double dot3_rn(const double* a, const double* b) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
}
__device__ __forceinline__ double norm3_rn(const double* a) { return __dsqrt_rn(dot3_rn(a, a)); }

// ---- one vertex-segment evaluation ------------------------------------------------------------------
// nearest_point_from_point_to_line (interaction.rs:497-509) + calculate_morse_interaction (:151-182)
// for point p and segment (a, b): f = force on p (f_own); the force on the nearest point is -f.
// ab = b - a and p - a are exact (Sterbenz) in the reference and here; t uses a reciprocal of |ab|^2
// (relative error ~2e-16); the nearest point keeps the reference's a*(1-t) + b*t rounding.
struct RodPot {
    double rsum, stiffness, cutoff, m2sa;
};
__device__ __forceinline__ void rod_eval(const RodPot& M, const double* p, const double* a,
                                         const double* b, const double* ab, double inv_ab2, double* f,
                                         double* t_out) {
    const double ap0 = p[0] - a[0], ap1 = p[1] - a[1], ap2 = p[2] - a[2];
    double t = fma(ab[2], ap2, fma(ab[1], ap1, ab[0] * ap0)) * inv_ab2;
    t = !(t > 0.0) ? 0.0 : t;  // clamp(t, 0, 1) as compare-selects (NaN -> 0 like fmax): fmin/fmax on f64
    t = t > 1.0 ? 1.0 : t;     // cost 15 instructions per evaluation (ncu), this costs 6
    const double omt = __dsub_rn(1.0, t);
    double z[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
        z[c] = __dsub_rn(p[c], __dadd_rn(__dmul_rn(a[c], omt), __dmul_rn(b[c], t)));
    const double d2 = fma(z[2], z[2], fma(z[1], z[1], z[0] * z[0]));
    const double rinv = rsqrt_nr(d2);
    const double dist = d2 * rinv;
    // branch-free (a branch here stops the compiler from interleaving the five evaluations of a segment:
    // measured 3.7 -> 4.9 ms).  x <= stiffness * rsum < 700 is checked on the host (rods_set_agents); below
    // -700 the true value is < 1e-304 and the clamp changes the force by less than that.
    double x = -M.stiffness * (dist - M.rsum);
    x = x < -700.0 ? -700.0 : x;
    const double e = exp_bounded(x);
    double cf = M.m2sa * e * (1.0 - e) * rinv;
    cf = (d2 == 0.0 || dist > M.cutoff) ? 0.0 : cf;
#pragma unroll
    for (int c = 0; c < 3; ++c) f[c] = z[c] * cf;
    *t_out = t;
}

// ---- SubDomainForce of CartesianSubDomainRods (bacterial_rods.rs:433-477), added to the force of step 1
// (SubDomainBox::calculate_custom_domain_force, chili/update_mechanics.rs:248-274): every vertex feels
// (0, 0, -gel_pressure); a moving vertex within surface_friction_distance of a wall of the domain
// additionally -vel/|vel| * gel_pressure * surface_friction.  Reference operation order.
template <int V>
__device__ __forceinline__ void rod_domain_force(const RodParams& P, const double (*pos)[3], const double (*vel)[3],
                                                 double (*acc)[3]) {
    if (P.dforce != CRB_DOMAIN_FORCE_ROD_GEL) return;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        double f[3] = {0.0, 0.0, -P.gel_pressure};
        bool near = false;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            near |= fabs(__dsub_rn(pos[v][c], P.g.mn[c])) <= P.surface_friction_distance;
            near |= fabs(__dsub_rn(pos[v][c], P.g.mx[c])) <= P.surface_friction_distance;
        }
        const double q = norm3_rn(vel[v]);
        if (q != 0.0 && near) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double dir = __ddiv_rn(vel[v][c], q);
                f[c] = __dadd_rn(f[c], __dmul_rn(__dmul_rn(-dir, P.gel_pressure), P.surface_friction));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[v][c] = __dadd_rn(acc[v][c], f[c]);
    }
}

// ---- boundary + voxel key + bookkeeping shared by k_rod_force and k_rod_divide ----------------------
// CartesianSubDomainRods::apply_boundary (bacterial_rods.rs:489-533), SortCells::get_voxel_index_of
// (mean of the vertices, :550-554), sort key of sort_cells_in_voxels (update_mechanics.rs:427-476).
template <int V>
__device__ __forceinline__ void rod_finish(const RodParams& P, const u64 slot, double (*pos)[3],
                                           double (*vel)[3], const u32 v_old, const u64 rank_old,
                                           const u64 id_for_error) {
    const Grid& g = P.g;
    bool outside = false;
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double p = pos[v][c], w = vel[v][c];
            if (p < g.mn[c]) {
                p = __dsub_rn(__dmul_rn(g.mn[c], 2.0), p);
                w = fabs(w);
            }
            if (p > g.mx[c]) {
                p = __dsub_rn(__dmul_rn(g.mx[c], 2.0), p);
                w = -fabs(w);
            }
            outside |= (p < g.mn[c]) | (p > g.mx[c]);
            pos[v][c] = p;
            vel[v][c] = w;
        }
    if (outside) latch_error(P.err, CRB_ERR_BOUNDARY, id_for_error);
    // mean: nalgebra row_sum = fold(0, +) per column, then / V
    double mean[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < V; ++v) s = __dadd_rn(s, pos[v][c]);
        mean[c] = __ddiv_rn(s, (double)V);
    }
    double h2 = 0.0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const double a = pos[v][0] - mean[0], b = pos[v][1] - mean[1], c = pos[v][2] - mean[2];
        h2 = fmax(h2, fma(c, c, fma(b, b, a * a)));
    }
    const double hr = sqrt(h2) * (1.0 + 1e-9);
    int q0, q1, q2;
    u32 nk = v_old;
    const int n1 = g.ng[1], n2 = g.ng[2];
    if (voxel_of<3>(g, mean, &q0, &q1, &q2)) {
        const int l0 = q0 - g.p0;
        if (l0 >= 0 && l0 < g.L0) nk = ((u32)l0 * (u32)n1 + (u32)q1) * (u32)n2 + (u32)q2;
    }
    const u64 gp_old = (u64)v_old + (u64)g.p0 * n1 * n2;
    const u64 rk = rank_old > CRB_SK_RANK_MASK ? CRB_SK_RANK_MASK : rank_old;
    u64 sk = (gp_old << CRB_SK_RANK_BITS) | rk;
    if (nk != v_old) {
        sk |= CRB_SK_MOVED;
        // a rod that changes chili subdomain is appended after the local movers (crb_internal.cuh)
        if (g.sm_nsub > 1 && g.sub_of(gp_old) != g.sub_of((u64)nk + (u64)g.p0 * n1 * n2)) sk |= CRB_SK_REMOTE;
    }
    P.sortkey[slot] = sk;
    P.newkey[slot] = nk;
    P.rank_arb[slot] = atomicAdd(&P.count_new[nk], 1u);
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            P.B.pos[(u64)(v * 3 + c) * P.cap + slot] = pos[v][c];
            P.B.vel[(u64)(v * 3 + c) * P.cap + slot] = vel[v][c];
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) P.B.cen[(u64)c * P.cap + slot] = mean[c];
    P.B.cen[3ull * P.cap + slot] = hr;
}

// ------------------------------------------------------------------------------------------------
// Per-rod epilogue of a step, given the complete step-1 force `acc` and the neighbour count `nb`:
// RodMechanics::calculate_increment (bacterial_rods.rs:157-213), AB wrapper (chili/solvers.rs:156-300),
// react_to_neighbors (main.rs:66-73), update_cycle (:87-98); rods without a Division event finish here
// (rod_finish), the others are completed by k_rod_divide.  Shared by the fused kernel (k_rod_force) and the
// split pair (k_rod_force<FORCE_ONLY> + k_rod_update); ADD_DOMAIN = false when `acc` already holds the
// DomainForce contribution.
// ------------------------------------------------------------------------------------------------
template <int V, bool ADD_DOMAIN>
__device__ __forceinline__ void rod_epilogue(const RodParams& P, const u64 i, double (*own)[3], double (*acc)[3],
                                             const u32 nb, const u32 vk, const u32 own_lo) {
    const u64 cap = P.cap;
    // ---- RodMechanics::calculate_increment (bacterial_rods.rs:157-213), reference op order ----
    double vel[V][3];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) vel[v][c] = P.A.vel[(u64)(v * 3 + c) * cap + i];
    if (ADD_DOMAIN) rod_domain_force<V>(P, own, vel, acc);  // aspect DomainForce (no-op unless CRB_DOMAIN_FORCE_ROD_GEL)
    double len = P.A.len[i];
    double dist[V - 1][3], nrm[V - 1];
#pragma unroll
    for (int s = 0; s + 1 < V; ++s) {
#pragma unroll
        for (int c = 0; c < 3; ++c) dist[s][c] = __dsub_rn(own[s][c], own[s + 1][c]);
        nrm[s] = norm3_rn(dist[s]);
    }
#pragma unroll
    for (int s = 0; s + 1 < V; ++s) {
        if (nrm[s] != 0.0) {
            const double stretch = __dsub_rn(nrm[s], len);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double dir = __ddiv_rn(dist[s][c], nrm[s]);
                const double fi = __dmul_rn(__dmul_rn(-dir, P.tension), stretch);
                acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(fi, 0.5));
                acc[s + 1][c] = __dadd_rn(acc[s + 1][c], __dmul_rn(-fi, 0.5));
            }
        }
    }
#pragma unroll
    for (int s = 0; s + 2 < V; ++s) {
        const double prod = dot3_rn(dist[s], dist[s + 1]);
        double angle = 0.0;
        if (!(nrm[s] == 0.0 || nrm[s + 1] == 0.0)) {
            double cang = __ddiv_rn(prod, __dmul_rn(nrm[s], nrm[s + 1]));
            cang = cang < -1.0 ? -1.0 : cang;
            cang = cang > 1.0 ? 1.0 : cang;
            angle = acos(cang);
        }
        double fd[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) fd[c] = __dsub_rn(dist[s][c], dist[s + 1][c]);
        const double nfd = norm3_rn(fd);
        const double tn = tan(__dmul_rn(0.5, angle));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double dir = nfd != 0.0 ? __ddiv_rn(fd[c], nfd) : fd[c];
            const double f = __dmul_rn(__dmul_rn(__dmul_rn(dir, 2.0), P.rigidity), tn);
            acc[s][c] = __dadd_rn(acc[s][c], __dmul_rn(-f, 0.5));
            acc[s + 1][c] = __dadd_rn(acc[s + 1][c], f);
            acc[s + 2][c] = __dadd_rn(acc[s + 2][c], __dmul_rn(-f, 0.5));
        }
    }
    // ---- get_random_contribution (bacterial_rods.rs:215-240): dpos[v][c] = ((0 + sqrt(dt) z) * 2^-2 * D) / dt with
    // the matrix filled column-major from the generator (the k-th number goes to row k % V, column k / V); dvel = 0.
    // The update loop below runs in that order (coordinate outer, vertex inner), so only four numbers are live.
    const int noise_mode = P.noise_mode;
    u64 nkey = 0;
    const double* zrow = nullptr;
    double sd = 0.0;
    if (noise_mode != CRB_NOISE_NONE) {
        nkey = rod_noise_key(P.A.ident[i], P.A.ident[cap + i], P.A.ident[2 * cap + i]);
        sd = sqrt(P.dt);
        if (noise_mode == CRB_NOISE_TABLE)
            zrow = P.noise + (P.noise_cursor * P.noise_ids + nkey % (P.noise_ids ? P.noise_ids : 1)) * (u64)(3 * V);
    }
    // ---- AB wrapper (chili/solvers.rs:156-300): dx = vel, dv = total force - vel * damping -------
    const int N = P.order;
    const int hlen = P.A.hlen[i];
    const int n_prev = N == 0 ? 0 : (hlen + 1 < N ? hlen + 1 : N);
    double zn[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int k = c * V + v;
            double rxv = 0.0;
            if (noise_mode != CRB_NOISE_NONE) {
                double z;
                if (noise_mode == CRB_NOISE_TABLE) {
                    z = zrow[k];
                } else {
                    if ((k & 3) == 0) philox_normal4(P.seed, nkey, P.iteration, (u32)(k >> 2), zn);
                    z = zn[k & 3];
                }
                const double m = __dadd_rn(0.0, __dmul_rn(sd, z));
                rxv = __ddiv_rn(__dmul_rn(__dmul_rn(m, 0.25), P.diffusion), P.dt);
            }
            const u64 o = (u64)(v * 3 + c) * cap + i;
            const double dxv = vel[v][c];
            const double dvv = __dsub_rn(acc[v][c], __dmul_rn(vel[v][c], P.damping));
            const double oldx = P.A.h1x[o], oldv = P.A.h1v[o];
            double pn, wn;
            if (N == 2 && n_prev == 2) {
                pn = helper_ab3(own[v][c], dxv, oldx, dxv, P.dt, rxv);
                wn = helper_ab3(vel[v][c], dvv, oldv, dvv, P.dt, 0.0);
            } else if (n_prev == 1) {
                pn = helper_ab2(own[v][c], dxv, dxv, P.dt, rxv);
                wn = helper_ab2(vel[v][c], dvv, dvv, P.dt, 0.0);
            } else {
                pn = helper_euler(own[v][c], dxv, P.dt, rxv);
                wn = helper_euler(vel[v][c], dvv, P.dt, 0.0);
            }
            own[v][c] = pn;
            vel[v][c] = wn;
            P.B.h1x[o] = N >= 1 ? dxv : oldx;
            P.B.h1v[o] = N >= 1 ? dvv : oldv;
            if (P.save_prev) {
                P.B.h0x[o] = oldx;
                P.B.h0v[o] = oldv;
            }
        }
    P.B.hlen[i] = (u8)n_prev;
    // ---- react_to_neighbors (main.rs:66-73) and update_cycle (:87-98) ---------------------------
    double growth;
    if (nb > 0)
        growth = fmax(__ddiv_rn(__dmul_rn(P.growth_base, __dsub_rn(8.0, (double)nb)), 8.0), 0.0);
    else
        growth = P.growth_base;
    len = __dadd_rn(len, __dmul_rn(growth, P.dt));
    P.B.growth[i] = growth;
    P.B.len[i] = len;
    const bool divide = len > P.threshold;
    P.div_flag[i] = divide ? 1 : 0;
    if (divide) {
        P.div_list[atomicAdd(P.n_div + 1, 1u)] = (u32)i;
        // completed by k_rod_divide (geometry changes before boundary and voxel key)
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                P.B.pos[(u64)(v * 3 + c) * cap + i] = own[v][c];
                P.B.vel[(u64)(v * 3 + c) * cap + i] = vel[v][c];
            }
    } else {
        rod_finish<V>(P, i, own, vel, vk, i - own_lo, P.A.ident[cap + i]);
    }
}

// ------------------------------------------------------------------------------------------------
// k_rod_force: one thread per rod of A (sorted by voxel).
//
// Phase 1 (filter): the thread walks the <= 9 contiguous candidate ranges of its 3^D neighbourhood and
//   rejects candidates whose centre distance exceeds reject + h_i + h_j (every vertex / segment point
//   of a rod lies within h of its centre, so a rejected pair has all vertex-segment distances beyond
//   the Morse cutoff AND beyond 2 radius: its contribution is exactly zero).  Survivors go to a
//   per-thread list in shared memory.
// Phase 2 (drain, warp-synchronous): every lane evaluates its own list in ascending candidate order:
//   RodInteraction::calculate_force_between (bacterial_rods.rs:292-336) from both sides (own half always,
//   ext half per the chili pair weights, update_mechanics.rs:186-234) and the 25 vertex-vertex checks of
//   accumulate_information (examples/bacterial_rods/src/main.rs:48-64).
// Fixed order per rod => bitwise run-to-run determinism.
// Epilogue: RodMechanics::calculate_increment (bacterial_rods.rs:157-213), AB wrapper
//   (chili/solvers.rs:156-300), react_to_neighbors (main.rs:66-73), update_cycle (:87-98); rods without
//   a Division event finish here (rod_finish), the others are completed by k_rod_divide.
// ------------------------------------------------------------------------------------------------
template <int V, bool FORCE_ONLY>
__global__ void __launch_bounds__(kRodThreads) k_rod_force(const RodParams P) {
    __shared__ u32 s_list[kRodList * kRodThreads];
    __shared__ double s_ext[3 * V * kRodThreads];
    const int tid = threadIdx.x;
    const u64 t = (u64)blockIdx.x * kRodThreads + tid;
    const bool valid = t < P.n;
    const u64 i = P.first + (valid ? t : 0);
    const Grid& g = P.g;
    const int n1 = g.ng[1], n2 = g.ng[2];
    const u64 cap = P.cap;

    double own[V][3];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) own[v][c] = P.A.pos[(u64)(v * 3 + c) * cap + i];
    const double cx = P.A.cen[i], cy = P.A.cen[cap + i], cz = P.A.cen[2 * cap + i];
    const double hi_r = P.A.cen[3 * cap + i] + P.reject;

    const u32 vk = P.key[i];
    const int i2 = (int)(vk % (u32)n2);
    const u32 tq = vk / (u32)n2;
    const int i1 = (int)(tq % (u32)n1);
    const int i0 = (int)(tq / (u32)n1);
    const u32 own_lo = P.start[vk], own_hi = P.start[vk + 1];
    u32 sub_lo = 0, sub_hi = 0;
    if (P.weight_mode == CRB_WEIGHT_CHILI) {
        const u64 plane = (u64)n1 * n2;
        const u64 gp = ((u64)(i0 + g.p0) * n1 + i1) * n2 + i2;
        const u64 s = g.sub_of(gp);
        const u64 base = (u64)g.p0 * plane, nloc = g.nvox_local();
        u64 f = g.sub_first_voxel(s), l = g.sub_first_voxel(s + 1);
        f = f < base ? 0 : (f - base > nloc ? nloc : f - base);
        l = l < base ? 0 : (l - base > nloc ? nloc : l - base);
        sub_lo = P.start[f];
        sub_hi = P.start[l];
    }

    RodPot M;
    M.rsum = P.radius + P.radius;
    M.stiffness = P.stiffness;
    M.cutoff = P.cutoff;
    M.m2sa = -2.0 * P.strength * P.stiffness;

    double acc[V][3];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[v][c] = 0.0;
    u32 nb = 0;
    int cnt = 0;

    // own segments: ab and 1/|ab|^2 are needed by every ext-half evaluation
    double oab[V - 1][3], oinv[V - 1];
#pragma unroll
    for (int s = 0; s + 1 < V; ++s) {
#pragma unroll
        for (int c = 0; c < 3; ++c) oab[s][c] = own[s + 1][c] - own[s][c];
        oinv[s] = 1.0 / fma(oab[s][2], oab[s][2], fma(oab[s][1], oab[s][1], oab[s][0] * oab[s][0]));
    }

    // the candidate rod is staged in shared memory (one column per thread: only the owning thread
    // touches it, so no barrier) to allow rolled loops over its vertices / segments
    double* const ex = s_ext + tid;
    auto EXT = [&](int v, int c) -> double { return ex[(v * 3 + c) * kRodThreads]; };
    auto pair = [&](const u32 j) {
#pragma unroll
        for (int w = 0; w < 3 * V; ++w) ex[w * kRodThreads] = __ldg(&P.A.pos[(u64)w * cap + j]);
        const bool same_voxel = (j >= own_lo) & (j < own_hi);
        const bool same_sub = (j >= sub_lo) & (j < sub_hi);
        const bool keep = same_voxel | !same_sub;
        // own half: F(own = i, ext = j): own vertices against ext segments
#pragma unroll 1
        for (int s = 0; s + 1 < V; ++s) {
            double a[3], b[3], ab[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                a[c] = EXT(s, c);
                b[c] = EXT(s + 1, c);
                ab[c] = b[c] - a[c];
            }
            const double inv = 1.0 / fma(ab[2], ab[2], fma(ab[1], ab[1], ab[0] * ab[0]));
#pragma unroll
            for (int v = 0; v < V; ++v) {
                double f[3], tt;
                rod_eval(M, own[v], a, b, ab, inv, f, &tt);
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[v][c] = fma(0.5, f[c], acc[v][c]);
            }
        }
        // ext half: F(own = j, ext = i): ext vertices against own segments; the force on the nearest
        // point (-f) is split (1 - t, t) over the two vertices of the own segment
        if (keep) {
#pragma unroll 1
            for (int v = 0; v < V; ++v) {
                double p[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) p[c] = EXT(v, c);
#pragma unroll
                for (int s = 0; s + 1 < V; ++s) {
                    double f[3], tt;
                    rod_eval(M, p, own[s], own[s + 1], oab[s], oinv[s], f, &tt);
                    const double w1 = -0.5 * tt, w0 = -0.5 * (1.0 - tt);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        acc[s][c] = fma(w0, f[c], acc[s][c]);
                        acc[s + 1][c] = fma(w1, f[c], acc[s + 1][c]);
                    }
                }
            }
        }
        // NeighborSensing::accumulate_information: (own_point - ext_point).norm() < 2 radius
#pragma unroll 1
        for (int b = 0; b < V; ++b) {
            double e[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) e[c] = EXT(b, c);
#pragma unroll
            for (int a = 0; a < V; ++a) {
                double z[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) z[c] = __dsub_rn(own[a][c], e[c]);
                nb += dot3_rn(z, z) < P.nb_thresh2 ? 1u : 0u;
            }
        }
    };
    auto drain = [&]() {
        int kmax = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        for (int k = 0; k < kmax; ++k)
            if (k < cnt) pair(s_list[k * kRodThreads + tid]);
        cnt = 0;
    };

    // ---- phase 1: candidate walk ---------------------------------------------------------------
    int r = -1;
    u32 j = 0, je = 0;
    auto next_range = [&]() -> bool {
        while (++r < 9) {
            const int X = i0 + r / 3 - 1, Y = i1 + r % 3 - 1;
            if (X < 0 || X >= g.L0 || Y < 0 || Y >= n1) continue;
            const int z_lo = i2 > 0 ? i2 - 1 : 0, z_hi = i2 + 1 < n2 ? i2 + 1 : n2 - 1;
            const u32 row = ((u32)X * (u32)n1 + (u32)Y) * (u32)n2;
            j = P.start[row + z_lo];
            je = P.start[row + z_hi + 1];
            if (j < je) return true;
        }
        return false;
    };
    bool active = valid && next_range();
    for (;;) {
        const bool any_active = __any_sync(0xffffffffu, active);
        if (active) {
            const u32 jj = j;
            if (jj != (u32)i) {
                const double dx = cx - __ldg(&P.A.cen[jj]), dy = cy - __ldg(&P.A.cen[cap + jj]),
                             dz = cz - __ldg(&P.A.cen[2 * cap + jj]);
                const double lim = hi_r + __ldg(&P.A.cen[3 * cap + jj]);
                if (fma(dz, dz, fma(dy, dy, dx * dx)) <= lim * lim) {
                    // (a second-level test on the 25 vertex-vertex distances was measured: it rejects too few
                    //  of these candidates to pay for its 15 gathered loads, 4.21 ms vs 3.71 ms at 2 M rods)
                    s_list[cnt * kRodThreads + tid] = jj;
                    ++cnt;
                }
            }
            if (++j == je) active = next_range();
        }
        if (!any_active || __any_sync(0xffffffffu, cnt == kRodList)) drain();
        if (!any_active) break;
    }
    if (!valid) return;

    if constexpr (FORCE_ONLY) {
        if (P.dforce == CRB_DOMAIN_FORCE_ROD_GEL) {
            double vel[V][3];
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
                for (int c = 0; c < 3; ++c) vel[v][c] = P.A.vel[(u64)(v * 3 + c) * cap + i];
            rod_domain_force<V>(P, own, vel, acc);
        }
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
            for (int c = 0; c < 3; ++c) P.force_out[(u64)(v * 3 + c) * cap + i] = acc[v][c];
        P.nb_out[i] = nb;
        return;
    } else {
        rod_epilogue<V, true>(P, i, own, acc, nb, vk, own_lo);
    }
}

// split variant: k_rod_force<V, true> writes force_out / nb_out, this kernel runs the epilogue at full
// occupancy (the fused kernel holds 8 warps per SM at 239 registers and pays the cold instruction fetches of
// this once-per-rod code in full)
template <int V>
__global__ void __launch_bounds__(128, 4) k_rod_update(const RodParams P) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n) return;
    const u64 i = P.first + t;
    const u64 cap = P.cap;
    // the epilogue reads velocity and increment history plane by plane inside its update loop, two loads at a time
    // between stores: ask for all of them now (no registers; 0.65 -> 0.62 ms at 2 M rods, same-box A/B)
#pragma unroll
    for (int w = 0; w < 3 * V; ++w) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&P.A.vel[(u64)w * cap + i]));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&P.A.h1x[(u64)w * cap + i]));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&P.A.h1v[(u64)w * cap + i]));
    }
    double own[V][3], acc[V][3];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            own[v][c] = P.A.pos[(u64)(v * 3 + c) * cap + i];
            acc[v][c] = P.force_out[(u64)(v * 3 + c) * cap + i];
        }
    const u32 vk = P.key[i];
    rod_epilogue<V, false>(P, i, own, acc, (u32)P.nb_out[i], vk, P.start[vk]);
}

// ------------------------------------------------------------------------------------------------
// k_rod_divide: Voxel::update_cell_cycle_4 (chili/update_cycle.rs:9-69) + Cycle::divide of the example
// (main.rs:100-108) + RodMechanics::divide (bacterial_rods.rs:586-663).  One thread per source slot with
// a pending Division event.  In voxel v with k dividing mothers (in-voxel order m = 0..k-1) and counter
// c0: mother m becomes Division(v, c0 + m + 1), her daughter Division(v, c0 + k + m + 1) and is appended
// behind all cells of v in mother order; the daughter starts with a fresh AuxStorage.
// ------------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void rod_divide_one(const RodParams& P, const u64 i);
template <int V>
__global__ void __launch_bounds__(128) k_rod_divide(const RodParams P) {
    // one thread per entry of the division list (a few per mille of the rods divide in a step: a thread per rod
    // spent 0.09 ms at 2 M rods on reading flags with the 254 registers of this kernel)
    const u32 nl = P.n_div[1];
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < nl; e += gridDim.x * blockDim.x) rod_divide_one<V>(P, P.div_list[e]);
}
template <int V>
__device__ __forceinline__ void rod_divide_one(const RodParams& P, const u64 i) {
    const u64 cap = P.cap;
    const u32 vk = P.key[i];
    const u32 lo = P.start[vk], hi = P.start[vk + 1];
    u32 m = 0, k = 0;
    for (u32 q = lo; q < hi; ++q) {
        const u32 fl = P.div_flag[q];
        k += fl;
        m += q < (u32)i ? fl : 0u;
    }
    const u32 c0 = P.id_counter[vk];
    const u64 d = P.first + P.n + atomicAdd(P.n_div, 1u);
    if (d >= P.src_limit) {
        latch_error(P.err, CRB_ERR_CAPACITY, P.A.ident[cap + i]);
        return;
    }
    double pos[V][3], vel[V][3];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            pos[v][c] = P.B.pos[(u64)(v * 3 + c) * cap + i];
            vel[v][c] = P.B.vel[(u64)(v * 3 + c) * cap + i];
        }
    // ---- RodMechanics::divide, reference op order ----------------------------------------------
    const double len = P.B.len[i];
    const double div_factor = __dsub_rn(0.5, __ddiv_rn(P.radius, __dmul_rn((double)V, len)));
    const double new_len = __dmul_rn(div_factor, len);
    double seg[V - 1];
#pragma unroll
    for (int s = 0; s + 1 < V; ++s) {
        double z[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) z[c] = __dsub_rn(pos[s][c], pos[s + 1][c]);
        seg[s] = norm3_rn(z);
    }
    // prefix sums as the reference computes them: take(n).sum() and rev().take(n).sum()
    double pre[V], rev[V];
    pre[0] = 0.0;
    rev[0] = 0.0;
#pragma unroll
    for (int s = 0; s + 1 < V; ++s) {
        pre[s + 1] = __dadd_rn(pre[s], seg[s]);
        rev[s + 1] = __dadd_rn(rev[s], seg[V - 2 - s]);
    }
    const double segment_length =
        __ddiv_rn(__ddiv_rn(__dsub_rn(pre[V - 1], __dmul_rn(2.0, P.radius)), (double)(V - 1)), 2.0);
    double c1[V][3], c2[V][3];
    bool ok = true;
#pragma unroll
    for (int nv = 0; nv < V; ++nv) {
        const double target = __dmul_rn((double)nv, segment_length);
        int kk = -1, mm = -1;
#pragma unroll
        for (int n = 0; n < V - 1; ++n)
            if (pre[n] <= target) kk = n;
#pragma unroll
        for (int n = 0; n < V; ++n)
            if (rev[n] <= target) mm = n;
        if (kk < 0 || mm < 0 || V - mm - 2 < 0) {
            ok = false;
            kk = 0;
            mm = 0;
        }
        double segk = seg[0], prek = pre[0], revm = rev[0], segm = seg[V - 2];
        double pk[3], pk1[3], qm[3], qm1[3];
#pragma unroll
        for (int n = 0; n < V - 1; ++n)
            if (n == kk) {
                segk = seg[n];
                prek = pre[n];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    pk[c] = pos[n][c];
                    pk1[c] = pos[n + 1][c];
                }
            }
#pragma unroll
        for (int n = 0; n < V - 1; ++n)
            if (n == mm) {
                revm = rev[n];
                segm = seg[V - n - 2];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    qm[c] = pos[V - n - 1][c];
                    qm1[c] = pos[V - n - 2][c];
                }
            }
        const double q = __ddiv_rn(__dsub_rn(target, prek), segk);
        const double p = __ddiv_rn(__dsub_rn(target, revm), segm);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            c1[nv][c] = __dadd_rn(__dmul_rn(pk[c], __dsub_rn(1.0, q)), __dmul_rn(pk1[c], q));
            c2[V - nv - 1][c] = __dadd_rn(__dmul_rn(qm[c], __dsub_rn(1.0, p)), __dmul_rn(qm1[c], p));
        }
    }
    if (!ok) {
        latch_error(P.err, CRB_ERR_DIVISION, P.A.ident[cap + i]);
        return;
    }
    // ---- identifiers (update_cycle.rs:33-39, 58-66) -----------------------------------------------
    const u64 gv = (u64)vk + (u64)P.g.p0 * P.g.ng[1] * P.g.ng[2];
    const u64 pk0 = P.A.ident[i], pa = P.A.ident[cap + i], pb = P.A.ident[2 * cap + i];
    P.A.parent[i] = pk0;
    P.A.parent[cap + i] = pa;
    P.A.parent[2 * cap + i] = pb;
    P.A.ident[i] = 0;
    P.A.ident[cap + i] = gv;
    P.A.ident[2 * cap + i] = (u64)c0 + m + 1;
    P.A.parent[d] = pk0;
    P.A.parent[cap + d] = pa;
    P.A.parent[2 * cap + d] = pb;
    P.A.ident[d] = 0;
    P.A.ident[cap + d] = gv;
    P.A.ident[2 * cap + d] = (u64)c0 + k + m + 1;
    // ---- state of mother and daughter -----------------------------------------------------------
    P.B.len[i] = new_len;
    P.B.len[d] = new_len;
    P.B.growth[d] = P.B.growth[i];
    P.B.hlen[d] = 0;
    P.div_flag[d] = 0;
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const u64 o = (u64)(v * 3 + c) * cap + d;
            P.B.h1x[o] = 0.0;
            P.B.h1v[o] = 0.0;
            P.B.h0x[o] = 0.0;
            P.B.h0v[o] = 0.0;
        }
    double vel2[V][3];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) vel2[v][c] = vel[v][c];
    rod_finish<V>(P, i, c1, vel, vk, i - lo, pa);
    rod_finish<V>(P, d, c2, vel2, vk, (u64)(hi - lo) + m, pa);
}

// Voxel::id_counter += 2 per division (one for the mother, one for the daughter)
__global__ void k_rod_bump_counters(const u32* __restrict__ div_list, const u32* __restrict__ n_div,
                                    const u32* __restrict__ key, u32* __restrict__ id_counter) {
    const u32 nl = n_div[1];
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < nl; e += gridDim.x * blockDim.x)
        atomicAdd(&id_counter[key[div_list[e]]], 2u);
}

// insert_initial_cells (chili/datastructures.rs:278-299): voxel.id_counter = number of initial cells
__global__ void k_rod_init_counters(const u32* __restrict__ start, u32* __restrict__ id_counter, u64 nvox) {
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nvox) id_counter[v] = start[v + 1] - start[v];
}

// ------------------------------------------------------------------------------------------------
// stable counting-sort scatter of the rod state (same ranking rule as k_reorder)
// ------------------------------------------------------------------------------------------------
struct RodReorderParams {
    RodState A, B, Y;
    const u32 *newkey, *start_new, *tmp_idx;
    const u64* sortkey;
    u32* key_out;
    u64 n, cap;
    u64 first;  // destination slots [first, first + n)
    int W;  // 3 V planes
    int save_prev;
};
// latency bound like k_reorder: occupancy buys bandwidth.  Groups of 4 planes x 4 arrays (16 loads in flight) at
// (256, 4) = 64 registers without spills: 0.58 -> 0.45 ms at 2 M rods against the unbounded 118-register build with
// groups of 6 (groups of 6 at (256, 3): 0.49 ms; groups of 3 at (256, 5): 0.46 ms, spills) — one box, round 2
constexpr int kRodReorderGroup = 4;
__global__ void __launch_bounds__(256, 4) k_rod_reorder(const RodReorderParams R) {
    const u64 tp = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (tp >= R.n) return;
    const u64 p = R.first + tp;
    const u32 i = R.tmp_idx[p];
    const u32 v = R.newkey[i];
    const u32 sb = R.start_new[v], se = R.start_new[v + 1];
    const u64 my = R.sortkey[i];
    u32 rank = 0;
    for (u32 q = sb; q < se; ++q) rank += (R.sortkey[R.tmp_idx[q]] < my) ? 1u : 0u;
    const u64 d = (u64)sb + rank;
    R.key_out[d] = v;
    // latency-bound gather: groups of kRodReorderGroup planes x 4 arrays of independent loads in flight per thread
    // before the first store (one long-scoreboard stall per group instead of one per load)
    for (int w0 = 0; w0 < R.W; w0 += kRodReorderGroup) {
        double a[kRodReorderGroup], b[kRodReorderGroup], c[kRodReorderGroup], e[kRodReorderGroup];
#pragma unroll
        for (int k = 0; k < kRodReorderGroup; ++k) {
            if (w0 + k < R.W) {
                const u64 oi = (u64)(w0 + k) * R.cap + i;
                a[k] = __ldg(&R.B.pos[oi]);
                b[k] = __ldg(&R.B.vel[oi]);
                c[k] = __ldg(&R.B.h1x[oi]);
                e[k] = __ldg(&R.B.h1v[oi]);
            }
        }
#pragma unroll
        for (int k = 0; k < kRodReorderGroup; ++k) {
            if (w0 + k < R.W) {
                const u64 od = (u64)(w0 + k) * R.cap + d;
                R.Y.pos[od] = a[k];
                R.Y.vel[od] = b[k];
                R.Y.h1x[od] = c[k];
                R.Y.h1v[od] = e[k];
            }
        }
    }
    if (R.save_prev) {
        for (int w = 0; w < R.W; ++w) {
            const u64 oi = (u64)w * R.cap + i, od = (u64)w * R.cap + d;
            R.Y.h0x[od] = R.B.h0x[oi];
            R.Y.h0v[od] = R.B.h0v[oi];
        }
    }
    for (int w = 0; w < 4; ++w) R.Y.cen[(u64)w * R.cap + d] = R.B.cen[(u64)w * R.cap + i];
    for (int w = 0; w < 3; ++w) {
        R.Y.ident[(u64)w * R.cap + d] = R.A.ident[(u64)w * R.cap + i];
        R.Y.parent[(u64)w * R.cap + d] = R.A.parent[(u64)w * R.cap + i];
    }
    R.Y.len[d] = R.B.len[i];
    R.Y.growth[d] = R.B.growth[i];
    R.Y.hlen[d] = R.B.hlen[i];
}

// ---- ingestion ---------------------------------------------------------------------------------
// agent-major host rows (width W) <-> planes
__global__ void k_rod_unpack(const double* __restrict__ rows, double* __restrict__ planes, int W, u64 cap,
                             u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < W; ++w) planes[(u64)w * cap + i] = rows[i * W + w];
}
__global__ void k_rod_pack(double* __restrict__ rows, const double* __restrict__ planes, int W, u64 cap,
                           u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < W; ++w) rows[i * W + w] = planes[(u64)w * cap + i];
}
__global__ void k_rod_pack_u64(u64* __restrict__ rows, const u64* __restrict__ planes, int W, u64 cap, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < W; ++w) rows[i * W + w] = planes[(u64)w * cap + i];
}
// crb_set_agents_async: the check of the parameter rows (every RodMechanics / MorsePotential field except
// spring_length and growth_rate is one value for all rods) and the extraction of the two per-rod columns happen on
// the device, from the rows as they were uploaded; a mismatch is latched like any other device-side error
struct RodUniformRow {
    double mp[8], ip[4];
};
__global__ void k_rod_param_rows(const double* __restrict__ mech_rows, const double* __restrict__ inter_rows,
                                 const RodUniformRow U, double* __restrict__ len, double* __restrict__ growth,
                                 u64 n, int* err) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool bad = false;
    double l = U.mp[3], g = U.mp[5];
    if (mech_rows) {
        const double* r = mech_rows + i * 8;
        bad = (r[0] != U.mp[0]) | (r[1] != U.mp[1]) | (r[2] != U.mp[2]) | (r[4] != U.mp[4]) | (r[6] != U.mp[6]) |
              (r[7] != U.mp[7]);
        l = r[3];
        g = r[5];
    }
    if (inter_rows) {
        const double* r = inter_rows + i * 4;
        bad |= (r[0] != U.ip[0]) | (r[1] != U.ip[1]) | (r[2] != U.ip[2]) | (r[3] != U.ip[3]);
    }
    len[i] = l;
    growth[i] = g;
    if (bad) latch_error(err, CRB_ERR_INVALID_ARG, i);
}
// clears as a kernel (see k_zero_cols: a copy-engine memset of a pipelined handle queues behind the host transfer of
// the handle it alternates with)
struct RodZeroRanges {
    double* p[8];
    u64 count[8];
    u8* b[2];
    u64 nb[2];
    int n;
};
__global__ void k_rod_zero(const RodZeroRanges Z) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 t0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < Z.n; ++k)
        for (u64 i = t0; i < Z.count[k]; i += stride) Z.p[k][i] = 0.0;
    for (int k = 0; k < 2; ++k)
        for (u64 i = t0; i < Z.nb[k]; i += stride) Z.b[k][i] = 0;
}
// the division count and / or the error latch into page-locked host words (device-accessible under unified
// addressing), read by the host after the stream synchronises
__global__ void k_rod_report(const u32* __restrict__ n_div, const int* __restrict__ err, u64* host_count, int* host_err) {
    const int t = threadIdx.x;
    if (host_count && t == 0) host_count[0] = n_div[0];
    if (host_err && t < 4) host_err[t] = err[t];
    __threadfence_system();
}
__global__ void k_rod_init_ident(RodState X, const u64* __restrict__ ids, u64 cap, u64 n, u64 first) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u64 i = first + t;
    X.ident[i] = 1;  // Initial(n)
    X.ident[cap + i] = ids ? ids[t] : t;
    X.ident[2 * cap + i] = 0;
    X.parent[i] = ~0ull;
    X.parent[cap + i] = 0;
    X.parent[2 * cap + i] = 0;
}

// setup-time binning: CartesianCuboidRods::get_voxel_index_of = mean of the vertices through
// get_voxel_index_of_raw (bacterial_rods.rs:368-372, cartesian_cuboid_n.rs:320-337)
template <int V>
__global__ void k_rod_ingest(const RodParams P) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n) return;
    const u64 i = P.first + t;
    const Grid& g = P.g;
    const u64 cap = P.cap;
    double mean[3];
    bool below = false, nan = false;
    double pos[V][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            pos[v][c] = P.B.pos[(u64)(v * 3 + c) * cap + i];
            s = __dadd_rn(s, pos[v][c]);
        }
        mean[c] = __ddiv_rn(s, (double)V);
        below |= (g.mn[c] >= mean[c]);
        nan |= !(mean[c] == mean[c]);
    }
    double h2 = 0.0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const double a = pos[v][0] - mean[0], b = pos[v][1] - mean[1], c = pos[v][2] - mean[2];
        h2 = fmax(h2, fma(c, c, fma(b, b, a * a)));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) P.B.cen[(u64)c * cap + i] = mean[c];
    P.B.cen[3 * cap + i] = sqrt(h2) * (1.0 + 1e-9);
    int q0 = 0, q1 = 0, q2 = 0;
    u32 nk = 0;
    if (below || nan) {
        latch_error(P.err, CRB_ERR_BOUNDARY, i);
    } else if (!voxel_of<3>(g, mean, &q0, &q1, &q2)) {
        latch_error(P.err, CRB_ERR_INDEX, i);
    } else {
        const int l0 = q0 - g.p0;
        if (l0 < g.own_lo || l0 >= g.own_hi)
            latch_error(P.err, CRB_ERR_INDEX, i);
        else
            nk = ((u32)l0 * (u32)g.ng[1] + (u32)q1) * (u32)g.ng[2] + (u32)q2;
    }
    P.newkey[i] = nk;
    P.sortkey[i] = t;
    P.rank_arb[i] = atomicAdd(&P.count_new[nk], 1u);
}

// ------------------------------------------------------------------------------------------------
// slab exchange of rods (the PosInformation / SendCell messages of chili/update_mechanics.rs:12-61
// between subdomains on different devices).  Same protocol as the point agents (crb_kernels.cuh): one fixed-size
// halo buffer and one fixed-size migration buffer per face, every column a plane of `cap` entries.
//   halo:      [n, mode | rel[plane + 1] | pos 3V planes | cen 4 planes]           (what k_rod_force reads of a candidate)
//   migration: [n, mode | pos, vel, h1x, h1v, h0x, h0v: 6 x 3V planes | len | growth | cen 4 | ident 3 | parent 3 |
//               sortkey | global voxel | hlen]
// ------------------------------------------------------------------------------------------------
struct RodHaloLayout {
    u64 gcap, plane, bytes;
    u64 off_rel, off_pos, off_cen;  // pos: 3V planes of gcap doubles, cen: 4 planes
    int W;
};
__global__ void k_rod_halo_export(const RodState A, u64 cap, const u32* __restrict__ start, u32 v0, RodHaloLayout L,
                                  unsigned char* __restrict__ buf, int* err) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 a = start[v0], b = start[v0 + (u32)L.plane];
    const u64 n = b - a;
    if (t == 0) {
        reinterpret_cast<u64*>(buf)[0] = n;
        reinterpret_cast<u64*>(buf)[1] = (u64)L.W;
        if (n > L.gcap) latch_error(err, CRB_ERR_CAPACITY, n);
    }
    if (t <= L.plane) reinterpret_cast<u32*>(buf + L.off_rel)[t] = start[v0 + (u32)t] - a;
    if (t < n && t < L.gcap) {
        const u64 i = a + t;
        double* pos = reinterpret_cast<double*>(buf + L.off_pos);
        double* cen = reinterpret_cast<double*>(buf + L.off_cen);
        for (int w = 0; w < L.W; ++w) pos[(u64)w * L.gcap + t] = A.pos[(u64)w * cap + i];
        for (int w = 0; w < 4; ++w) cen[(u64)w * L.gcap + t] = A.cen[(u64)w * cap + i];
    }
}
// lower ghost plane: the rods end at `edge` (= first owned slot); upper: they begin at `edge` (= first + n)
__global__ void k_rod_halo_import(RodState A, u64 cap, u32* __restrict__ start, u32 v0, u64 edge, int lower,
                                  RodHaloLayout L, const unsigned char* __restrict__ buf, int* err) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 n = reinterpret_cast<const u64*>(buf)[0];
    if (n > L.gcap || reinterpret_cast<const u64*>(buf)[1] != (u64)L.W) {
        if (t == 0) latch_error(err, n > L.gcap ? CRB_ERR_CAPACITY : CRB_ERR_STATE, n);
        return;
    }
    const u64 base = lower ? edge - n : edge;
    const u32* rel = reinterpret_cast<const u32*>(buf + L.off_rel);
    if (t <= L.plane) start[v0 + t] = (u32)(base + rel[t]);
    if (t < n) {
        const u64 i = base + t;
        const double* pos = reinterpret_cast<const double*>(buf + L.off_pos);
        const double* cen = reinterpret_cast<const double*>(buf + L.off_cen);
        for (int w = 0; w < L.W; ++w) A.pos[(u64)w * cap + i] = pos[(u64)w * L.gcap + t];
        for (int w = 0; w < 4; ++w) A.cen[(u64)w * cap + i] = cen[(u64)w * L.gcap + t];
    }
}

struct RodMigLayout {
    u64 mcap, bytes;
    u64 off_dyn;  // 6 arrays x W planes of mcap doubles: pos, vel, h1x, h1v, h0x, h0v
    u64 off_len, off_growth, off_cen, off_ident, off_parent, off_sortkey, off_gvox, off_hlen;
    int W;
};
struct RodMigParams {
    RodState A, B;  // A: identifiers (pre-step slots), B: integrated state
    u64 cap;
    u32* newkey;
    u64* sortkey;
    u32 *rank_arb, *count_new;
    u64 first, n;  // export: sources [first, first + n); import: arrivals go to [first, first + n)
    u32 lo_end, hi_begin;
    u64 vox_base, nvox;
    RodMigLayout L;
    unsigned char* buf[2];
    u64* counter;  // [2]
    int* err;
};
__global__ void k_rod_mig_export(const RodMigParams M) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        if (M.buf[0]) reinterpret_cast<u64*>(M.buf[0])[1] = (u64)M.L.W;
        if (M.buf[1]) reinterpret_cast<u64*>(M.buf[1])[1] = (u64)M.L.W;
    }
    if (t >= M.n) return;
    const u64 i = M.first + t;
    const u32 k = M.newkey[i];
    int side;
    if (k < M.lo_end) side = 0;
    else if (k >= M.hi_begin && k != CRB_KEY_GONE) side = 1;
    else return;
    const u64 slot = atomicAdd(&M.counter[side], 1ull);
    M.newkey[i] = CRB_KEY_GONE;
    const u64 cap = M.cap;
    if (slot >= M.L.mcap) {
        latch_error(M.err, CRB_ERR_CAPACITY, M.A.ident[cap + i]);
        return;
    }
    unsigned char* buf = M.buf[side];
    const RodMigLayout& L = M.L;
    const u64 mc = L.mcap;
    double* dyn = reinterpret_cast<double*>(buf + L.off_dyn);
    const double* src[6] = {M.B.pos, M.B.vel, M.B.h1x, M.B.h1v, M.B.h0x, M.B.h0v};
    for (int a = 0; a < 6; ++a)
        for (int w = 0; w < L.W; ++w) dyn[((u64)a * L.W + w) * mc + slot] = src[a][(u64)w * cap + i];
    reinterpret_cast<double*>(buf + L.off_len)[slot] = M.B.len[i];
    reinterpret_cast<double*>(buf + L.off_growth)[slot] = M.B.growth[i];
    for (int w = 0; w < 4; ++w) reinterpret_cast<double*>(buf + L.off_cen)[(u64)w * mc + slot] = M.B.cen[(u64)w * cap + i];
    for (int w = 0; w < 3; ++w) {
        reinterpret_cast<u64*>(buf + L.off_ident)[(u64)w * mc + slot] = M.A.ident[(u64)w * cap + i];
        reinterpret_cast<u64*>(buf + L.off_parent)[(u64)w * mc + slot] = M.A.parent[(u64)w * cap + i];
    }
    reinterpret_cast<u64*>(buf + L.off_sortkey)[slot] = M.sortkey[i];
    reinterpret_cast<u64*>(buf + L.off_gvox)[slot] = M.vox_base + k;
    (buf + L.off_hlen)[slot] = M.B.hlen[i];
}
// arrivals of ONE face -> unsorted source slots [first, first + n) of B (identifiers into A at the same slots)
__global__ void k_rod_mig_import(const RodMigParams M, int side) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M.n) return;
    const unsigned char* buf = M.buf[side];
    const u64 i = M.first + t;
    if (reinterpret_cast<const u64*>(buf)[1] != (u64)M.L.W) {
        if (t == 0) latch_error(M.err, CRB_ERR_STATE, reinterpret_cast<const u64*>(buf)[1]);
        M.newkey[i] = CRB_KEY_GONE;
        return;
    }
    const RodMigLayout& L = M.L;
    const u64 mc = L.mcap, cap = M.cap;
    const double* dyn = reinterpret_cast<const double*>(buf + L.off_dyn);
    double* dst[6] = {M.B.pos, M.B.vel, M.B.h1x, M.B.h1v, M.B.h0x, M.B.h0v};
    for (int a = 0; a < 6; ++a)
        for (int w = 0; w < L.W; ++w) dst[a][(u64)w * cap + i] = dyn[((u64)a * L.W + w) * mc + t];
    M.B.len[i] = reinterpret_cast<const double*>(buf + L.off_len)[t];
    M.B.growth[i] = reinterpret_cast<const double*>(buf + L.off_growth)[t];
    for (int w = 0; w < 4; ++w) M.B.cen[(u64)w * cap + i] = reinterpret_cast<const double*>(buf + L.off_cen)[(u64)w * mc + t];
    for (int w = 0; w < 3; ++w) {
        M.A.ident[(u64)w * cap + i] = reinterpret_cast<const u64*>(buf + L.off_ident)[(u64)w * mc + t];
        M.A.parent[(u64)w * cap + i] = reinterpret_cast<const u64*>(buf + L.off_parent)[(u64)w * mc + t];
    }
    M.sortkey[i] = reinterpret_cast<const u64*>(buf + L.off_sortkey)[t];
    M.B.hlen[i] = (buf + L.off_hlen)[t];
    const u64 gv = reinterpret_cast<const u64*>(buf + L.off_gvox)[t];
    if (gv < M.vox_base || gv - M.vox_base >= M.nvox) {
        latch_error(M.err, CRB_ERR_INDEX, M.A.ident[cap + i]);
        M.newkey[i] = CRB_KEY_GONE;
        return;
    }
    const u32 nk = (u32)(gv - M.vox_base);
    M.newkey[i] = nk;
    M.rank_arb[i] = atomicAdd(&M.count_new[nk], 1u);
}

}  // namespace crb
