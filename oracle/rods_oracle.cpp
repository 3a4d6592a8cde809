// oracle/rods_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of one time step of cellular_raza's `chili` backend for the bacterial-rods agent
// (cellular_raza-examples/bacterial_rods/src/main.rs) with
//   aspects [Mechanics, Interaction, Cycle, NeighborSensing]
// i.e. RodMechanics + RodInteraction<MorsePotential> on a CartesianCuboidRods domain, neighbour
// counting -> growth rate, growth of the spring length and division (dynamic agent count).
// Iteration and accumulation order is chili's (SURVEY.md §8a pseudocode and rows A6d, A7, A9, A11);
// every function cites the reference file:line it follows (paths relative to /root/reference).
//
// PARITY STATUS: the Rust reference cannot be built here (no cargo/rustc).  Pinned by the
// known-answer tests the reference holds for this path (tests/test_oracle_rods.py): nearest point
// doctest (interaction.rs:478-496) and exact cases (:820-851), rod division doctest
// (bacterial_rods.rs:560-585).  Segment/curvature forces, the pair weights and the cycle ordering have
// no numeric expectations in the reference: PARITY UNPINNED there (defined by the source text).
// Third-party arithmetic restated from its published behaviour: nalgebra 0.34 (`norm`, `normalize`,
// `angle`, `row_sum`, `dot`), f64::{acos, tan, exp, sqrt} via libm.
//
// Only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke() may load this
// library.  The product (cellular_raza_b200/) never does.
//
// Build: g++ -O2 -ffp-contract=off -pthread -shared -fPIC  (see oracle/Makefile)
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <thread>

#include "../include/crb.h"
#include "chili_common.hpp"

using namespace chili;

namespace {

constexpr int MAXV = CRB_MAX_ROD_VERTICES;

// mech row of CRB_MECH_ROD (include/crb.h)
enum { MP_DIFF = 0, MP_TENSION, MP_RIGIDITY, MP_LENGTH, MP_DAMPING, MP_GROWTH, MP_THRESHOLD, MP_GROWTH_BASE };
// inter row of CRB_INT_MORSE
enum { IP_RADIUS = 0, IP_STIFFNESS, IP_CUTOFF, IP_STRENGTH };

// CellIdentifier (cellular_raza-concepts/src/cell.rs:18-25); kind numbering = declaration order
struct Ident {
    uint64_t kind = 1;  // 0 Division(voxel, counter), 1 Initial(n), 2 Inserted
    uint64_t a = 0, b = 0;
};
constexpr uint64_t kNoParent = ~0ull;

struct Mat {
    double m[MAXV][MAXD];
};

struct Rod {
    Mat pos, vel;
    double mp[CRB_MAX_MECH_PARAMS] = {0};
    double ip[CRB_MAX_INTER_PARAMS] = {0};
    Ident id;
    bool has_parent = false;
    Ident parent;
    // AuxStorage: mechanics ring buffers (chili/aux_storage.rs:41-46), current force, neighbour
    // accumulator (:322-336) and cycle events
    Mat ring_dx[2], ring_dv[2];
    uint32_t ring_len = 0;
    Mat force;
    uint64_t neighbors = 0;
    bool ev_division = false;
    double abs_sum = 0;  // not chili: sum of |half-force contributions|_inf (tolerance metric)
};

struct PosInfo {
    uint64_t index_sender, index_receiver;
    Mat pos, vel;
    double radius;
    uint64_t cell_index_in_vector;
    uint64_t src_sub;
};
struct ForceInfo {
    Mat force;
    double abs_sum;
    uint64_t cell_index_in_vector, index_sender;
};
struct SendCell {
    uint64_t old_voxel;
    Rod cell;
};

struct Oracle {
    crb_config cfg;
    Domain dom;
    SubdomainMap smap;
    int D = 3, V = 5;
    std::vector<std::vector<Rod>> vox;
    std::vector<uint64_t> id_counter;  // Voxel::id_counter (chili/datastructures.rs:176)
    uint64_t iteration = 0;
    uint64_t n_divisions = 0;
    int n_threads = 1;
    // shared noise table: z[(step * n_ids + row) * D V + c V + v] standard normal numbers
    std::vector<double> noise;
    uint64_t noise_steps = 0, noise_ids = 0, noise_cursor = 0;
    std::vector<std::vector<std::vector<PosInfo>>> q_pos;
    std::vector<std::vector<std::vector<ForceInfo>>> q_force;
    std::vector<std::vector<std::vector<SendCell>>> q_cell;
};

thread_local std::string g_err;

template <class F>
void parallel_for(int64_t n, int n_threads, F&& fn) {
    if (n_threads <= 1 || n <= 1) {
        for (int64_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n) break;
            fn(i);
        }
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<int64_t>(n_threads, n);
    for (int t = 1; t < nt; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

// nalgebra: dot of two D-vectors = a0*b0 + a1*b1 (+ a2*b2), left to right (base/blas.rs dotx)
inline double dotD(const double* a, const double* b, int D) {
    double s = a[0] * b[0];
    for (int i = 1; i < D; ++i) s = s + a[i] * b[i];
    return s;
}
inline double normD(const double* z, int D) { return std::sqrt(dotD(z, z, D)); }

// nearest_point_from_point_to_line (cell_building_blocks/interaction.rs:497-509)
inline void nearest_point(const double* p, const double* a, const double* b, int D, double* dist,
                          double* nearest, double* t_out) {
    double ab[MAXD] = {0, 0, 0}, ap[MAXD] = {0, 0, 0};
    for (int i = 0; i < D; ++i) {
        ab[i] = b[i] - a[i];
        ap[i] = p[i] - a[i];
    }
    double t = dotD(ab, ap, D) / dotD(ab, ab, D);
    // f64::clamp: NaN stays NaN
    if (t < 0.0) t = 0.0;
    if (t > 1.0) t = 1.0;
    double z[MAXD] = {0, 0, 0};
    for (int i = 0; i < D; ++i) {
        nearest[i] = a[i] * (1.0 - t) + b[i] * t;
        z[i] = p[i] - nearest[i];
    }
    *dist = normD(z, D);
    *t_out = t;
}

// calculate_morse_interaction (interaction.rs:151-182)
inline void morse(const double* own_pos, const double* ext_pos, double own_radius, double ext_radius,
                  double cutoff, double strength, double stiffness, int D, double* f_own,
                  double* f_ext) {
    double z[MAXD] = {0, 0, 0};
    for (int i = 0; i < D; ++i) z[i] = own_pos[i] - ext_pos[i];
    const double dist = normD(z, D);
    if (dist > cutoff || dist == 0.0) {
        for (int i = 0; i < D; ++i) f_own[i] = f_ext[i] = 0.0;
        return;
    }
    const double r = own_radius + ext_radius;
    const double e = std::exp(-stiffness * (dist - r));
    const double force = -2.0 * strength * stiffness * e * (1.0 - e);
    for (int i = 0; i < D; ++i) {
        const double dir = z[i] / dist;
        f_own[i] = dir * force;
        f_ext[i] = -dir * force;
    }
}

// RodInteraction::calculate_force_between (cell_building_blocks/bacterial_rods.rs:292-336)
void rod_force_between(const Oracle& o, const Rod& own, const Mat& own_pos, const Mat& ext_pos,
                       double ext_radius, Mat* force_own, Mat* force_ext) {
    const int D = o.D, V = o.V;
    std::memset(force_own, 0, sizeof(Mat));
    std::memset(force_ext, 0, sizeof(Mat));
    for (int i = 0; i < V; ++i) {
        const double* p1 = own_pos.m[i];
        for (int j = 0; j + 1 < V; ++j) {
            double dist, nearest[MAXD], rel;
            nearest_point(p1, ext_pos.m[j], ext_pos.m[j + 1], D, &dist, nearest, &rel);
            double f_own[MAXD], f_ext[MAXD];
            morse(p1, nearest, own.ip[IP_RADIUS], ext_radius, own.ip[IP_CUTOFF], own.ip[IP_STRENGTH],
                  own.ip[IP_STIFFNESS], D, f_own, f_ext);
            const int j1 = (j + 1) % V;
            for (int c = 0; c < D; ++c) {
                force_own->m[i][c] += f_own[c];
                force_ext->m[j][c] += f_ext[c] * (1.0 - rel);
                force_ext->m[j1][c] += f_ext[c] * rel;
            }
        }
    }
}

// NeighborSensing::accumulate_information (examples/bacterial_rods/src/main.rs:48-64)
inline void accumulate_information(const Oracle& o, const Rod& own, const Mat& own_pos,
                                   const Mat& ext_pos, uint64_t* acc) {
    for (int i = 0; i < o.V; ++i)
        for (int j = 0; j < o.V; ++j) {
            double z[MAXD] = {0, 0, 0};
            for (int c = 0; c < o.D; ++c) z[c] = own_pos.m[i][c] - ext_pos.m[j][c];
            if (normD(z, o.D) < 2.0 * own.ip[IP_RADIUS]) *acc += 1;
        }
}

inline double mat_absmax(const Mat& f, int V, int D) {
    double m = 0;
    for (int v = 0; v < V; ++v)
        for (int c = 0; c < D; ++c) m = std::max(m, std::fabs(f.m[v][c]));
    return m;
}
// aux_storage.add_force(f.xa(one_half))
inline void add_half_force(const Oracle& o, Rod& r, const Mat& f) {
    for (int v = 0; v < o.V; ++v)
        for (int c = 0; c < o.D; ++c) r.force.m[v][c] += xa(f.m[v][c], 0.5);
    r.abs_sum += 0.5 * mat_absmax(f, o.V, o.D);
}

// Voxel::calculate_force_between_cells_external (chili/update_mechanics.rs:65-108)
void force_between_cells_external(const Oracle& o, std::vector<Rod>& cells, const Mat& ext_pos,
                                  double ext_radius, Mat* ret, double* ret_abs, bool* has_force) {
    *has_force = false;
    for (Rod& cell : cells) {
        Mat f1, f2;
        rod_force_between(o, cell, cell.pos, ext_pos, ext_radius, &f1, &f2);
        add_half_force(o, cell, f1);
        if (*has_force) {
            for (int v = 0; v < o.V; ++v)
                for (int c = 0; c < o.D; ++c) ret->m[v][c] = xapy(f2.m[v][c], 0.5, ret->m[v][c]);
        } else {
            for (int v = 0; v < o.V; ++v)
                for (int c = 0; c < o.D; ++c) ret->m[v][c] = xa(f2.m[v][c], 0.5);
            *has_force = true;
            *ret_abs = 0;
        }
        *ret_abs += 0.5 * mat_absmax(f2, o.V, o.D);
        // neighbor_sensing_func(aux_storage, &cell, &cell.pos(), &ext_pos, &ext_inf)  (:105)
        accumulate_information(o, cell, cell.pos, ext_pos, &cell.neighbors);
    }
}

// update_mechanics_interaction_step_1 (chili/update_mechanics.rs:153-238)
void step_1(Oracle& o, uint64_t s) {
    std::vector<uint64_t> neighbors;
    const uint64_t v_lo = o.smap.first_voxel(s), v_hi = o.smap.first_voxel(s + 1);
    for (uint64_t v = v_lo; v < v_hi; ++v) {
        std::vector<Rod>& cells = o.vox[v];
        const size_t n_cells = cells.size();
        if (n_cells == 0) continue;
        neighbor_plain_indices(o.dom, v, &neighbors);
        for (size_t n = 0; n < n_cells; ++n) {
            const Mat p1 = cells[n].pos;
            const Mat v1 = cells[n].vel;
            const double r1 = cells[n].ip[IP_RADIUS];
            for (size_t m = n + 1; m < n_cells; ++m) {
                Rod& c1 = cells[n];
                Rod& c2 = cells[m];
                Mat fa, fb;
                rod_force_between(o, c1, p1, c2.pos, c2.ip[IP_RADIUS], &fa, &fb);
                add_half_force(o, c1, fa);
                add_half_force(o, c2, fb);
                Mat fb2, fa2;
                rod_force_between(o, c2, c2.pos, p1, r1, &fb2, &fa2);
                add_half_force(o, c1, fa2);
                add_half_force(o, c2, fb2);
                accumulate_information(o, c1, p1, c2.pos, &c1.neighbors);  // (:201-202)
                accumulate_information(o, c2, c2.pos, p1, &c2.neighbors);
            }
            bool have = false;
            Mat force;
            std::memset(&force, 0, sizeof(force));
            double force_abs = 0;
            for (uint64_t w : neighbors) {
                const uint64_t sw = o.smap.sub_of(w);
                if (sw == s) {
                    Mat f;
                    double fabs_ = 0;
                    bool has = false;
                    force_between_cells_external(o, o.vox[w], p1, r1, &f, &fabs_, &has);
                    if (has) {
                        if (have) {
                            for (int a = 0; a < o.V; ++a)
                                for (int c = 0; c < o.D; ++c)
                                    force.m[a][c] = xapy(f.m[a][c], 1.0, force.m[a][c]);
                        } else {
                            force = f;
                            have = true;
                        }
                        force_abs += fabs_;
                    }
                } else {
                    PosInfo pi;
                    pi.index_sender = v;
                    pi.index_receiver = w;
                    pi.pos = p1;
                    pi.vel = v1;
                    pi.radius = r1;
                    pi.cell_index_in_vector = n;
                    pi.src_sub = s;
                    o.q_pos[s][sw].push_back(pi);
                }
            }
            // chili as written drops `force` (update_mechanics.rs:206-234)
            if (o.cfg.pair_weight_mode == CRB_WEIGHT_PHYSICAL && have) {
                for (int a = 0; a < o.V; ++a)
                    for (int c = 0; c < o.D; ++c) cells[n].force.m[a][c] += force.m[a][c];
                cells[n].abs_sum += force_abs;
            }
        }
    }
}

// update_mechanics_interaction_step_2 (chili/update_mechanics.rs:285-340)
void step_2(Oracle& o, uint64_t s) {
    std::vector<PosInfo> received;
    for (uint64_t src = 0; src < o.smap.n_sub; ++src) {
        auto& q = o.q_pos[src][s];
        received.insert(received.end(), q.begin(), q.end());
        q.clear();
    }
    std::stable_sort(received.begin(), received.end(), [](const PosInfo& a, const PosInfo& b) {
        return a.index_sender < b.index_sender;
    });
    for (const PosInfo& pi : received) {
        ForceInfo fi;
        bool has = false;
        fi.abs_sum = 0;
        force_between_cells_external(o, o.vox[pi.index_receiver], pi.pos, pi.radius, &fi.force,
                                     &fi.abs_sum, &has);
        if (has) {
            fi.cell_index_in_vector = pi.cell_index_in_vector;
            fi.index_sender = pi.index_sender;
            o.q_force[s][pi.src_sub].push_back(fi);
        }
    }
}

// update_mechanics_interaction_step_3 (chili/update_mechanics.rs:344-382)
void step_3(Oracle& o, uint64_t s) {
    std::vector<ForceInfo> received;
    for (uint64_t src = 0; src < o.smap.n_sub; ++src) {
        auto& q = o.q_force[src][s];
        received.insert(received.end(), q.begin(), q.end());
        q.clear();
    }
    std::stable_sort(received.begin(), received.end(), [](const ForceInfo& a, const ForceInfo& b) {
        return a.index_sender < b.index_sender;
    });
    for (const ForceInfo& fi : received) {
        Rod& c = o.vox[fi.index_sender][fi.cell_index_in_vector];
        for (int a = 0; a < o.V; ++a)
            for (int k = 0; k < o.D; ++k) c.force.m[a][k] += fi.force.m[a][k];
        c.abs_sum += fi.abs_sum;
    }
}

// SubDomainBox::calculate_custom_domain_force (chili/update_mechanics.rs:248-274), appended to step 1 when
// the aspects contain DomainForce (run_sim.rs:277-278), with the SubDomainForce of CartesianSubDomainRods
// (bacterial_rods.rs:433-477): force = rows (0, 0, -gel_pressure); for every vertex with |v| != 0 that lies
// within surface_friction_distance of domain_min or domain_max in ANY component:
// row += -(v / |v|) * gel_pressure * surface_friction; then aux.add_force(force).
// cfg.domain_force_strength = {gel_pressure, surface_friction, surface_friction_distance}.
void domain_force_step(Oracle& o, uint64_t s) {
    if (o.cfg.domain_force_kind != CRB_DOMAIN_FORCE_ROD_GEL) return;
    const double gel = o.cfg.domain_force_strength[0], fric = o.cfg.domain_force_strength[1],
                 dist = o.cfg.domain_force_strength[2];
    const uint64_t v_lo = o.smap.first_voxel(s), v_hi = o.smap.first_voxel(s + 1);
    for (uint64_t v = v_lo; v < v_hi; ++v)
        for (Rod& r : o.vox[v]) {
            Mat f;
            for (int a = 0; a < o.V; ++a) {
                for (int c = 0; c < o.D; ++c) f.m[a][c] = (c == 2) ? -gel : 0.0;
                bool near = false;
                for (int c = 0; c < o.D; ++c) {
                    near = near || std::fabs(r.pos.m[a][c] - o.cfg.domain_min[c]) <= dist;
                    near = near || std::fabs(r.pos.m[a][c] - o.cfg.domain_max[c]) <= dist;
                }
                const double q = normD(r.vel.m[a], o.D);
                if (q != 0.0 && near)
                    for (int c = 0; c < o.D; ++c) {
                        const double dir = r.vel.m[a][c] / q;
                        f.m[a][c] += ((-dir) * gel) * fric;
                    }
            }
            for (int a = 0; a < o.V; ++a)
                for (int c = 0; c < o.D; ++c) r.force.m[a][c] += f.m[a][c];
            r.abs_sum += mat_absmax(f, o.V, o.D);
        }
}

void run_forces(Oracle& o) {
    const int64_t S = (int64_t)o.smap.n_sub;
    parallel_for(S, o.n_threads, [&](int64_t s) {
        step_1(o, (uint64_t)s);
        domain_force_step(o, (uint64_t)s);
    });
    parallel_for(S, o.n_threads, [&](int64_t s) { step_2(o, (uint64_t)s); });
    parallel_for(S, o.n_threads, [&](int64_t s) { step_3(o, (uint64_t)s); });
}

// RodMechanics::calculate_increment (bacterial_rods.rs:157-213)
void rod_increment(int D, int V, const Mat& pos, const Mat& vel, const double* mp, const Mat& force,
                   Mat* dx, Mat* dv) {
    const double one_half = 1.0 / (1.0 + 1.0), two = 1.0 + 1.0;
    Mat total = force;
    double dist[MAXV][MAXD];
    for (int i = 0; i + 1 < V; ++i)
        for (int c = 0; c < D; ++c) dist[i][c] = pos.m[i][c] - pos.m[i + 1][c];
    for (int i = 0; i + 1 < V; ++i) {
        const double nrm = normD(dist[i], D);
        if (nrm != 0.0) {
            for (int c = 0; c < D; ++c) {
                const double dir = dist[i][c] / nrm;  // normalize(): unscale by the norm
                const double fi = -dir * mp[MP_TENSION] * (nrm - mp[MP_LENGTH]);
                total.m[i][c] += fi * one_half;
                total.m[i + 1][c] += -fi * one_half;
            }
        }
    }
    for (int i = 0; i + 2 < V; ++i) {
        const double* c1 = dist[i];
        const double* c2 = dist[i + 1];
        // nalgebra Matrix::angle: acos(clamp(dot / (n1 * n2), -1, 1)), 0 if a norm is zero
        const double prod = dotD(c1, c2, D);
        const double n1 = normD(c1, D), n2 = normD(c2, D);
        double angle = 0.0;
        if (!(n1 == 0.0 || n2 == 0.0)) {
            double cang = prod / (n1 * n2);
            if (cang < -1.0) cang = -1.0;
            if (cang > 1.0) cang = 1.0;
            angle = std::acos(cang);
        }
        double fd[MAXD];
        for (int c = 0; c < D; ++c) fd[c] = c1[c] - c2[c];
        const double nfd = normD(fd, D);
        const double tn = std::tan(one_half * angle);
        for (int c = 0; c < D; ++c) {
            const double dir = nfd != 0.0 ? fd[c] / nfd : fd[c];
            const double f = dir * two * mp[MP_RIGIDITY] * tn;
            total.m[i][c] += -f * one_half;
            total.m[i + 1][c] += f;
            total.m[i + 2][c] += -f * one_half;
        }
    }
    for (int v = 0; v < V; ++v)
        for (int c = 0; c < D; ++c) {
            total.m[v][c] -= vel.m[v][c] * mp[MP_DAMPING];
            dx->m[v][c] = vel.m[v][c];
            dv->m[v][c] = total.m[v][c];
        }
}

// row of the noise table of a rod: Initial(k) rods use k, rods born by division fold their Division(voxel, counter)
// identifier (the same rule as rod_noise_key of the device code; a contract of the shared table, not of chili)
uint64_t rod_noise_key(const Ident& id) {
    return id.kind == 1 ? id.a : ((id.a * 0x9E3779B97F4A7C15ull + id.b) * 0xBF58476D1CE4E5B9ull) >> 11;
}

// local_mechanics_update -> Adams-Bashforth wrapper (chili/solvers.rs:156-300).  Random contribution of
// RodMechanics (bacterial_rods.rs:215-240): dpos = Matrix::from_distribution(nrows, Normal(0, sqrt(dt)), rng)
// * 2^-2 * diffusion_constant / dt, dvel = 0; nalgebra fills the matrix column-major, so the k-th number drawn goes
// to row k % nrows, column k / nrows; Normal samples mean + std_dev * z.  The ChaCha8 stream is replaced by the shared
// table (noise_mode TABLE); noise_mode NONE gives dpos = 0.
void local_mechanics_update(const Oracle& o, Rod& r, double dt) {
    const int D = o.D, V = o.V;
    const uint32_t N = o.cfg.solver_order;
    Mat rx;
    std::memset(&rx, 0, sizeof(rx));
    if (o.cfg.noise_mode == CRB_NOISE_TABLE && r.mp[MP_DIFF] != 0.0 && o.noise_ids) {
        const uint64_t row = rod_noise_key(r.id) % o.noise_ids;
        const double* z = o.noise.data() + (o.noise_cursor * o.noise_ids + row) * (uint64_t)(D * V);
        const double sd = std::sqrt(dt);
        for (int v = 0; v < V; ++v)
            for (int c = 0; c < D; ++c) {
                const double m = 0.0 + sd * z[c * V + v];
                rx.m[v][c] = m * 0.25 * r.mp[MP_DIFF] / dt;
            }
    }
    const Mat force = r.force;
    std::memset(&r.force, 0, sizeof(Mat));
    r.abs_sum = 0;
    Mat dx, dv;
    rod_increment(D, V, r.pos, r.vel, r.mp, force, &dx, &dv);
    // push before read
    if (N > 0) {
        if (r.ring_len < N) {
            r.ring_dx[r.ring_len] = dx;
            r.ring_dv[r.ring_len] = dv;
            ++r.ring_len;
        } else {
            for (uint32_t k = 1; k < N; ++k) {
                r.ring_dx[k - 1] = r.ring_dx[k];
                r.ring_dv[k - 1] = r.ring_dv[k];
            }
            r.ring_dx[N - 1] = dx;
            r.ring_dv[N - 1] = dv;
        }
    }
    const uint32_t n_prev = N == 0 ? 0 : r.ring_len;
    for (int v = 0; v < V; ++v)
        for (int c = 0; c < D; ++c) {
            const double x = r.pos.m[v][c], w = r.vel.m[v][c];
            double nx, nw;
            if (N == 2 && n_prev == 2) {
                nx = helper_ab3(x, dx.m[v][c], r.ring_dx[0].m[v][c], r.ring_dx[1].m[v][c], dt, rx.m[v][c]);
                nw = helper_ab3(w, dv.m[v][c], r.ring_dv[0].m[v][c], r.ring_dv[1].m[v][c], dt, 0.0);
            } else if (n_prev == 1) {
                nx = helper_ab2(x, dx.m[v][c], r.ring_dx[0].m[v][c], dt, rx.m[v][c]);
                nw = helper_ab2(w, dv.m[v][c], r.ring_dv[0].m[v][c], dt, 0.0);
            } else {
                nx = helper_euler(x, dx.m[v][c], dt, rx.m[v][c]);
                nw = helper_euler(w, dv.m[v][c], dt, 0.0);
            }
            r.pos.m[v][c] = nx;
            r.vel.m[v][c] = nw;
        }
}

// local_interaction_react_to_neighbors (chili/update_mechanics.rs:562-577) ->
// react_to_neighbors (examples/bacterial_rods/src/main.rs:66-73), clear_accumulator (:75-77)
inline void react_to_neighbors(Rod& r) {
    const double base = r.mp[MP_GROWTH_BASE];
    if (r.neighbors > 0)
        r.mp[MP_GROWTH] = std::max(base * (8.0 - (double)r.neighbors) / 8.0, 0.0);
    else
        r.mp[MP_GROWTH] = base;
    r.neighbors = 0;
}

// local_cycle_update (chili/update_cycle.rs:103-127) -> Cycle::update_cycle (main.rs:87-98)
inline void cycle_update(Rod& r, double dt) {
    r.mp[MP_LENGTH] += r.mp[MP_GROWTH] * dt;
    if (r.mp[MP_LENGTH] > r.mp[MP_THRESHOLD]) r.ev_division = true;
}

// RodMechanics::divide (bacterial_rods.rs:586-663).  Returns false on DivisionError.
bool rod_divide(int D, int V, Mat* c1_pos, double* spring_length, double radius, Mat* c2_pos) {
    const Mat pos = *c1_pos;
    *c2_pos = pos;
    const double two = 1.0 + 1.0, one_half = 1.0 / two;
    const double div_factor = one_half - radius / ((double)V * *spring_length);
    *spring_length = div_factor * *spring_length;
    double segments[MAXV];
    const int ns = V - 1;
    for (int i = 0; i < ns; ++i) {
        double z[MAXD] = {0, 0, 0};
        for (int c = 0; c < D; ++c) z[c] = pos.m[i][c] - pos.m[i + 1][c];
        segments[i] = normD(z, D);
    }
    auto sum_take = [&](int n) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += segments[i];
        return s;
    };
    auto sum_take_rev = [&](int n) {
        double s = 0.0;
        for (int i = 0; i < n && i < ns; ++i) s += segments[ns - 1 - i];
        return s;
    };
    const double segment_length = (sum_take(ns) - two * radius) / (double)(V - 1) / two;
    for (int nv = 0; nv < V; ++nv) {
        const double target = (double)nv * segment_length;
        int k = -1;
        for (int n = 0; n < ns; ++n)
            if (sum_take(n) <= target) k = std::max(k, n);
        if (k < 0) return false;
        const double q = (target - sum_take(k)) / segments[k];
        for (int c = 0; c < D; ++c) c1_pos->m[nv][c] = pos.m[k][c] * (1.0 - q) + pos.m[k + 1][c] * q;
        int m = -1;
        for (int n = 0; n < V; ++n)
            if (sum_take_rev(n) <= target) m = std::max(m, n);
        if (m < 0 || V - m - 2 < 0) return false;
        const double p = (target - sum_take_rev(m)) / segments[V - m - 2];
        for (int c = 0; c < D; ++c)
            c2_pos->m[V - nv - 1][c] = pos.m[V - m - 1][c] * (1.0 - p) + pos.m[V - m - 2][c] * p;
    }
    return true;
}

// Voxel::update_cell_cycle_4 (chili/update_cycle.rs:9-69) with Cycle::divide (main.rs:100-108)
bool update_cell_cycle_4(Oracle& o, uint64_t v, uint64_t* n_div) {
    std::vector<Rod>& cells = o.vox[v];
    std::vector<Rod> new_cells;
    for (Rod& r : cells) {
        if (!r.ev_division) continue;
        r.ev_division = false;
        Mat c2_pos;
        if (!rod_divide(o.D, o.V, &r.pos, &r.mp[MP_LENGTH], r.ip[IP_RADIUS], &c2_pos)) return false;
        Rod c2 = r;  // cell.clone() after the mother has been shortened
        c2.pos = c2_pos;
        const Ident parent = r.id;
        o.id_counter[v] += 1;
        r.id.kind = 0;
        r.id.a = v;
        r.id.b = o.id_counter[v];
        r.has_parent = true;
        r.parent = parent;
        c2.has_parent = true;
        c2.parent = parent;
        new_cells.push_back(c2);
        ++*n_div;
    }
    for (Rod& c2 : new_cells) {
        // default aux storage (cellular_raza-core-proc-macro/src/aux_storage.rs default initializer)
        c2.ring_len = 0;
        std::memset(c2.ring_dx, 0, sizeof(c2.ring_dx));
        std::memset(c2.ring_dv, 0, sizeof(c2.ring_dv));
        std::memset(&c2.force, 0, sizeof(Mat));
        c2.neighbors = 0;
        c2.ev_division = false;
        c2.abs_sum = 0;
        o.id_counter[v] += 1;
        c2.id.kind = 0;
        c2.id.a = v;
        c2.id.b = o.id_counter[v];
        cells.push_back(c2);
    }
    return true;
}

// CartesianSubDomainRods::apply_boundary (bacterial_rods.rs:489-533)
bool apply_boundary(const Oracle& o, Rod& r) {
    for (int v = 0; v < o.V; ++v)
        for (int i = 0; i < o.D; ++i) {
            double& p = r.pos.m[v][i];
            double& w = r.vel.m[v][i];
            if (p < o.dom.min[i]) {
                p = o.dom.min[i] * 2.0 - p;
                w = std::fabs(w);
            }
            if (p > o.dom.max[i]) {
                p = o.dom.max[i] * 2.0 - p;
                w = -std::fabs(w);
            }
        }
    for (int v = 0; v < o.V; ++v)
        for (int i = 0; i < o.D; ++i)
            if (r.pos.m[v][i] < o.dom.min[i] || r.pos.m[v][i] > o.dom.max[i]) return false;
    return true;
}

// SortCells::get_voxel_index_of for rods: mean of the vertices, nalgebra row_sum = per column
// fold(0, +) over the rows, then / nrows (bacterial_rods.rs:368-372, 550-554)
inline void rod_mean(const Oracle& o, const Rod& r, double* mean) {
    for (int c = 0; c < o.D; ++c) {
        double s = 0.0;
        for (int v = 0; v < o.V; ++v) s = s + r.pos.m[v][c];
        mean[c] = s / (double)o.V;
    }
}
inline uint64_t voxel_key_or_stay(const Oracle& o, const Rod& r, uint64_t current) {
    double mean[MAXD];
    rod_mean(o, r, mean);
    uint64_t idx[MAXD] = {0, 0, 0};
    if (!get_index_of(o.dom, mean, idx)) return current;
    if (!index_in_domain(o.dom, idx)) return current;
    return o.dom.plain(idx);
}

// sort_cells_in_voxels_step_1 / _2 (chili/update_mechanics.rs:418-513)
void sort_step_1(Oracle& o, uint64_t s) {
    const uint64_t v_lo = o.smap.first_voxel(s), v_hi = o.smap.first_voxel(s + 1);
    std::vector<std::pair<uint64_t, Rod>> find_new_home;
    for (uint64_t v = v_lo; v < v_hi; ++v) {
        std::vector<Rod>& cells = o.vox[v];
        if (cells.empty()) continue;
        std::vector<Rod> stay;
        stay.reserve(cells.size());
        for (Rod& c : cells) {
            if (voxel_key_or_stay(o, c, v) != v)
                find_new_home.emplace_back(v, c);
            else
                stay.push_back(c);
        }
        cells.swap(stay);
    }
    for (auto& vc : find_new_home) {
        const uint64_t nv = voxel_key_or_stay(o, vc.second, vc.first);
        const uint64_t sw = o.smap.sub_of(nv);
        if (sw == s)
            o.vox[nv].push_back(vc.second);
        else
            o.q_cell[s][sw].push_back({vc.first, vc.second});
    }
}
void sort_step_2(Oracle& o, uint64_t s) {
    std::vector<SendCell> received;
    for (uint64_t src = 0; src < o.smap.n_sub; ++src) {
        auto& q = o.q_cell[src][s];
        received.insert(received.end(), q.begin(), q.end());
        q.clear();
    }
    std::stable_sort(received.begin(), received.end(),
                     [](const SendCell& a, const SendCell& b) { return a.old_voxel < b.old_voxel; });
    for (SendCell& sc : received) {
        const uint64_t nv = voxel_key_or_stay(o, sc.cell, sc.old_voxel);
        o.vox[nv].push_back(sc.cell);
    }
}

}  // namespace

// =================================================================================================
extern "C" {

const char* rods_oracle_last_error() { return g_err.c_str(); }

void* rods_oracle_create(const crb_config* cfg) {
    if (cfg->dim < 2 || cfg->dim > 3 || cfg->rod_n_vertices < 2 || cfg->rod_n_vertices > (uint32_t)MAXV) {
        g_err = "rods: dim must be 2 or 3 and 2 <= rod_n_vertices <= 8";
        return nullptr;
    }
    Oracle* o = new Oracle();
    o->cfg = *cfg;
    o->D = (int)cfg->dim;
    o->V = (int)cfg->rod_n_vertices;
    o->dom.dim = o->D;
    for (int i = 0; i < o->D; ++i) {
        o->dom.min[i] = cfg->domain_min[i];
        o->dom.max[i] = cfg->domain_max[i];
        o->dom.dx[i] = cfg->dx[i];
        o->dom.n_voxels[i] = cfg->n_voxels[i];
    }
    const uint64_t nsub = cfg->ref_n_subdomains ? cfg->ref_n_subdomains : 1;
    if (!o->smap.init(o->dom.n_total(), nsub)) {
        g_err = "DecomposeError: Could not find a suiting decomposition";
        delete o;
        return nullptr;
    }
    o->vox.resize(o->dom.n_total());
    o->id_counter.assign(o->dom.n_total(), 0);
    o->q_pos.assign(nsub, std::vector<std::vector<PosInfo>>(nsub));
    o->q_force.assign(nsub, std::vector<std::vector<ForceInfo>>(nsub));
    o->q_cell.assign(nsub, std::vector<std::vector<SendCell>>(nsub));
    return o;
}
void rods_oracle_destroy(void* h) { delete (Oracle*)h; }
void rods_oracle_set_threads(void* h, int n) { ((Oracle*)h)->n_threads = n < 1 ? 1 : n; }

// insert_initial_cells (chili/datastructures.rs:278-299): input order inside a voxel,
// voxel.id_counter += 1 per inserted cell; ids default to Initial(k).
int rods_oracle_set_agents(void* h, uint64_t n, const double* pos, const double* vel,
                           const double* mech_params, const double* inter_params, const uint64_t* ids) {
    Oracle& o = *(Oracle*)h;
    const int D = o.D, V = o.V;
    for (auto& v : o.vox) v.clear();
    std::fill(o.id_counter.begin(), o.id_counter.end(), 0);
    for (uint64_t k = 0; k < n; ++k) {
        Rod r;
        std::memset(&r.pos, 0, sizeof(Mat));
        std::memset(&r.vel, 0, sizeof(Mat));
        std::memset(&r.force, 0, sizeof(Mat));
        std::memset(r.ring_dx, 0, sizeof(r.ring_dx));
        std::memset(r.ring_dv, 0, sizeof(r.ring_dv));
        for (int v = 0; v < V; ++v)
            for (int c = 0; c < D; ++c) {
                r.pos.m[v][c] = pos[(k * V + v) * D + c];
                r.vel.m[v][c] = vel ? vel[(k * V + v) * D + c] : 0.0;
            }
        for (int j = 0; j < 8; ++j) r.mp[j] = mech_params[k * 8 + j];
        for (int j = 0; j < 4; ++j) r.ip[j] = inter_params[k * 4 + j];
        r.id.kind = 1;
        r.id.a = ids ? ids[k] : k;
        r.id.b = 0;
        double mean[MAXD];
        rod_mean(o, r, mean);
        uint64_t idx[MAXD] = {0, 0, 0};
        if (!get_voxel_index_of_raw(o.dom, mean, idx)) {
            g_err = "BoundaryError: agent " + std::to_string(k) + " at or below domain minimum";
            return CRB_ERR_BOUNDARY;
        }
        if (!index_in_domain(o.dom, idx)) {
            g_err = "IndexError: could not find voxel index for agent " + std::to_string(k);
            return CRB_ERR_INDEX;
        }
        const uint64_t p = o.dom.plain(idx);
        o.vox[p].push_back(r);
        o.id_counter[p] += 1;
    }
    return 0;
}

uint64_t rods_oracle_n_agents(void* h) {
    Oracle& o = *(Oracle*)h;
    uint64_t n = 0;
    for (auto& v : o.vox) n += v.size();
    return n;
}
uint64_t rods_oracle_n_divisions(void* h) { return ((Oracle*)h)->n_divisions; }

// one iteration of the generated loop (run_sim.rs:380-403) for
// aspects [Mechanics, Interaction, Cycle, NeighborSensing] (run_sim.rs:240-290)
int rods_oracle_step(void* h, uint64_t n_steps) {
    Oracle& o = *(Oracle*)h;
    const int64_t S = (int64_t)o.smap.n_sub;
    const double dt = o.cfg.dt;
    for (uint64_t it = 0; it < n_steps; ++it) {
        if (o.cfg.noise_mode == CRB_NOISE_TABLE && o.noise_ids && o.noise_cursor >= o.noise_steps) {
            g_err = "noise table exhausted";
            return CRB_ERR_STATE;
        }
        run_forces(o);
        std::atomic<int> ok{1};
        std::atomic<uint64_t> ndiv{0};
        parallel_for(S, o.n_threads, [&](int64_t s) {
            const uint64_t v_lo = o.smap.first_voxel(s), v_hi = o.smap.first_voxel(s + 1);
            // run_local_cell_funcs: mechanics, react_to_neighbors, cycle per cell (run_sim.rs:259-283)
            for (uint64_t v = v_lo; v < v_hi; ++v)
                for (Rod& r : o.vox[v]) {
                    local_mechanics_update(o, r, dt);
                    react_to_neighbors(r);
                    cycle_update(r, dt);
                }
            // step_4: update_cell_cycle_4, apply_boundary, sort_cells_in_voxels_step_1
            uint64_t nd = 0;
            for (uint64_t v = v_lo; v < v_hi; ++v)
                if (!update_cell_cycle_4(o, v, &nd)) ok = -4;
            ndiv += nd;
            for (uint64_t v = v_lo; v < v_hi; ++v)
                for (Rod& r : o.vox[v])
                    if (!apply_boundary(o, r)) ok = -2;
        });
        o.n_divisions = ndiv;
        if (ok == -4) {
            g_err = "DivisionError: Could not calculate new segment lengths";
            return CRB_ERR_DIVISION;
        }
        if (ok == -2) {
            g_err = "BoundaryError: Particle is out of domain";
            return CRB_ERR_BOUNDARY;
        }
        parallel_for(S, o.n_threads, [&](int64_t s) { sort_step_1(o, (uint64_t)s); });
        parallel_for(S, o.n_threads, [&](int64_t s) { sort_step_2(o, (uint64_t)s); });
        ++o.iteration;
        ++o.noise_cursor;
    }
    return 0;
}

int rods_oracle_set_noise(void* h, const double* z, uint64_t n_steps, uint64_t n_ids) {
    Oracle& o = *(Oracle*)h;
    o.noise.assign(z, z + n_steps * n_ids * (uint64_t)(o.D * o.V));
    o.noise_steps = n_steps;
    o.noise_ids = n_ids;
    o.noise_cursor = 0;
    return 0;
}

// forces and neighbour counts after update_mechanics_interaction_step_3 for the current state in
// (voxel, in-voxel) order; the state is left untouched.
int rods_oracle_forces(void* h, double* force, double* abs_sum, uint64_t* neighbors) {
    Oracle& o = *(Oracle*)h;
    run_forces(o);
    uint64_t k = 0;
    for (auto& v : o.vox)
        for (Rod& r : v) {
            for (int a = 0; a < o.V; ++a)
                for (int c = 0; c < o.D; ++c) force[(k * o.V + a) * o.D + c] = r.force.m[a][c];
            std::memset(&r.force, 0, sizeof(Mat));
            if (abs_sum) abs_sum[k] = r.abs_sum;
            r.abs_sum = 0;
            if (neighbors) neighbors[k] = r.neighbors;
            r.neighbors = 0;
            ++k;
        }
    return 0;
}

// ident / parent: 3 u64 per agent (kind, a, b); parent kind = ~0 when None
int rods_oracle_get_agents(void* h, double* pos, double* vel, uint64_t* ident, uint64_t* parent,
                           uint64_t* voxel) {
    Oracle& o = *(Oracle*)h;
    uint64_t k = 0;
    for (uint64_t v = 0; v < o.vox.size(); ++v)
        for (const Rod& r : o.vox[v]) {
            for (int a = 0; a < o.V; ++a)
                for (int c = 0; c < o.D; ++c) {
                    if (pos) pos[(k * o.V + a) * o.D + c] = r.pos.m[a][c];
                    if (vel) vel[(k * o.V + a) * o.D + c] = r.vel.m[a][c];
                }
            if (ident) {
                ident[k * 3 + 0] = r.id.kind;
                ident[k * 3 + 1] = r.id.a;
                ident[k * 3 + 2] = r.id.b;
            }
            if (parent) {
                parent[k * 3 + 0] = r.has_parent ? r.parent.kind : kNoParent;
                parent[k * 3 + 1] = r.has_parent ? r.parent.a : 0;
                parent[k * 3 + 2] = r.has_parent ? r.parent.b : 0;
            }
            if (voxel) voxel[k] = v;
            ++k;
        }
    return 0;
}

int rods_oracle_get_params(void* h, double* mech_params) {
    Oracle& o = *(Oracle*)h;
    uint64_t k = 0;
    for (auto& v : o.vox)
        for (const Rod& r : v) {
            for (int j = 0; j < 8; ++j) mech_params[k * 8 + j] = r.mp[j];
            ++k;
        }
    return 0;
}

// ring buffers oldest-first: 2 * V * D doubles per agent
int rods_oracle_get_aux(void* h, double* hist_dx, double* hist_dv, uint32_t* hist_len) {
    Oracle& o = *(Oracle*)h;
    const int W = o.V * o.D;
    uint64_t k = 0;
    for (auto& v : o.vox)
        for (const Rod& r : v) {
            for (int e = 0; e < 2; ++e)
                for (int a = 0; a < o.V; ++a)
                    for (int c = 0; c < o.D; ++c) {
                        const bool live = e < (int)r.ring_len;
                        hist_dx[(k * 2 + e) * W + a * o.D + c] = live ? r.ring_dx[e].m[a][c] : 0.0;
                        hist_dv[(k * 2 + e) * W + a * o.D + c] = live ? r.ring_dv[e].m[a][c] : 0.0;
                    }
            hist_len[k] = r.ring_len;
            ++k;
        }
    return 0;
}

int rods_oracle_get_voxel_counts(void* h, uint32_t* counts) {
    Oracle& o = *(Oracle*)h;
    for (size_t v = 0; v < o.vox.size(); ++v) counts[v] = (uint32_t)o.vox[v].size();
    return 0;
}

// ordered candidate rod pairs (j in the 3^D voxels around i, j != i)
uint64_t rods_oracle_count_candidate_pairs(void* h) {
    Oracle& o = *(Oracle*)h;
    std::vector<uint64_t> nb;
    uint64_t total = 0;
    for (uint64_t v = 0; v < o.vox.size(); ++v) {
        const uint64_t n = o.vox[v].size();
        if (!n) continue;
        neighbor_plain_indices(o.dom, v, &nb);
        uint64_t c = n - 1;
        for (uint64_t w : nb) c += o.vox[w].size();
        total += n * c;
    }
    return total;
}

// ---- known-answer taps ------------------------------------------------------------------------
void rods_oracle_nearest_point(uint32_t D, const double* p, const double* a, const double* b,
                               double* dist, double* nearest, double* t) {
    nearest_point(p, a, b, (int)D, dist, nearest, t);
}
// pos: V*D doubles (vertex-major); out pos1/pos2: mother / daughter; returns 0 or CRB_ERR_DIVISION
int rods_oracle_divide(uint32_t D, uint32_t V, const double* pos, double spring_length, double radius,
                       double* pos1, double* pos2, double* new_spring_length) {
    Mat c1, c2;
    std::memset(&c1, 0, sizeof(c1));
    for (uint32_t v = 0; v < V; ++v)
        for (uint32_t c = 0; c < D; ++c) c1.m[v][c] = pos[v * D + c];
    double l = spring_length;
    if (!rod_divide((int)D, (int)V, &c1, &l, radius, &c2)) return CRB_ERR_DIVISION;
    for (uint32_t v = 0; v < V; ++v)
        for (uint32_t c = 0; c < D; ++c) {
            pos1[v * D + c] = c1.m[v][c];
            pos2[v * D + c] = c2.m[v][c];
        }
    *new_spring_length = l;
    return 0;
}
void rods_oracle_increment(uint32_t D, uint32_t V, const double* pos, const double* vel,
                           const double* mech_row, const double* force, double* dx, double* dv) {
    Mat p, w, f, ox, ov;
    std::memset(&p, 0, sizeof(p));
    std::memset(&w, 0, sizeof(w));
    std::memset(&f, 0, sizeof(f));
    for (uint32_t v = 0; v < V; ++v)
        for (uint32_t c = 0; c < D; ++c) {
            p.m[v][c] = pos[v * D + c];
            w.m[v][c] = vel[v * D + c];
            f.m[v][c] = force ? force[v * D + c] : 0.0;
        }
    rod_increment((int)D, (int)V, p, w, mech_row, f, &ox, &ov);
    for (uint32_t v = 0; v < V; ++v)
        for (uint32_t c = 0; c < D; ++c) {
            dx[v * D + c] = ox.m[v][c];
            dv[v * D + c] = ov.m[v][c];
        }
}
// own/ext: V*D doubles; inter_row of own; returns force_own, force_ext (V*D each)
void rods_oracle_pair_force(uint32_t D, uint32_t V, const double* own, const double* ext,
                            const double* inter_row, double ext_radius, double* f_own, double* f_ext) {
    Oracle o;
    o.D = (int)D;
    o.V = (int)V;
    Rod r;
    for (int j = 0; j < 4; ++j) r.ip[j] = inter_row[j];
    Mat a, b, fo, fe;
    std::memset(&a, 0, sizeof(a));
    std::memset(&b, 0, sizeof(b));
    for (uint32_t v = 0; v < V; ++v)
        for (uint32_t c = 0; c < D; ++c) {
            a.m[v][c] = own[v * D + c];
            b.m[v][c] = ext[v * D + c];
        }
    rod_force_between(o, r, a, b, ext_radius, &fo, &fe);
    for (uint32_t v = 0; v < V; ++v)
        for (uint32_t c = 0; c < D; ++c) {
            f_own[v * D + c] = fo.m[v][c];
            f_ext[v * D + c] = fe.m[v][c];
        }
}

}  // extern "C"
