// oracle/chili_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of one time step of cellular_raza's `chili` backend for
// aspects [Mechanics, Interaction] with point agents (NewtonDamped / Brownian / Langevin mechanics,
// cell_sorting species potential / Morse / BoundLennardJones / Mie interactions) on a
// CartesianCuboid in 2 or 3 dimensions.  The iteration and accumulation order is chili's
// (SURVEY.md §8a pseudocode); each function cites the reference file:line it follows.
//
// PARITY STATUS: the Rust reference cannot be built here (no cargo/rustc, no network).  The
// restatement is pinned against every known-answer test the reference holds for this path
// (tests/test_oracle_known_answers.py); pair forces, chili's cross-voxel weights and the
// Adams-Bashforth wrapper recurrence have no numeric expectations anywhere in the reference and
// are therefore PARITY UNPINNED (defined by the source text restated below).
//
// Only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke() may load
// this library.  The product (cellular_raza_b200/) never does.
//
// Build: g++ -O2 -ffp-contract=off -pthread -shared -fPIC  (see oracle/Makefile)
#include <algorithm>
#include <cstdio>
#include <atomic>
#include <functional>
#include <thread>

#include "../include/crb.h"
#include "chili_common.hpp"

using namespace chili;

namespace {

struct Cell {
    double pos[MAXD] = {0, 0, 0}, vel[MAXD] = {0, 0, 0};
    double mp[CRB_MAX_MECH_PARAMS] = {0};
    double ip[CRB_MAX_INTER_PARAMS] = {0};
    uint8_t species = 0;
    uint64_t id = 0;
    // AuxStorageMechanics (chili/aux_storage.rs:41-46)
    Ring<MAXD> ring_dx, ring_dv;
    double force[MAXD] = {0, 0, 0};
    // not part of chili: sum of |half-force contributions|_inf, used by the tolerance metric of
    // SURVEY.md §8d ("parity metrics")
    double abs_sum = 0;
};

// PosInformation / ForceInformation / SendCell (chili/update_mechanics.rs:12-61)
struct PosInfo {
    uint64_t index_sender, index_receiver;
    double pos[MAXD];
    double radius;  // InteractionInformation of the sender
    uint8_t species;
    uint64_t cell_index_in_vector;
    uint64_t src_sub;
};
struct ForceInfo {
    double force[MAXD];
    double abs_sum;
    uint64_t cell_index_in_vector;
    uint64_t index_sender;
};
struct SendCell {
    uint64_t old_voxel;
    Cell cell;
};
struct Migration {
    uint64_t id, old_voxel, new_voxel;
};

struct Oracle {
    crb_config cfg;
    Domain dom;
    SubdomainMap smap;
    int D = 3;
    std::vector<std::vector<Cell>> vox;  // Voxel::cells by VoxelPlainIndex
    std::vector<double> noise;           // z[step][id][d]
    uint64_t noise_steps = 0, noise_ids = 0, noise_cursor = 0;
    uint64_t iteration = 0;
    std::vector<Migration> migrations;
    std::string err;
    int n_threads = 1;
    // queues [src][dst]
    std::vector<std::vector<std::vector<PosInfo>>> q_pos;
    std::vector<std::vector<std::vector<ForceInfo>>> q_force;
    std::vector<std::vector<std::vector<SendCell>>> q_cell;
};

thread_local std::string g_err;

// one worker per emulated chili subdomain at a time (chili: one OS thread per SubDomainBox,
// run_sim.rs:29-59), work-shared over n_threads host threads.
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

// ------------------------------------------------------------------------------------------------
// Interaction::calculate_force_between for the closed model set.  `own` supplies the parameters
// (self), ext_* is the InteractionInformation of the other agent.  Returns false on CalcError.
// ------------------------------------------------------------------------------------------------
inline double norm(const double* z, int D) {
    // nalgebra norm_squared for static vectors: ((x*x + y*y) + z*z), then sqrt
    double s = z[0] * z[0];
    for (int i = 1; i < D; ++i) s = s + z[i] * z[i];
    return std::sqrt(s);
}

// cellular_raza-examples/cell_sorting/src/main.rs:43-97 (sign = +1) and the benchmark copy
// cellular_raza-benchmarks/src/cell_sorting.rs:34-88 (sign = -1: the two returned forces swapped)
bool force_cell_sorting(const Cell& own, const double* own_pos, const double* ext_pos,
                        double ext_radius, uint8_t ext_species, int D, bool bench_sign,
                        double* f_own, double* f_ext) {
    const double cell_radius = own.ip[0], potential_strength = own.ip[1],
                 relative_interaction_range = own.ip[2];
    const double min_relative_distance_to_center = 0.3162277660168379;
    double z[MAXD], dir[MAXD];
    for (int i = 0; i < D; ++i) z[i] = own_pos[i] - ext_pos[i];
    double nz = norm(z, D);
    double r;
    if (!(nz < cell_radius * min_relative_distance_to_center)) {
        r = nz;
        for (int i = 0; i < D; ++i) dir[i] = z[i] / nz;  // normalize() = self / self.norm()
    } else {
        bool equal = true;
        for (int i = 0; i < D; ++i) equal = equal && (own_pos[i] == ext_pos[i]);
        if (equal) {
            for (int i = 0; i < D; ++i) f_own[i] = f_ext[i] = 0.0;
            return true;
        }
        for (int i = 0; i < D; ++i) dir[i] = z[i] / nz;
        r = cell_radius * min_relative_distance_to_center;
    }
    double sigma = r / (cell_radius + ext_radius);
    double bound = 4.0 + 1.0 / sigma;
    double spatial_cutoff =
        (1.0 + rust_signum(relative_interaction_range * (cell_radius + ext_radius) - r)) * 0.5;
    double val = std::pow(1.0 / sigma, 2.0) - std::pow(1.0 / sigma, 4.0);
    val = std::fmax(std::fmin(val, bound), -bound);  // .min(bound).max(-bound)
    double strength = potential_strength * val;
    double sp = std::fmax(strength, 0.0), sn = std::fmin(strength, 0.0);
    for (int i = 0; i < D; ++i) {
        double att = dir[i] * sp * spatial_cutoff;
        double rep = dir[i] * sn * spatial_cutoff;
        double a, b;
        if (ext_species == own.species) {
            a = -rep - att;
            b = rep + att;
        } else {
            a = -rep;
            b = rep;
        }
        if (!bench_sign) {
            f_own[i] = a;
            f_ext[i] = b;
        } else {
            // benchmark copy: Ok((rep + att, -rep - att)) / Ok((rep, -rep))
            f_own[i] = b;
            f_ext[i] = a;
        }
    }
    return true;
}

// calculate_morse_interaction (interaction.rs:151-182) called from MorsePotential (:226-260)
bool force_morse(const Cell& own, const double* own_pos, const double* ext_pos, double ext_radius,
                 int D, double* f_own, double* f_ext) {
    const double own_radius = own.ip[0], a = own.ip[1], cutoff = own.ip[2], s = own.ip[3];
    double z[MAXD];
    for (int i = 0; i < D; ++i) z[i] = own_pos[i] - ext_pos[i];
    double dist = norm(z, D);
    if (dist > cutoff || dist == 0.0) {
        for (int i = 0; i < D; ++i) f_own[i] = f_ext[i] = 0.0;
        return true;
    }
    double r = own_radius + ext_radius;
    double e = std::exp(-a * (dist - r));
    double force = -2.0 * s * a * e * (1.0 - e);
    for (int i = 0; i < D; ++i) {
        double dir = z[i] / dist;
        f_own[i] = dir * force;
        f_ext[i] = -dir * force;
    }
    return true;
}

// BoundLennardJones (interaction.rs:122-138).  No r == 0 guard in the reference: NaN propagates.
bool force_bound_lj(const Cell& own, const double* own_pos, const double* ext_pos, int D,
                    double* f_own, double* f_ext) {
    const double epsilon = own.ip[0], sigma = own.ip[1], bound = own.ip[2], cutoff = own.ip[3];
    double z[MAXD];
    for (int i = 0; i < D; ++i) z[i] = own_pos[i] - ext_pos[i];
    double r = norm(z, D);
    double val = 4.0 * epsilon / r *
                 (12.0 * std::pow(sigma / r, 11.0) - 6.0 * std::pow(sigma / r, 5.0));
    double max = bound / r;
    double q = (cutoff >= r) ? 1.0 : 0.0;
    double mn = std::fmin(max, val);
    for (int i = 0; i < D; ++i) {
        double dir = z[i] / r;
        f_own[i] = -dir * q * mn;
        f_ext[i] = dir * q * mn;
    }
    return true;
}

// MiePotential (interaction.rs:345-389)
bool force_mie(const Cell& own, const double* own_pos, const double* ext_pos, double ext_radius,
               int D, double* f_own, double* f_ext) {
    const double radius = own.ip[0], strength = own.ip[1], bound = own.ip[2], cutoff = own.ip[3],
                 en = own.ip[4], em = own.ip[5];
    double z[MAXD];
    for (int i = 0; i < D; ++i) z[i] = own_pos[i] - ext_pos[i];
    double r = norm(z, D);
    if (r == 0.0) return false;  // CalcError
    if (r > cutoff) {
        for (int i = 0; i < D; ++i) f_own[i] = f_ext[i] = 0.0;
        return true;
    }
    double x = radius + ext_radius;
    double sigma = std::pow(em / en, 1.0 / (en - em)) * x;  // radius_to_sigma_factor() * x
    double mie_constant = en / (en - em) * std::pow(en / em, em / (en - em));
    double potential_part = en / sigma * std::pow(sigma / r, en + 1.0) -
                            em / sigma * std::pow(sigma / r, em + 1.0);
    double force = std::fmin(mie_constant * strength * potential_part, bound);
    for (int i = 0; i < D; ++i) {
        double dir = z[i] / r;
        f_own[i] = dir * force;
        f_ext[i] = -dir * force;
    }
    return true;
}

inline double info_radius(const Oracle& o, const Cell& c) {
    switch (o.cfg.interaction_kind) {
        case CRB_INT_CELL_SORTING:
        case CRB_INT_CELL_SORTING_BENCH:
        case CRB_INT_MORSE:
        case CRB_INT_MIE:
            return c.ip[0];  // get_interaction_information() -> radius
        default:
            return 0.0;  // BoundLennardJones: ()
    }
}

bool calculate_force_between(const Oracle& o, const Cell& own, const double* own_pos,
                             const double* ext_pos, double ext_radius, uint8_t ext_species,
                             double* f_own, double* f_ext) {
    switch (o.cfg.interaction_kind) {
        case CRB_INT_CELL_SORTING:
            return force_cell_sorting(own, own_pos, ext_pos, ext_radius, ext_species, o.D, false,
                                      f_own, f_ext);
        case CRB_INT_CELL_SORTING_BENCH:
            return force_cell_sorting(own, own_pos, ext_pos, ext_radius, ext_species, o.D, true,
                                      f_own, f_ext);
        case CRB_INT_MORSE:
            return force_morse(own, own_pos, ext_pos, ext_radius, o.D, f_own, f_ext);
        case CRB_INT_BOUND_LJ:
            return force_bound_lj(own, own_pos, ext_pos, o.D, f_own, f_ext);
        case CRB_INT_MIE:
            return force_mie(own, own_pos, ext_pos, ext_radius, o.D, f_own, f_ext);
    }
    return false;
}

inline double absmax(const double* f, int D) {
    double m = 0;
    for (int i = 0; i < D; ++i) m = std::fmax(m, std::fabs(f[i]));
    return m;
}

// aux_storage.add_force(f.xa(one_half))  (chili/update_mechanics.rs:95, 193-200)
inline void add_half_force(Cell& c, const double* f, int D) {
    for (int i = 0; i < D; ++i) c.force[i] += xa(f[i], 0.5);
    c.abs_sum += 0.5 * absmax(f, D);
}

// Voxel::calculate_force_between_cells_external (chili/update_mechanics.rs:65-108)
// returns has_force; `ret` = sum of 0.5*f2 accumulated as f2.xapy(0.5, ret)
bool force_between_cells_external(const Oracle& o, std::vector<Cell>& cells, const double* ext_pos,
                                  double ext_radius, uint8_t ext_species, double* ret,
                                  double* ret_abs, bool* has_force) {
    *has_force = false;
    for (Cell& cell : cells) {
        double f1[MAXD], f2[MAXD];
        if (!calculate_force_between(o, cell, cell.pos, ext_pos, ext_radius, ext_species, f1, f2))
            return false;
        add_half_force(cell, f1, o.D);
        if (*has_force) {
            for (int i = 0; i < o.D; ++i) ret[i] = xapy(f2[i], 0.5, ret[i]);
        } else {
            for (int i = 0; i < o.D; ++i) ret[i] = xa(f2[i], 0.5);
            *has_force = true;
            *ret_abs = 0;
        }
        *ret_abs += 0.5 * absmax(f2, o.D);
    }
    return true;
}

// SubDomainBox::update_mechanics_interaction_step_1 (chili/update_mechanics.rs:153-238)
bool step_1(Oracle& o, uint64_t s) {
    const int D = o.D;
    std::vector<uint64_t> neighbors;
    const uint64_t v_lo = o.smap.first_voxel(s), v_hi = o.smap.first_voxel(s + 1);
    for (uint64_t v = v_lo; v < v_hi; ++v) {
        std::vector<Cell>& cells = o.vox[v];
        const size_t n_cells = cells.size();
        if (n_cells == 0) continue;
        neighbor_plain_indices(o.dom, v, &neighbors);
        for (size_t n = 0; n < n_cells; ++n) {
            // p1, v1, i1 are clones taken before the inner loop (:181-183)
            double p1[MAXD];
            std::memcpy(p1, cells[n].pos, sizeof(p1));
            const double r1 = info_radius(o, cells[n]);
            const uint8_t s1 = cells[n].species;
            for (size_t m = n + 1; m < n_cells; ++m) {
                Cell& c1 = cells[n];
                Cell& c2 = cells[m];
                double fa[MAXD], fb[MAXD];
                if (!calculate_force_between(o, c1, p1, c2.pos, info_radius(o, c2), c2.species, fa,
                                             fb))
                    return false;
                add_half_force(c1, fa, D);
                add_half_force(c2, fb, D);
                double fb2[MAXD], fa2[MAXD];
                if (!calculate_force_between(o, c2, c2.pos, p1, r1, s1, fb2, fa2)) return false;
                add_half_force(c1, fa2, D);
                add_half_force(c2, fb2, D);
            }
            bool have = false;
            double force[MAXD] = {0, 0, 0};
            double force_abs = 0;
            for (uint64_t w : neighbors) {
                const uint64_t sw = o.smap.sub_of(w);
                if (sw == s) {
                    double f[MAXD], fabs_ = 0;
                    bool has = false;
                    if (!force_between_cells_external(o, o.vox[w], p1, r1, s1, f, &fabs_, &has))
                        return false;
                    if (has) {
                        if (have) {
                            for (int i = 0; i < D; ++i) force[i] = xapy(f[i], 1.0, force[i]);
                        } else {
                            std::memcpy(force, f, sizeof(force));
                            have = true;
                        }
                        force_abs += fabs_;
                    }
                } else {
                    PosInfo pi;
                    pi.index_sender = v;
                    pi.index_receiver = w;
                    std::memcpy(pi.pos, p1, sizeof(p1));
                    pi.radius = r1;
                    pi.species = s1;
                    pi.cell_index_in_vector = n;
                    pi.src_sub = s;
                    o.q_pos[s][sw].push_back(pi);
                }
            }
            // chili as written drops `force` here (update_mechanics.rs:206-234); the legacy
            // backend adds it (cpu_os_threads/domain_decomposition.rs:969-971).
            if (o.cfg.pair_weight_mode == CRB_WEIGHT_PHYSICAL && have) {
                for (int i = 0; i < D; ++i) cells[n].force[i] += force[i];
                cells[n].abs_sum += force_abs;
            }
        }
    }
    return true;
}

// SubDomainBox::update_mechanics_interaction_step_2 (chili/update_mechanics.rs:285-340)
bool step_2(Oracle& o, uint64_t s) {
    std::vector<PosInfo> received;
    // channel arrival order across senders is arbitrary in chili; per sender it is FIFO.  With
    // determinism=true (run_sim.rs:141) the vector is stably sorted by index_sender, which makes
    // the order unique because one sender voxel lives in exactly one subdomain.
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
        if (!force_between_cells_external(o, o.vox[pi.index_receiver], pi.pos, pi.radius,
                                          pi.species, fi.force, &fi.abs_sum, &has))
            return false;
        if (has) {
            fi.cell_index_in_vector = pi.cell_index_in_vector;
            fi.index_sender = pi.index_sender;
            o.q_force[s][pi.src_sub].push_back(fi);
        }
    }
    return true;
}

// SubDomainBox::update_mechanics_interaction_step_3 (chili/update_mechanics.rs:344-382)
bool step_3(Oracle& o, uint64_t s) {
    std::vector<ForceInfo> received;
    // tie-break for equal index_sender coming from different responders: responder subdomain
    // ascending (chili leaves this to channel arrival order, i.e. it is racy there).
    for (uint64_t src = 0; src < o.smap.n_sub; ++src) {
        auto& q = o.q_force[src][s];
        received.insert(received.end(), q.begin(), q.end());
        q.clear();
    }
    std::stable_sort(received.begin(), received.end(), [](const ForceInfo& a, const ForceInfo& b) {
        return a.index_sender < b.index_sender;
    });
    for (const ForceInfo& fi : received) {
        Cell& c = o.vox[fi.index_sender][fi.cell_index_in_vector];
        for (int i = 0; i < o.D; ++i) c.force[i] += fi.force[i];
        c.abs_sum += fi.abs_sum;
    }
    return true;
}

// Mechanics::calculate_increment (mechanics.rs:158-173, 336-357, 434-463)
void calculate_increment(const Oracle& o, const Cell& c, const double* force, double* dx,
                         double* dv) {
    const int D = o.D;
    switch (o.cfg.mechanics_kind) {
        case CRB_MECH_NEWTON_DAMPED: {
            const double damping = c.mp[0], mass = c.mp[1];
            for (int i = 0; i < D; ++i) {
                dx[i] = c.vel[i];
                dv[i] = force[i] / mass - damping * c.vel[i];
            }
            break;
        }
        case CRB_MECH_BROWNIAN: {
            const double f = c.mp[0] / c.mp[1];  // diffusion_constant / kb_temperature
            for (int i = 0; i < D; ++i) {
                dx[i] = f * force[i];
                dv[i] = 0.0;
            }
            break;
        }
        case CRB_MECH_LANGEVIN: {
            const double mass = c.mp[0], damping = c.mp[1];
            for (int i = 0; i < D; ++i) {
                dx[i] = c.vel[i];
                double dv1 = 1.0 / mass * force[i];
                double dv2 = -damping * c.vel[i];
                dv[i] = dv1 + dv2;
            }
            break;
        }
    }
}

// Mechanics::get_random_contribution with the noise contract of SURVEY.md §8a A7:
// wiener_process (mechanics.rs:211-226) = Normal(0, sqrt(dt)) sample / dt, where the sample is
// mean + std_dev * z (rand_distr 0.5) and z comes from the uploaded table instead of ChaCha8.
void random_contribution(const Oracle& o, const Cell& c, double dt, double* rx, double* rv) {
    const int D = o.D;
    for (int i = 0; i < D; ++i) rx[i] = rv[i] = 0.0;
    if (o.cfg.noise_mode != CRB_NOISE_TABLE) return;
    if (o.cfg.mechanics_kind == CRB_MECH_NEWTON_DAMPED) return;
    const double* z = &o.noise[(o.noise_cursor * o.noise_ids + c.id) * D];
    const double std_dev = std::sqrt(dt);
    double w[MAXD];
    for (int i = 0; i < D; ++i) w[i] = (0.0 + std_dev * z[i]) / dt;
    if (o.cfg.mechanics_kind == CRB_MECH_BROWNIAN) {
        const double f = std::sqrt(2.0 * c.mp[0]);
        for (int i = 0; i < D; ++i) rx[i] = f * w[i];
    } else if (o.cfg.mechanics_kind == CRB_MECH_LANGEVIN) {
        const double mass = c.mp[0], damping = c.mp[1], kbt = c.mp[2];
        const double f = std::sqrt(2.0 * damping * kbt / mass);
        for (int i = 0; i < D; ++i) rv[i] = f * w[i];
    }
}

// local_mechanics_update -> mechanics_adams_bashforth_3 / _2 / mechanics_euler
// (chili/update_mechanics.rs:530-559, chili/solvers.rs:23-55, 156-300)
void local_mechanics_update(const Oracle& o, Cell& c, double dt) {
    const int D = o.D;
    const uint32_t N = o.cfg.solver_order;
    double force[MAXD];
    std::memcpy(force, c.force, sizeof(force));  // get_current_force_and_reset
    for (int i = 0; i < MAXD; ++i) c.force[i] = 0.0;
    c.abs_sum = 0.0;
    const bool brownian = o.cfg.mechanics_kind == CRB_MECH_BROWNIAN;
    double velocity[MAXD], position[MAXD];
    for (int i = 0; i < D; ++i) {
        velocity[i] = brownian ? 0.0 : c.vel[i];  // Brownian::velocity() is zero (mechanics.rs:371-378)
        position[i] = c.pos[i];
    }
    double dx[MAXD] = {0, 0, 0}, dv[MAXD] = {0, 0, 0}, rx[MAXD], rv[MAXD];
    calculate_increment(o, c, force, dx, dv);
    random_contribution(o, c, dt, rx, rv);
    c.ring_dx.push(dx, N);  // set_last_position: pushed BEFORE the history is read
    c.ring_dv.push(dv, N);
    const uint32_t n_prev = c.ring_dx.len;
    double new_pos[MAXD], new_vel[MAXD];
    for (int i = 0; i < D; ++i) {
        if (N == 2 && n_prev == 2) {
            new_pos[i] = helper_ab3(position[i], dx[i], c.ring_dx.v[0][i], c.ring_dx.v[1][i], dt, rx[i]);
            new_vel[i] = helper_ab3(velocity[i], dv[i], c.ring_dv.v[0][i], c.ring_dv.v[1][i], dt, rv[i]);
        } else if ((N == 2 || N == 1) && n_prev == 1) {
            new_pos[i] = helper_ab2(position[i], dx[i], c.ring_dx.v[0][i], dt, rx[i]);
            new_vel[i] = helper_ab2(velocity[i], dv[i], c.ring_dv.v[0][i], dt, rv[i]);
        } else {
            new_pos[i] = helper_euler(position[i], dx[i], dt, rx[i]);
            new_vel[i] = helper_euler(velocity[i], dv[i], dt, rv[i]);
        }
    }
    for (int i = 0; i < D; ++i) {
        c.pos[i] = new_pos[i];
        if (!brownian) c.vel[i] = new_vel[i];  // Brownian::set_velocity is a no-op
    }
}

// CartesianSubDomain::apply_boundary (cartesian_cuboid_n.rs:594-631) via
// SubDomainBox::apply_boundary (chili/update_mechanics.rs:387-406)
bool apply_boundary(const Oracle& o, Cell& c) {
    const int D = o.D;
    const bool brownian = o.cfg.mechanics_kind == CRB_MECH_BROWNIAN;
    double position[MAXD], velocity[MAXD];
    for (int i = 0; i < D; ++i) {
        position[i] = c.pos[i];
        velocity[i] = brownian ? 0.0 : c.vel[i];
    }
    for (int i = 0; i < D; ++i) {
        if (position[i] < o.dom.min[i]) {
            position[i] = 2.0 * o.dom.min[i] - position[i];
            velocity[i] = std::fabs(velocity[i]);
        }
        if (position[i] > o.dom.max[i]) {
            position[i] = 2.0 * o.dom.max[i] - position[i];
            velocity[i] = -std::fabs(velocity[i]);
        }
    }
    for (int i = 0; i < D; ++i)
        if (position[i] < o.dom.min[i] || position[i] > o.dom.max[i]) return false;
    for (int i = 0; i < D; ++i) {
        c.pos[i] = position[i];
        if (!brownian) c.vel[i] = velocity[i];
    }
    return true;
}

// key used by the partition predicate of sort_cells_in_voxels_step_1
// (chili/update_mechanics.rs:427-456): lookup failures are collected and silently dropped, i.e.
// the cell stays where it is.
inline uint64_t voxel_key_or_stay(const Oracle& o, const Cell& c, uint64_t current) {
    uint64_t idx[MAXD] = {0, 0, 0};
    if (!get_index_of(o.dom, c.pos, idx)) return current;
    if (!index_in_domain(o.dom, idx)) return current;
    return o.dom.plain(idx);
}

// sort_cells_in_voxels_step_1 (chili/update_mechanics.rs:418-477)
void sort_step_1(Oracle& o, uint64_t s, std::vector<Migration>& mig) {
    const uint64_t v_lo = o.smap.first_voxel(s), v_hi = o.smap.first_voxel(s + 1);
    std::vector<std::pair<uint64_t, Cell>> find_new_home;
    for (uint64_t v = v_lo; v < v_hi; ++v) {
        std::vector<Cell>& cells = o.vox[v];
        if (cells.empty()) continue;
        std::vector<Cell> stay;
        stay.reserve(cells.size());
        for (Cell& c : cells) {
            if (voxel_key_or_stay(o, c, v) != v)
                find_new_home.emplace_back(v, c);
            else
                stay.push_back(c);
        }
        cells.swap(stay);
    }
    for (auto& vc : find_new_home) {
        const uint64_t nv = voxel_key_or_stay(o, vc.second, vc.first);
        mig.push_back({vc.second.id, vc.first, nv});
        const uint64_t sw = o.smap.sub_of(nv);
        if (sw == s)
            o.vox[nv].push_back(vc.second);
        else
            o.q_cell[s][sw].push_back({vc.first, vc.second});
    }
}

// sort_cells_in_voxels_step_2 (chili/update_mechanics.rs:485-513)
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

// SubDomainBox::calculate_custom_domain_force (chili/update_mechanics.rs:248-274), appended to step 1
// by the DomainForce aspect (run_sim.rs:276-278): every cell of the subdomain, voxels ascending, gets
// SubDomainForce::calculate_custom_force(pos, vel, info) added to its force.  The force law is user code
// in the reference; CRB_DOMAIN_FORCE_SPRING restates cellular_raza/tests/domain_force_spring.rs:56-68
// (force_x = -spring_strength * (pos.x - 0.0), one strength and anchor per axis here).
void domain_force_step(Oracle& o, uint64_t s) {
    if (o.cfg.domain_force_kind != CRB_DOMAIN_FORCE_SPRING) return;
    const uint64_t v_lo = o.smap.first_voxel(s), v_hi = o.smap.first_voxel(s + 1);
    for (uint64_t v = v_lo; v < v_hi; ++v)
        for (Cell& c : o.vox[v])
            for (int i = 0; i < o.D; ++i) {
                const double f = -o.cfg.domain_force_strength[i] * (c.pos[i] - o.cfg.domain_force_anchor[i]);
                c.force[i] += f;
                c.abs_sum += std::fabs(f);
            }
}

bool run_forces(Oracle& o) {
    const int64_t S = (int64_t)o.smap.n_sub;
    std::atomic<int> ok{1};
    const bool pairs = o.cfg.interaction_kind != CRB_INT_NONE;  // aspects without Interaction
    parallel_for(S, o.n_threads, [&](int64_t s) {
        if (pairs && !step_1(o, (uint64_t)s)) ok = 0;
        domain_force_step(o, (uint64_t)s);
    });
    if (!ok) return false;
    if (!pairs) return true;
    parallel_for(S, o.n_threads, [&](int64_t s) {
        if (!step_2(o, (uint64_t)s)) ok = 0;
    });
    if (!ok) return false;
    parallel_for(S, o.n_threads, [&](int64_t s) { step_3(o, (uint64_t)s); });
    return true;
}

}  // namespace

// =================================================================================================
// C interface (ctypes)
// =================================================================================================
extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

void* oracle_create(const crb_config* cfg) {
    if (cfg->dim < 1 || cfg->dim > 3) {
        g_err = "dim must be 1..3";
        return nullptr;
    }
    Oracle* o = new Oracle();
    o->cfg = *cfg;
    o->D = (int)cfg->dim;
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
    o->q_pos.assign(nsub, std::vector<std::vector<PosInfo>>(nsub));
    o->q_force.assign(nsub, std::vector<std::vector<ForceInfo>>(nsub));
    o->q_cell.assign(nsub, std::vector<std::vector<SendCell>>(nsub));
    o->n_threads = 1;
    return o;
}

void oracle_destroy(void* h) { delete (Oracle*)h; }

void oracle_set_threads(void* h, int n) { ((Oracle*)h)->n_threads = n < 1 ? 1 : n; }

// Domain::decompose + insert_initial_cells (concepts-derive/src/domain.rs:292-420,
// chili/datastructures.rs:278-299): cells keep input order inside their voxel.
int oracle_set_agents(void* h, uint64_t n, const double* pos, const double* vel,
                      const double* mech_params, const double* inter_params,
                      const uint8_t* species, const uint64_t* ids) {
    Oracle& o = *(Oracle*)h;
    const int D = o.D;
    const uint32_t nm = n_mech_params(o.cfg.mechanics_kind);
    const uint32_t ni = n_inter_params(o.cfg.interaction_kind);
    for (auto& v : o.vox) v.clear();
    for (uint64_t k = 0; k < n; ++k) {
        Cell c;
        for (int i = 0; i < D; ++i) {
            c.pos[i] = pos[k * D + i];
            c.vel[i] = vel ? vel[k * D + i] : 0.0;
        }
        for (uint32_t j = 0; j < nm; ++j) c.mp[j] = mech_params[k * nm + j];
        for (uint32_t j = 0; j < ni; ++j) c.ip[j] = inter_params[k * ni + j];
        c.species = species ? species[k] : 0;
        c.id = ids ? ids[k] : k;
        uint64_t idx[MAXD] = {0, 0, 0};
        if (!get_voxel_index_of_raw(o.dom, c.pos, idx)) {
            g_err = "BoundaryError: agent " + std::to_string(k) + " at or below domain minimum";
            return CRB_ERR_BOUNDARY;
        }
        if (!index_in_domain(o.dom, idx)) {
            g_err = "IndexError: could not find voxel index for agent " + std::to_string(k);
            return CRB_ERR_INDEX;
        }
        o.vox[o.dom.plain(idx)].push_back(c);
    }
    return 0;
}

int oracle_set_noise(void* h, const double* z, uint64_t n_steps, uint64_t n_ids) {
    Oracle& o = *(Oracle*)h;
    o.noise.assign(z, z + n_steps * n_ids * o.D);
    o.noise_steps = n_steps;
    o.noise_ids = n_ids;
    o.noise_cursor = 0;
    return 0;
}

uint64_t oracle_n_agents(void* h) {
    Oracle& o = *(Oracle*)h;
    uint64_t n = 0;
    for (auto& v : o.vox) n += v.size();
    return n;
}

// one iteration of the generated loop (run_sim.rs:380-403) for aspects [Mechanics, Interaction]
int oracle_step(void* h, uint64_t n_steps) {
    Oracle& o = *(Oracle*)h;
    const int64_t S = (int64_t)o.smap.n_sub;
    const double dt = o.cfg.dt;
    for (uint64_t it = 0; it < n_steps; ++it) {
        if (o.cfg.noise_mode == CRB_NOISE_TABLE && o.cfg.mechanics_kind != CRB_MECH_NEWTON_DAMPED &&
            o.noise_cursor >= o.noise_steps) {
            g_err = "noise table exhausted";
            return CRB_ERR_STATE;
        }
        if (!run_forces(o)) {
            g_err = "CalcError: identical position for two objects";
            return CRB_ERR_CALC;
        }
        std::atomic<int> ok{1};
        // run_local_cell_funcs (chili/datastructures.rs:305-326) then apply_boundary
        parallel_for(S, o.n_threads, [&](int64_t s) {
            const uint64_t v_lo = o.smap.first_voxel(s), v_hi = o.smap.first_voxel(s + 1);
            for (uint64_t v = v_lo; v < v_hi; ++v)
                for (Cell& c : o.vox[v]) local_mechanics_update(o, c, dt);
            for (uint64_t v = v_lo; v < v_hi; ++v)
                for (Cell& c : o.vox[v])
                    if (!apply_boundary(o, c)) ok = 0;
        });
        if (!ok) {
            g_err = "BoundaryError: Particle is out of domain";
            return CRB_ERR_BOUNDARY;
        }
        o.migrations.clear();
        std::vector<std::vector<Migration>> mig(S);
        parallel_for(S, o.n_threads, [&](int64_t s) { sort_step_1(o, (uint64_t)s, mig[s]); });
        parallel_for(S, o.n_threads, [&](int64_t s) { sort_step_2(o, (uint64_t)s); });
        for (auto& m : mig) o.migrations.insert(o.migrations.end(), m.begin(), m.end());
        ++o.iteration;
        ++o.noise_cursor;
    }
    return 0;
}

// forces after update_mechanics_interaction_step_3 for the current state, in (voxel, in-voxel)
// order; the state is left untouched (forces are reset afterwards).  abs_sum may be NULL.
int oracle_forces(void* h, double* force, double* abs_sum) {
    Oracle& o = *(Oracle*)h;
    if (!run_forces(o)) {
        g_err = "CalcError: identical position for two objects";
        return CRB_ERR_CALC;
    }
    uint64_t k = 0;
    for (auto& v : o.vox)
        for (Cell& c : v) {
            for (int i = 0; i < o.D; ++i) {
                force[k * o.D + i] = c.force[i];
                c.force[i] = 0.0;
            }
            if (abs_sum) abs_sum[k] = c.abs_sum;
            c.abs_sum = 0.0;
            ++k;
        }
    return 0;
}

int oracle_get_agents(void* h, double* pos, double* vel, uint64_t* ids, uint64_t* voxel) {
    Oracle& o = *(Oracle*)h;
    uint64_t k = 0;
    for (uint64_t v = 0; v < o.vox.size(); ++v)
        for (const Cell& c : o.vox[v]) {
            for (int i = 0; i < o.D; ++i) {
                if (pos) pos[k * o.D + i] = c.pos[i];
                if (vel) vel[k * o.D + i] = c.vel[i];
            }
            if (ids) ids[k] = c.id;
            if (voxel) voxel[k] = v;
            ++k;
        }
    return 0;
}

// ring buffers oldest-first, 2*D doubles per agent
int oracle_get_aux(void* h, double* hist_dx, double* hist_dv, uint32_t* hist_len) {
    Oracle& o = *(Oracle*)h;
    uint64_t k = 0;
    for (auto& v : o.vox)
        for (const Cell& c : v) {
            for (int r = 0; r < 2; ++r)
                for (int i = 0; i < o.D; ++i) {
                    hist_dx[(k * 2 + r) * o.D + i] = r < (int)c.ring_dx.len ? c.ring_dx.v[r][i] : 0.0;
                    hist_dv[(k * 2 + r) * o.D + i] = r < (int)c.ring_dv.len ? c.ring_dv.v[r][i] : 0.0;
                }
            hist_len[k] = c.ring_dx.len;
            ++k;
        }
    return 0;
}

uint64_t oracle_n_migrations(void* h) { return ((Oracle*)h)->migrations.size(); }
int oracle_get_migrations(void* h, uint64_t* ids, uint64_t* old_voxel, uint64_t* new_voxel) {
    Oracle& o = *(Oracle*)h;
    for (size_t k = 0; k < o.migrations.size(); ++k) {
        ids[k] = o.migrations[k].id;
        old_voxel[k] = o.migrations[k].old_voxel;
        new_voxel[k] = o.migrations[k].new_voxel;
    }
    return 0;
}

int oracle_get_voxel_counts(void* h, uint32_t* counts) {
    Oracle& o = *(Oracle*)h;
    for (size_t v = 0; v < o.vox.size(); ++v) counts[v] = (uint32_t)o.vox[v].size();
    return 0;
}

// ordered candidate pairs: (i,j), j != i, j in the clipped 3^D box around i
uint64_t oracle_count_candidate_pairs(void* h) {
    Oracle& o = *(Oracle*)h;
    std::vector<uint64_t> nb;
    uint64_t total = 0;
    for (uint64_t v = 0; v < o.vox.size(); ++v) {
        const uint64_t nv = o.vox[v].size();
        if (!nv) continue;
        neighbor_plain_indices(o.dom, v, &nb);
        uint64_t c = nv - 1;
        for (uint64_t w : nb) c += o.vox[w].size();
        total += nv * c;
    }
    return total;
}

// ---- known-answer taps ---------------------------------------------------------------------------
int oracle_decomp_res(uint64_t n_voxel, uint64_t n_regions, uint64_t* n, uint64_t* m,
                      uint64_t* avg) {
    return get_decomp_res(n_voxel, n_regions, n, m, avg) ? 0 : CRB_ERR_DECOMPOSE;
}
int oracle_domain_from_interaction_range(uint32_t dim, const double* min, const double* max,
                                         double range, uint64_t* n_voxels, double* dx) {
    Domain d;
    if (!domain_from_interaction_range((int)dim, min, max, range, &d)) return CRB_ERR_BOUNDARY;
    for (uint32_t i = 0; i < dim; ++i) {
        n_voxels[i] = d.n_voxels[i];
        dx[i] = d.dx[i];
    }
    return 0;
}
int oracle_domain_from_n_voxels(uint32_t dim, const double* min, const double* max,
                                const uint64_t* n_voxels, double* dx) {
    Domain d;
    if (!domain_from_n_voxels((int)dim, min, max, n_voxels, &d)) return CRB_ERR_BOUNDARY;
    for (uint32_t i = 0; i < dim; ++i) dx[i] = d.dx[i];
    return 0;
}
// sizes of the chunks create_subdomains produces (cartesian_cuboid_n.rs:500-583)
int oracle_subdomain_sizes(uint64_t n_voxel, uint64_t n_regions, uint64_t* sizes,
                           uint64_t* n_sizes) {
    SubdomainMap sm;
    if (!sm.init(n_voxel, n_regions)) return CRB_ERR_DECOMPOSE;
    std::vector<uint64_t> cnt;
    for (uint64_t v = 0; v < n_voxel; ++v) {
        uint64_t s = sm.sub_of(v);
        if (s >= cnt.size()) cnt.resize(s + 1, 0);
        cnt[s]++;
    }
    *n_sizes = cnt.size();
    for (size_t i = 0; i < cnt.size(); ++i) sizes[i] = cnt[i];
    return 0;
}
uint64_t oracle_sub_of(void* h, uint64_t plain) { return ((Oracle*)h)->smap.sub_of(plain); }
int oracle_voxel_index(void* h, const double* pos, uint64_t* idx, int setup_variant) {
    Oracle& o = *(Oracle*)h;
    bool ok = setup_variant ? get_voxel_index_of_raw(o.dom, pos, idx) : get_index_of(o.dom, pos, idx);
    return ok ? 0 : CRB_ERR_BOUNDARY;
}
uint64_t oracle_neighbors(void* h, uint64_t plain, uint64_t* out) {
    std::vector<uint64_t> nb;
    neighbor_plain_indices(((Oracle*)h)->dom, plain, &nb);
    for (size_t i = 0; i < nb.size(); ++i) out[i] = nb[i];
    return nb.size();
}
double oracle_helper_euler(double x, double dx, double dt, double r) { return helper_euler(x, dx, dt, r); }
double oracle_helper_ab2(double x, double dx0, double dx1, double dt, double r) {
    return helper_ab2(x, dx0, dx1, dt, r);
}
double oracle_helper_ab3(double x, double dx0, double dx1, double dx2, double dt, double r) {
    return helper_ab3(x, dx0, dx1, dx2, dt, r);
}
double oracle_xapy(double x, double a, double y) { return xapy(x, a, y); }
// pushes `n` scalars into a RingBuffer of capacity cap, returns the iteration order
uint32_t oracle_ring(const double* values, uint32_t n, uint32_t cap, double* out) {
    Ring<1> r;
    for (uint32_t i = 0; i < n; ++i) r.push(&values[i], cap);
    for (uint32_t i = 0; i < r.len; ++i) out[i] = r.v[i][0];
    return r.len;
}
// one pair evaluation: Interaction::calculate_force_between(own, ext) with parameter rows
int oracle_pair_force(uint32_t interaction_kind, uint32_t dim, const double* own_pos,
                      const double* ext_pos, const double* own_params, uint8_t own_species,
                      double ext_radius, uint8_t ext_species, double* f_own, double* f_ext) {
    Oracle o;
    o.cfg.interaction_kind = interaction_kind;
    o.D = (int)dim;
    Cell own;
    for (uint32_t j = 0; j < n_inter_params(interaction_kind); ++j) own.ip[j] = own_params[j];
    own.species = own_species;
    return calculate_force_between(o, own, own_pos, ext_pos, ext_radius, ext_species, f_own, f_ext)
               ? 0
               : CRB_ERR_CALC;
}

}  // extern "C"
