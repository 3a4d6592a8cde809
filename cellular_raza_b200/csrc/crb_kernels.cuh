// crb_kernels.cuh — device code of the mechanics/interaction step (sm_100a).
//
// Stages of one step:
//   k_force_integrate  pair forces over the 3^D voxel neighbourhood (A5/A6), increment +
//                      Adams-Bashforth/Euler update (A7/A8), reflective boundary (A9), new voxel
//                      key + per-voxel histogram (A1/A10)             -> FP64-pipe bound
//   k_scan_*           exclusive prefix scan of the histogram          -> HBM bound (tiny)
//   k_fill, k_reorder  stable counting-sort scatter of the full state  -> HBM bound
// Row labels (A1..A10) refer to SURVEY.md §8a; reference file:line are cited per function.
#pragma once
#include "crb_internal.cuh"

namespace crb {

constexpr int kForceThreads = 128;

__device__ __forceinline__ void latch_error(int* err, int code, u64 id) {
    if (atomicCAS(&err[1], 0, 1) == 0) {
        err[0] = code;
        *reinterpret_cast<u64*>(err + 2) = id;
        __threadfence();
    }
}

// A1: CartesianSubDomain::get_index_of (cartesian_cuboid_n.rs:469-490) for one axis.
// IEEE subtraction + IEEE division + truncation (num-traits to_usize: Some iff -1 < q < 2^64).
// Never a multiply by a reciprocal: the result must be bit-exact.
__device__ __forceinline__ bool voxel_axis_index(double p, double mn, double dx, u64* idx) {
    const double q = __ddiv_rn(__dsub_rn(p, mn), dx);
    if (q > -1.0 && q < 18446744073709551616.0) {
        *idx = q <= 0.0 ? 0ull : (u64)q;
        return true;
    }
    return false;
}

// global plain index of a position, or false when chili's lookup would fail (to_usize None or
// index not a key of voxel_index_to_plain_index): the agent then stays where it is
// (chili/update_mechanics.rs:430-455 collects these errors and drops them).
template <int D>
__device__ __forceinline__ bool voxel_of(const Grid& g, const double* p, int* i0, int* i1,
                                         int* i2) {
    u64 a0, a1 = 0, a2;
    if (!voxel_axis_index(p[0], g.mn[0], g.dx[0], &a0)) return false;
    if (D == 3) {
        if (!voxel_axis_index(p[1], g.mn[1], g.dx[1], &a1)) return false;
    }
    if constexpr (D == 1) {
        a2 = 0;  // 1-D: the only axis is the slab axis a0
    } else {
        if (!voxel_axis_index(p[D - 1], g.mn[D - 1], g.dx[D - 1], &a2)) return false;
    }
    if (a0 >= (u64)g.ng[0] || a1 >= (u64)g.ng[1] || a2 >= (u64)g.ng[2]) return false;
    *i0 = (int)a0;
    *i1 = (int)a1;
    *i2 = (int)a2;
    return true;
}

// ------------------------------------------------------------------------------------------------
// A6: pair potentials.  Every potential of the closed model set returns
//   f_own = z * coef,  f_ext = -f_own,   z = own_pos - ext_pos,
// so one scalar `coef` per (own parameters, ext information, r^2) describes the pair.
// Tolerance contract: 1e-12 relative on the accumulated force (BASELINE.json north_star), which is
// why reciprocal square roots and multiplication chains replace the reference's
// norm/normalize/powf sequence (SURVEY.md "hard parts").
// ------------------------------------------------------------------------------------------------
// 1/sqrt(x) for normal, positive x: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~20 good bits) refined by
// two Newton steps y <- y + (y/2)(1 - x y^2)  => relative error < 2^-52 (checked on the device by
// tests/test_gpu_kernels.py::test_rsqrt_accuracy).  Replaces CUDA's rsqrt(), whose slow path for
// denormals/inf costs a CALL per pair.  x == 0 gives inf/NaN: callers test r2 == 0 themselves.
__device__ __forceinline__ double rsqrt_nr(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double e = fma(-hx * y, y, 0.5);  // (1 - x y^2) / 2
        y = fma(y, e, y);
    }
    return y;
}

// exp(x) for |x| < 700 without the special-case handling of CUDA's exp() (overflow / underflow / NaN paths):
// k = round(x / ln 2) by the 1.5 * 2^52 shift, r = x - k ln 2 in two steps (|r| <= 0.3466), Taylor polynomial
// of degree 13 (remainder r^14 / 14! < 5e-18), exponent add.  Relative error < 2.3e-16 (checked on the device
// by tests/test_gpu_kernels.py::test_exp_bounded_accuracy).  Used by the Morse potential, whose argument is
// clamped from below in the kernels and bounded from above by a host-side check (a * (R1 + R2) < 700).
__device__ __forceinline__ double exp_bounded(double x) {
    const double shift = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(x, 1.4426950408889634074, shift);
    const int k = __double2loint(t);
    const double kd = t - shift;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;          // 1/13!
    p = fma(p, r, 2.0876756987868100e-09);      // 1/12!
    p = fma(p, r, 2.5052108385441720e-08);      // 1/11!
    p = fma(p, r, 2.7557319223985888e-07);      // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);      // 1/9!
    p = fma(p, r, 2.4801587301587302e-05);      // 1/8!
    p = fma(p, r, 1.9841269841269841e-04);      // 1/7!
    p = fma(p, r, 1.3888888888888889e-03);      // 1/6!
    p = fma(p, r, 8.3333333333333332e-03);      // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);      // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);      // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

template <int INTER>
struct Pot;

// cellular_raza-examples/cell_sorting/src/main.rs:43-97; benchmark copy with the opposite sign
// cellular_raza-benchmarks/src/cell_sorting.rs:34-88.
template <bool BENCH>
struct PotCellSorting {
    static constexpr int NI = 3;
    static constexpr bool kUsesSpecies = true;
    double R, strength, range, rmin, inv_rmin;
    __device__ __forceinline__ void init(const double* ip) {
        R = ip[0];
        strength = ip[1];
        range = ip[2];
        rmin = R * 0.3162277660168379;
        inv_rmin = 1.0 / rmin;
    }
    // UPPER = false drops the upper clamp, which can only act for a negative radius sum (the tile
    // kernel is not planned for such parameters, see plan_tile)
    template <bool UPPER>
    __device__ __forceinline__ double coef_t(double r2, double ext_radius, bool same_species) const {
        // branch-free: r2 == 0 (own_pos == ext_pos -> zero force, :60-63) is selected at the end
        const double rinv = rsqrt_nr(r2);
        const double r_true = r2 * rinv;
        const bool clamp = r_true < rmin;  // (:51-52, 67-68): r is clamped, dir is not
        const double r = clamp ? rmin : r_true;
        const double rcinv = clamp ? inv_rmin : rinv;
        const double Rs = R + ext_radius;
        const double sinv = Rs * rcinv;  // 1 / sigma
        const double nbound = -4.0 - sinv;
        const double s2 = sinv * sinv;
        double val = s2 - s2 * s2;  // (1/sigma)^2 - (1/sigma)^4
        // clamp(val, -bound, bound): x - x^2 <= 1/4 < 4 <= bound whenever 1/sigma >= 0, so the upper
        // clamp can only act for a negative radius sum.  Compare-selects instead of fmin/fmax (no NaN
        // reaches a selected result: r2 == 0 is masked below)
        if (UPPER) {
            if (Rs < 0.0) val = val > -nbound ? -nbound : val;
        }
        val = val < nbound ? nbound : val;
        const double st = strength * val;
        // attracting part only between equal species (:88-96); spatial_cutoff =
        // (1 + signum(range*(R1+R2) - r)) / 2 with signum(+0) = 1 (:74-75)
        const bool on = (same_species | (st < 0.0)) & (range * Rs - r >= 0.0) & (r2 != 0.0);
        const double c = on ? st * rinv : 0.0;
        return BENCH ? c : -c;
    }
    __device__ __forceinline__ double coef(double r2, double ext_radius, bool same_species,
                                           int* flag) const {
        return coef_t<true>(r2, ext_radius, same_species);
    }
};
template <>
struct Pot<CRB_INT_CELL_SORTING> : PotCellSorting<false> {};
template <>
struct Pot<CRB_INT_CELL_SORTING_BENCH> : PotCellSorting<true> {};

// calculate_morse_interaction (interaction.rs:151-182)
template <>
struct Pot<CRB_INT_MORSE> {
    static constexpr int NI = 4;
    static constexpr bool kUsesSpecies = false;
    double radius, a, cutoff, m2sa;
    __device__ __forceinline__ void init(const double* ip) {
        radius = ip[0];
        a = ip[1];
        cutoff = ip[2];
        m2sa = -2.0 * ip[3] * ip[1];
    }
    __device__ __forceinline__ double coef(double r2, double ext_radius, bool, int*) const {
        const double rinv = rsqrt_nr(r2);
        const double dist = r2 * rinv;
        const double e = exp(-a * (dist - (radius + ext_radius)));
        const double c = m2sa * e * (1.0 - e) * rinv;
        return (r2 == 0.0 || dist > cutoff) ? 0.0 : c;  // (:158-162) zero beyond the cutoff / at r = 0
    }
    // tile kernels (uniform parameters): branch-free lean exponential.  The argument is at most
    // a * (R1 + R2) < 700 (plan_tile checks it, otherwise the per-agent kernel with coef() runs); below -700
    // the true value is < 1e-304 and the clamp changes the coefficient by less than that.
    __device__ __forceinline__ double coef_bounded(double r2, double ext_radius) const {
        const double rinv = rsqrt_nr(r2);
        const double dist = r2 * rinv;
        double x = -a * (dist - (radius + ext_radius));
        x = x < -700.0 ? -700.0 : x;
        const double e = exp_bounded(x);
        const double c = m2sa * e * (1.0 - e) * rinv;
        return (r2 == 0.0 || dist > cutoff) ? 0.0 : c;
    }
};

// BoundLennardJones (interaction.rs:122-138); no r == 0 guard in the reference -> NaN
template <>
struct Pot<CRB_INT_BOUND_LJ> {
    static constexpr int NI = 4;
    static constexpr bool kUsesSpecies = false;
    double eps4, sigma, bound, cutoff;
    __device__ __forceinline__ void init(const double* ip) {
        eps4 = 4.0 * ip[0];
        sigma = ip[1];
        bound = ip[2];
        cutoff = ip[3];
    }
    __device__ __forceinline__ double coef(double r2, double, bool, int*) const {
        if (r2 == 0.0) return __longlong_as_double(0x7ff8000000000000ll);
        const double rinv = rsqrt_nr(r2);
        const double r = r2 * rinv;
        const double x = sigma * rinv;
        const double x2 = x * x;
        const double x5 = x2 * x2 * x;
        const double x11 = x5 * x5 * x;
        const double val = eps4 * rinv * (12.0 * x11 - 6.0 * x5);
        const double mx = bound * rinv;
        const double q = cutoff >= r ? 1.0 : 0.0;
        return -(q * fmin(mx, val)) * rinv;
    }
};

// MiePotential (interaction.rs:345-389)
template <>
struct Pot<CRB_INT_MIE> {
    static constexpr int NI = 6;
    static constexpr bool kUsesSpecies = false;
    double radius, strength, bound, cutoff, en, em, sig_factor, mie_c;
    __device__ __forceinline__ void init(const double* ip) {
        radius = ip[0];
        strength = ip[1];
        bound = ip[2];
        cutoff = ip[3];
        en = ip[4];
        em = ip[5];
        sig_factor = pow(em / en, 1.0 / (en - em));              // radius_to_sigma_factor()
        mie_c = en / (en - em) * pow(en / em, em / (en - em));  // mie_constant
    }
    __device__ __forceinline__ double coef(double r2, double ext_radius, bool, int* flag) const {
        if (r2 == 0.0) {
            *flag = 1;  // CalcError: identical position for two objects
            return 0.0;
        }
        const double rinv = rsqrt_nr(r2);
        const double r = r2 * rinv;
        if (r > cutoff) return 0.0;
        const double sigma = sig_factor * (radius + ext_radius);
        const double x = sigma * rinv;
        const double pp = en / sigma * pow(x, en + 1.0) - em / sigma * pow(x, em + 1.0);
        const double force = fmin(mie_c * strength * pp, bound);
        return force * rinv;
    }
};

// aspects without Interaction: no pair forces (the neighbour loops are skipped at compile time)
template <>
struct Pot<CRB_INT_NONE> {
    static constexpr int NI = 0;
    static constexpr bool kUsesSpecies = false;
    __device__ __forceinline__ void init(const double*) {}
    __device__ __forceinline__ double coef(double, double, bool, int*) const { return 0.0; }
};

// SubDomainBox::calculate_custom_domain_force (chili/update_mechanics.rs:248-274): the force of the
// subdomain on one agent is ADDED to the pair forces accumulated so far (AuxStorage::add_force).
// CRB_DOMAIN_FORCE_SPRING = cellular_raza/tests/domain_force_spring.rs:56-68, same operation order.
template <int D>
__device__ __forceinline__ void add_domain_force(const KParams& P, const double* x, double* f) {
    if (P.dforce == CRB_DOMAIN_FORCE_SPRING) {
#pragma unroll
        for (int c = 0; c < D; ++c) f[c] = __dadd_rn(f[c], __dmul_rn(-P.dk[c], __dsub_rn(x[c], P.dx0[c])));
    }
}

// coefficient as the tile kernel evaluates it (same arithmetic; cell sorting without the upper clamp)
template <int INTER>
__device__ __forceinline__ double coef_tile(const Pot<INTER>& p, double r2, double ext_radius,
                                            bool same_species, int* flag) {
    if constexpr (INTER == CRB_INT_CELL_SORTING || INTER == CRB_INT_CELL_SORTING_BENCH)
        return p.template coef_t<false>(r2, ext_radius, same_species);
    else if constexpr (INTER == CRB_INT_MORSE)
        return p.coef_bounded(r2, ext_radius);
    else
        return p.coef(r2, ext_radius, same_species, flag);
}

// coefficient of a pair under UNIFORM parameters outside the tile kernels (k_force_integrate<UNIFORM>, the dense
// voxel fall-back of the tile kernels): Morse takes the same bounded exponential as the tile kernels, so that every
// uniform-parameter path gives the same bits (tile == per-agent kernel, slabs == single handle)
template <int INTER>
__device__ __forceinline__ double coef_uniform(const Pot<INTER>& p, double r2, double ext_radius,
                                               bool same_species, int* flag) {
    if constexpr (INTER == CRB_INT_MORSE)
        return p.coef_bounded(r2, ext_radius);
    else
        return p.coef(r2, ext_radius, same_species, flag);
}

// ------------------------------------------------------------------------------------------------
// A8 helper formulas with the reference's operation order, no fma
// (chili/solvers.rs:303-347; Xapy = self * a + y, concepts/src/reactions.rs:72-82)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double xapy(double x, double a, double y) {
    return __dadd_rn(__dmul_rn(x, a), y);
}
__device__ __forceinline__ double helper_euler(double x, double dx, double dt, double dr) {
    return xapy(xapy(dx, dt, x), 1.0, __dmul_rn(dr, dt));
}
__device__ __forceinline__ double helper_ab2(double x, double dx0, double dx1, double dt,
                                             double dr) {
    return xapy(xapy(xapy(dx0, 1.5, __dmul_rn(dx1, -0.5)), dt, x), 1.0, __dmul_rn(dr, dt));
}
__device__ __forceinline__ double helper_ab3(double x, double dx0, double dx1, double dx2,
                                             double dt, double dr) {
    const double f0 = 23.0 / 12.0, f1 = -16.0 / 12.0, f2 = 5.0 / 12.0;
    return xapy(xapy(xapy(dx0, f0, xapy(dx1, f1, __dmul_rn(dx2, f2))), dt, x), 1.0,
                __dmul_rn(dr, dt));
}

// counter-based standard normal numbers for CRB_NOISE_PHILOX: Philox4x32-10 keyed by the seed,
// counter = (agent id, iteration); Box-Muller on the four words.
__device__ __forceinline__ void philox_round(u32& c0, u32& c1, u32& c2, u32& c3, u32 k0, u32 k1) {
    const u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const u32 n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
}
__device__ __forceinline__ void philox_normal4(u64 seed, u64 id, u64 iter, u32 stream,
                                               double* z) {
    u32 c0 = (u32)id, c1 = (u32)(id >> 32), c2 = (u32)iter, c3 = (u32)(iter >> 32) ^ (stream << 24);
    u32 k0 = (u32)seed, k1 = (u32)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    // two Box-Muller pairs from 4 x 32 random bits (32-bit uniforms in (0,1])
    const double u0 = ((double)c0 + 1.0) * (1.0 / 4294967296.0);
    const double u1 = ((double)c1 + 0.5) * (1.0 / 4294967296.0);
    const double u2 = ((double)c2 + 1.0) * (1.0 / 4294967296.0);
    const double u3 = ((double)c3 + 0.5) * (1.0 / 4294967296.0);
    const double ra = sqrt(-2.0 * log(u0)), rb = sqrt(-2.0 * log(u2));
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    z[0] = ra * c;
    z[1] = ra * s;
    sincospi(2.0 * u3, &s, &c);
    z[2] = rb * c;
    z[3] = rb * s;
}

// ------------------------------------------------------------------------------------------------
// A7-A10 for one agent: increment, Adams-Bashforth/Euler update, reflective boundary, new voxel key,
// migration flag and histogram.  Shared by every force kernel (fused epilogue: the force never
// travels through HBM).
// ------------------------------------------------------------------------------------------------
// per-agent inputs of the integrator; loaded early (before the pair loop) so that their global-memory
// latency overlaps the force computation
template <int D>
struct AgentDyn {
    double vel[D], oldx[D], oldv[D];
    int hlen;
};
template <int D>
__device__ __forceinline__ AgentDyn<D> load_agent_dyn(const KParams& P, const u64 i) {
    AgentDyn<D> a;
    const bool brownian = P.mech == CRB_MECH_BROWNIAN;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        a.vel[c] = brownian ? 0.0 : P.A.vel[c][i];  // Brownian::velocity() == 0 (mechanics.rs:371-378)
        a.oldx[c] = P.A.h1x[c][i];
        a.oldv[c] = P.A.h1v[c][i];
    }
    a.hlen = P.A.hlen[i];
    return a;
}

template <int D, bool UNIFORM>
__device__ __forceinline__ void integrate_agent(const KParams& P, const u64 i, const u32 v,
                                                const u32 own_lo, const int i0, const int i1,
                                                const int i2, const double* xi, const double* acc_pairs,
                                                const AgentDyn<D>& dyn) {
    const Grid& g = P.g;
    double acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = acc_pairs[c];
    add_domain_force<D>(P, xi, acc);  // aspect DomainForce (no-op when dforce == NONE)
    const int n1 = g.ng[1], n2 = g.ng[2];
    // ---- A7: Mechanics::calculate_increment / get_random_contribution ---------------------------
    double mp[CRB_MAX_MECH_PARAMS];
#pragma unroll
    for (int k = 0; k < 3; ++k) mp[k] = UNIFORM ? P.u.mp[k] : (P.A.mp[k] ? P.A.mp[k][i] : 0.0);
    double vel[D], dxv[D], dvv[D], rx[D], rv[D];
    const bool brownian = P.mech == CRB_MECH_BROWNIAN;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        vel[c] = dyn.vel[c];
        rx[c] = 0.0;
        rv[c] = 0.0;
    }
    if (P.mech == CRB_MECH_NEWTON_DAMPED) {  // mechanics.rs:165-172
        const double damping = mp[0], mass = mp[1];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            dxv[c] = vel[c];
            // x / 1 == x exactly: the shipped mass (cell_sorting/src/main.rs:141) skips three IEEE divisions
            const double a_over_m = (UNIFORM && mass == 1.0) ? acc[c] : __ddiv_rn(acc[c], mass);
            dvv[c] = __dsub_rn(a_over_m, __dmul_rn(damping, vel[c]));
        }
    } else if (brownian) {  // mechanics.rs:349-356
        const double f = __ddiv_rn(mp[0], mp[1]);
#pragma unroll
        for (int c = 0; c < D; ++c) {
            dxv[c] = __dmul_rn(f, acc[c]);
            dvv[c] = 0.0;
        }
    } else {  // Langevin, mechanics.rs:450-462
        const double mass = mp[0], damping = mp[1];
        const double minv = __ddiv_rn(1.0, mass);
#pragma unroll
        for (int c = 0; c < D; ++c) {
            dxv[c] = vel[c];
            dvv[c] = __dadd_rn(__dmul_rn(minv, acc[c]), __dmul_rn(-damping, vel[c]));
        }
    }
    if (P.noise_mode != CRB_NOISE_NONE && P.mech != CRB_MECH_NEWTON_DAMPED) {
        // wiener_process (mechanics.rs:211-226): (mean + std_dev * z) / dt with mean = 0
        double zn[4];
        const u64 id = P.A.id[i];
        if (P.noise_mode == CRB_NOISE_TABLE) {
            // the table is indexed by the caller's agent id: an id beyond its rows is an IndexError, not a read
            // past the allocation
            const bool ok = id < P.noise_ids;
            if (!ok) latch_error(P.err, CRB_ERR_INDEX, id);
#pragma unroll
            for (int c = 0; c < D; ++c) zn[c] = ok ? P.noise[(P.noise_cursor * P.noise_ids + id) * D + c] : 0.0;
        } else {
            philox_normal4(P.seed, id, P.iteration, 0, zn);
        }
        const double sd = sqrt(P.dt);
        double scale;
        if (brownian)
            scale = sqrt(__dmul_rn(2.0, mp[0]));
        else
            scale = sqrt(__ddiv_rn(__dmul_rn(__dmul_rn(2.0, mp[1]), mp[2]), mp[0]));
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double w = __ddiv_rn(__dadd_rn(0.0, __dmul_rn(sd, zn[c])), P.dt);
            if (brownian)
                rx[c] = __dmul_rn(scale, w);
            else
                rv[c] = __dmul_rn(scale, w);
        }
    }

    // ---- A8: local_mechanics_update (chili/solvers.rs:156-300) -----------------------------------
    const int N = P.order;
    const int hlen = dyn.hlen;
    const int n_prev = N == 0 ? 0 : (hlen + 1 < N ? hlen + 1 : N);  // ring size AFTER the push
    double p_new[D], v_new[D], oldx[D], oldv[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        oldx[c] = dyn.oldx[c];
        oldv[c] = dyn.oldv[c];
        if (N == 2 && n_prev == 2) {
            // ring = [previous, current] oldest first; dx = [current, ring[0], ring[1]]
            p_new[c] = helper_ab3(xi[c], dxv[c], oldx[c], dxv[c], P.dt, rx[c]);
            v_new[c] = helper_ab3(vel[c], dvv[c], oldv[c], dvv[c], P.dt, rv[c]);
        } else if (n_prev == 1) {
            // ring = [current]; dx = [current, current]
            p_new[c] = helper_ab2(xi[c], dxv[c], dxv[c], P.dt, rx[c]);
            v_new[c] = helper_ab2(vel[c], dvv[c], dvv[c], P.dt, rv[c]);
        } else {
            p_new[c] = helper_euler(xi[c], dxv[c], P.dt, rx[c]);
            v_new[c] = helper_euler(vel[c], dvv[c], P.dt, rv[c]);
        }
    }

    // ---- A9: CartesianSubDomain::apply_boundary (cartesian_cuboid_n.rs:594-631) ------------------
    bool outside = false;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double p = p_new[c], w = brownian ? 0.0 : v_new[c];
        if (p < g.mn[c]) {
            p = __dsub_rn(__dmul_rn(2.0, g.mn[c]), p);
            w = fabs(w);
        }
        if (p > g.mx[c]) {
            p = __dsub_rn(__dmul_rn(2.0, g.mx[c]), p);
            w = -fabs(w);
        }
        outside |= (p < g.mn[c]) | (p > g.mx[c]);
        p_new[c] = p;
        v_new[c] = w;
    }
    if (outside) latch_error(P.err, CRB_ERR_BOUNDARY, P.A.id[i]);

    // ---- A1/A10: new voxel, migration flag, histogram --------------------------------------------
    int q0, q1, q2;
    u32 nk = v;
    const u64 gp_old = ((u64)(i0 + g.p0) * n1 + i1) * n2 + i2;
    u64 gp_new = gp_old;
    if (voxel_of<D>(g, p_new, &q0, &q1, &q2)) {
        const int l0 = q0 - g.p0;
        if (l0 >= 0 && l0 < g.L0) {
            nk = ((u32)l0 * (u32)n1 + (u32)q1) * (u32)n2 + (u32)q2;
            gp_new = ((u64)q0 * n1 + q1) * n2 + q2;
        } else {
            // a voxel of the domain that is neither owned nor in a ghost plane: the agent crossed more than one
            // plane of a neighbouring slab in ONE step (chili would move it anywhere; the slab protocol only
            // exchanges with the adjacent slabs).  Never silently: IndexError.
            latch_error(P.err, CRB_ERR_INDEX, P.A.id[i]);
        }
    }
    u64 rank_old = i - own_lo;
    if (rank_old > CRB_SK_RANK_MASK) rank_old = CRB_SK_RANK_MASK;
    u64 sk = (gp_old << CRB_SK_RANK_BITS) | rank_old;
    if (nk != v) {
        sk |= CRB_SK_MOVED;
        // a mover that changes chili subdomain travels as SendCell and is appended after the local movers
        if (g.sm_nsub > 1 && g.sub_of(gp_old) != g.sub_of(gp_new)) sk |= CRB_SK_REMOTE;
    }
    P.sortkey[i] = sk;
    P.newkey[i] = nk;
    P.rank_arb[i] = atomicAdd(&P.count_new[nk], 1u);

#pragma unroll
    for (int c = 0; c < D; ++c) {
        P.B.pos[c][i] = p_new[c];
        P.B.vel[c][i] = v_new[c];
        P.B.h1x[c][i] = N >= 1 ? dxv[c] : oldx[c];
        P.B.h1v[c][i] = N >= 1 ? dvv[c] : oldv[c];
        if (P.save_prev) {
            P.B.h0x[c][i] = oldx[c];
            P.B.h0v[c][i] = oldv[c];
        }
    }
    P.B.hlen[i] = (u8)n_prev;
}

// ------------------------------------------------------------------------------------------------
// force + integrate + boundary + new key, one thread per agent of A (sorted by voxel).
//
// Pair weights (SURVEY.md finding 2; chili/update_mechanics.rs:186-234, 285-382): agent i always
// receives 1/2 f_own(own = i, ext = j).  The other half, f_ext(own = j, ext = i), reaches i when j
// is in the same voxel, in another chili subdomain, or always in PHYSICAL mode.  Because chili's
// subdomains are contiguous plain-index chunks and agents are sorted by plain index, "same
// voxel" and "same subdomain" are index-range tests on j.
// ------------------------------------------------------------------------------------------------
template <int D, int INTER, bool UNIFORM, bool FORCE_ONLY>
__global__ void __launch_bounds__(kForceThreads)
    k_force_integrate(const KParams P, const u64 first, const u64 count) {
    const u64 t = (u64)blockIdx.x * kForceThreads + threadIdx.x;
    if (t >= (P.n_dev ? *P.n_dev : count)) return;
    const u64 i = first + t;
    const Grid& g = P.g;
    const int n1 = g.ng[1], n2 = g.ng[2];

    double xi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) xi[c] = P.A.pos[c][i];
    double ipi[CRB_MAX_INTER_PARAMS];
#pragma unroll
    for (int k = 0; k < Pot<INTER>::NI; ++k) ipi[k] = UNIFORM ? P.u.ip[k] : P.A.ip[k][i];
    Pot<INTER> own;
    own.init(ipi);
    const u8 spi = Pot<INTER>::kUsesSpecies ? P.A.species[i] : 0;

    const u32 v = P.key[i];
    const int i2 = (int)(v % (u32)n2);
    const u32 tq = v / (u32)n2;
    const int i1 = (int)(tq % (u32)n1);
    const int i0 = (int)(tq / (u32)n1);  // local plane

    const u32 own_lo = P.start[v], own_hi = P.start[v + 1];
    u32 sub_lo = 0, sub_hi = 0;  // agents of i's chili subdomain (empty range => all weights 1)
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

    double acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.0;
    int calc_flag = 0;

    const int z_lo = i2 > 0 ? i2 - 1 : 0;
    const int z_hi = i2 + 1 < n2 ? i2 + 1 : n2 - 1;
    for (int d0 = -1; d0 <= 1; ++d0) {
        if (INTER == CRB_INT_NONE) break;  // aspects without Interaction
        const int X = i0 + d0;
        if (X < 0 || X >= g.L0) continue;
        for (int d1 = -1; d1 <= 1; ++d1) {
            const int Y = i1 + d1;
            if (Y < 0 || Y >= n1) continue;
            const u32 row = ((u32)X * (u32)n1 + (u32)Y) * (u32)n2;
            const u32 jb = P.start[row + z_lo], je = P.start[row + z_hi + 1];
            for (u32 j = jb; j < je; ++j) {
                double z[D];
                double r2 = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    z[c] = xi[c] - __ldg(&P.A.pos[c][j]);
                    r2 = fma(z[c], z[c], r2);
                }
                if (j == (u32)i) continue;
                const bool same_voxel = (j >= own_lo) & (j < own_hi);
                const bool same_sub = (j >= sub_lo) & (j < sub_hi);
                const bool keep = same_voxel | !same_sub;  // second half reaches i
                const u8 spj = Pot<INTER>::kUsesSpecies ? __ldg(&P.A.species[j]) : 0;
                double cf;
                if (UNIFORM) {
                    const double ext_radius = ipi[0];
                    cf = coef_uniform<INTER>(own, r2, ext_radius, spi == spj, &calc_flag);
                    cf = keep ? cf : 0.5 * cf;
                } else {
                    double ipj[CRB_MAX_INTER_PARAMS];
#pragma unroll
                    for (int k = 0; k < Pot<INTER>::NI; ++k) ipj[k] = __ldg(&P.A.ip[k][j]);
                    const double ci = own.coef(r2, ipj[0], spi == spj, &calc_flag);
                    cf = 0.5 * ci;
                    if (keep) {
                        Pot<INTER> other;
                        other.init(ipj);
                        cf += 0.5 * other.coef(r2, ipi[0], spi == spj, &calc_flag);
                    }
                }
#pragma unroll
                for (int c = 0; c < D; ++c) acc[c] = fma(z[c], cf, acc[c]);
            }
        }
    }
    if (calc_flag) latch_error(P.err, CRB_ERR_CALC, P.A.id[i]);

    if constexpr (FORCE_ONLY) {
        add_domain_force<D>(P, xi, acc);
#pragma unroll
        for (int c = 0; c < D; ++c) P.force_out[(u64)c * P.cap + i] = acc[c];
    } else {
        integrate_agent<D, UNIFORM>(P, i, v, own_lo, i0, i1, i2, xi, acc, load_agent_dyn<D>(P, i));
    }
}

// the kernels below are not templates: only the translation unit of the C ABI (crb_api.cu) defines them; the tile
// kernel translation units (crb_tile*.cu) set CRB_TILE_TU
#ifndef CRB_TILE_TU
// ------------------------------------------------------------------------------------------------
// exclusive scan of count_new[0..n) -> start_new[0..n], start_new[n] = total.
// three passes: per-block scan (4096 items), scan of block sums (one block), add.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ u32 block_exclusive_scan(u32 x, u32* total) {
    __shared__ u32 warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 w = warp_sums[lane];
        u32 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        warp_sums[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    const u32 r = incl - x + warp_sums[wid];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads)
    k_scan_tiles(const u32* __restrict__ in, u32* __restrict__ out, u32* __restrict__ tile_sums,
                 u64 n) {
    __shared__ u32 tot;
    const u64 base = (u64)blockIdx.x * kScanTile + (u64)threadIdx.x * kScanItems;
    u32 v[kScanItems];
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? in[base + k] : 0u;
        s += v[k];
    }
    u32 off = block_exclusive_scan(s, &tot);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = off;
        off += v[k];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// the whole scan in one launch when the histogram fits one tile (n <= kScanTile voxels: every example the
// reference ships): out[k] = base + sum(in[0..k)), out[n] = base + total, like the three passes together
__global__ void __launch_bounds__(kScanThreads)
    k_scan_single(const u32* __restrict__ in, u32* __restrict__ out, u64 n, u32 base) {
    __shared__ u32 tot;
    const u64 b = (u64)threadIdx.x * kScanItems;
    u32 v[kScanItems];
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = b + k < n ? in[b + k] : 0u;
        s += v[k];
    }
    u32 off = block_exclusive_scan(s, &tot) + base;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (b + k < n) out[b + k] = off;
        off += v[k];
    }
    if (threadIdx.x == 0) out[n] = tot + base;
}

__global__ void __launch_bounds__(kScanThreads)
    k_scan_sums(u32* __restrict__ tile_sums, u32 n_tiles, u32* __restrict__ grand_total, u32 base) {
    __shared__ u32 tot;
    u32 carry = base;  // index of the first owned agent: start[] holds absolute agent indices
    for (u32 base = 0; base < n_tiles; base += kScanThreads) {
        const u32 idx = base + threadIdx.x;
        const u32 x = idx < n_tiles ? tile_sums[idx] : 0u;
        const u32 off = block_exclusive_scan(x, &tot);
        if (idx < n_tiles) tile_sums[idx] = off + carry;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(kScanThreads)
    k_scan_add(u32* __restrict__ out, const u32* __restrict__ tile_sums, u64 n) {
    const u64 base = (u64)blockIdx.x * kScanTile + (u64)threadIdx.x * kScanItems;
    const u32 add = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) out[base + k] += add;
}

// ------------------------------------------------------------------------------------------------
// A10: stable counting-sort scatter.  k_fill places every agent at an arbitrary slot of its new
// voxel's segment; k_reorder computes the stable rank inside the segment from the sortkeys and
// moves the whole state B (dynamic fields) + A (static fields) -> Y.
// ------------------------------------------------------------------------------------------------
// n_dev != nullptr: the number of sources is read from device memory (crb_slab_step: the host never learns the
// counts between steps; the grid is sized by the capacity)
// Four sources per thread (one per 256-thread stripe of the CTA's 1,024: every access stays coalesced): eight
// independent loads, then four gathers of start_new[], then four stores — the pass is a chain of three dependent
// memory operations per source and nothing else, so the loads in flight per thread decide its speed (one source per
// thread: 0.040 ms at 8 M cells, a tenth of the rebuild).
constexpr int kFillPerThread = 4;
__global__ void __launch_bounds__(256) k_fill(const u32* __restrict__ newkey, const u32* __restrict__ rank_arb,
                                              const u32* __restrict__ start_new, u32* __restrict__ tmp_idx, u64 first, u64 n,
                                              const u64* __restrict__ n_dev = nullptr) {
    if (n_dev) n = *n_dev;
    const u64 t0 = (u64)blockIdx.x * (256 * kFillPerThread) + threadIdx.x;
    u32 k[kFillPerThread], r[kFillPerThread];
#pragma unroll
    for (int j = 0; j < kFillPerThread; ++j) {
        const u64 t = t0 + (u64)j * 256;
        k[j] = t < n ? __ldg(&newkey[first + t]) : CRB_KEY_GONE;
        r[j] = t < n ? __ldg(&rank_arb[first + t]) : 0u;
    }
    u32 sb[kFillPerThread];
#pragma unroll
    for (int j = 0; j < kFillPerThread; ++j) sb[j] = k[j] != CRB_KEY_GONE ? __ldg(&start_new[k[j]]) : 0u;  // GONE: migrated away
#pragma unroll
    for (int j = 0; j < kFillPerThread; ++j)
        if (k[j] != CRB_KEY_GONE) tmp_idx[sb[j] + r[j]] = (u32)(first + t0 + (u64)j * 256);
}

struct ReorderParams {
    State A, B, Y;
    const u32* newkey;
    const u32* start_new;
    const u32* tmp_idx;
    const u64* sortkey;
    u32* key_out;
    u32* slot_of_src;  // optional: destination slot of every source index (ingestion only)
    u64 n;      // destination slots [first, first + n)
    u64 first;
    int D, V;
    int n_mp, n_ip;  // per-agent parameter columns to move (0 when uniform)
    int save_prev;
    int has_species;
    double mn[3];  // CartesianCuboid::min (origin of the FP32 filter coordinates)
    u32 n2;        // voxels along the fastest axis
    u64 nvox_dev;  // != 0: the number of destinations is start_new[nvox_dev] - first (device-side counts)
};

// Two groups of loads (positions + velocities + scalars, then the increment history) instead of one: at most nine
// doubles are in flight per thread, which lets the kernel fit 40 registers without spills = 48 resident warps per SM
// instead of 32.  k_reorder is latency bound — occupancy is what buys bandwidth: 0.443 -> 0.407 ms at 8 M cells,
// 0.0657 -> 0.0594 ms at 1 M (one box, round 2).  Capping the one-group form at 48 / 40 registers spilled and lost
// (0.53 / 0.56 ms); four groups at 32 registers spill as well.
__global__ void __launch_bounds__(256, 6) k_reorder(const ReorderParams R) {
    const u64 tp = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 n_dst = R.nvox_dev ? (u64)__ldg(&R.start_new[R.nvox_dev]) - R.first : R.n;
    if (tp >= n_dst) return;
    const u64 p = R.first + tp;
    const u32 i = R.tmp_idx[p];
    // ---- group 1: issued BEFORE the rank loop and before any store (one long-scoreboard stall for all of them)
    double bp[3], bv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c < R.D) {
            bp[c] = __ldg(&R.B.pos[c][i]);
            bv[c] = __ldg(&R.B.vel[c][i]);
        }
    }
    const u8 hl = __ldg(&R.B.hlen[i]);
    const u64 id = __ldg(&R.A.id[i]);
    const u8 sp = R.has_species ? __ldg(&R.A.species[i]) : (u8)0;
    const u32 v = __ldg(&R.newkey[i]);
    const u64 my = __ldg(&R.sortkey[i]);
    const u32 sb = __ldg(&R.start_new[v]), se = __ldg(&R.start_new[v + 1]);
    // stable rank inside the voxel's segment
    u32 rank = 0;
    for (u32 q = sb; q < se; ++q) rank += (__ldg(&R.sortkey[__ldg(&R.tmp_idx[q])]) < my) ? 1u : 0u;
    const u64 d = (u64)sb + rank;
    // ---- group 2: the increment history, in flight while group 1 is stored
    double bx[3], bw[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c < R.D) {
            bx[c] = __ldg(&R.B.h1x[c][i]);
            bw[c] = __ldg(&R.B.h1v[c][i]);
        }
    }
    R.key_out[d] = v;
    if (R.slot_of_src) R.slot_of_src[i] = (u32)d;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c < R.D) {
            R.Y.pos[c][d] = bp[c];
            R.Y.vel[c][d] = bv[c];
        }
    }
    {
        const double p0 = bp[0], p1 = R.D >= 2 ? bp[1] : R.mn[1];
        const double p2 = R.D == 3 ? bp[2] : R.mn[2];
        R.Y.posf[d] = make_float4((float)(p0 - R.mn[0]), (float)(p1 - R.mn[1]), (float)(p2 - R.mn[2]),
                                  __uint_as_float(v));
        // species | (voxel index along the fastest axis & 0xff) << 8: what the drain of the tile kernels
        // needs besides the voxel index (one 4-byte word per staged candidate)
        R.Y.meta[d] = (u32)sp | (((v % R.n2) & 0xffu) << 8);
    }
    R.Y.hlen[d] = hl;
    R.Y.id[d] = id;
    if (R.has_species) R.Y.species[d] = sp;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c < R.D) {
            R.Y.h1x[c][d] = bx[c];
            R.Y.h1v[c][d] = bw[c];
        }
    }
    if (R.save_prev) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (c < R.D) {
                R.Y.h0x[c][d] = __ldg(&R.B.h0x[c][i]);
                R.Y.h0v[c][d] = __ldg(&R.B.h0v[c][i]);
            }
        }
    }
    for (int k = 0; k < R.n_mp; ++k) R.Y.mp[k][d] = R.A.mp[k][i];
    for (int k = 0; k < R.n_ip; ++k) R.Y.ip[k][d] = R.A.ip[k][i];
}

// ------------------------------------------------------------------------------------------------
// ingestion (crb_set_agents): agent-major host rows -> SoA columns, setup-time voxel index
// (CartesianCuboid::get_voxel_index_of_raw, cartesian_cuboid_n.rs:320-337)
// ------------------------------------------------------------------------------------------------
__global__ void k_unpack_rows(const double* __restrict__ rows, int width, double* c0, double* c1, double* c2, double* c3, double* c4,
                              double* c5, double* c6, double* c7, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* cols[8] = {c0, c1, c2, c3, c4, c5, c6, c7};
    for (int k = 0; k < width; ++k)
        if (cols[k]) cols[k][i] = rows[i * width + k];
}

__global__ void k_pack_rows(double* __restrict__ rows, int width, const double* c0,
                            const double* c1, const double* c2, const double* c3,
                            const double* c4, const double* c5, const double* c6,
                            const double* c7, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* cols[8] = {c0, c1, c2, c3, c4, c5, c6, c7};
    for (int k = 0; k < width; ++k) rows[i * width + k] = cols[k] ? cols[k][i] : 0.0;
}

template <int D>
__global__ void k_ingest_keys(const KParams P) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n) return;
    const u64 i = P.first + t;
    const Grid& g = P.g;
    double p[D];
    bool below = false;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        p[c] = P.B.pos[c][i];
        below |= (g.mn[c] >= p[c]);  // check_min_max(min, pos) (:149-162, 321)
    }
    int q0 = 0, q1 = 0, q2 = 0;
    u32 nk = 0;
    if (below || !(p[0] == p[0])) {
        latch_error(P.err, CRB_ERR_BOUNDARY, t);
    } else if (!voxel_of<D>(g, p, &q0, &q1, &q2)) {
        latch_error(P.err, CRB_ERR_INDEX, t);
    } else {
        const int l0 = q0 - g.p0;
        if (l0 < g.own_lo || l0 >= g.own_hi)
            latch_error(P.err, CRB_ERR_INDEX, t);  // not in this device's slab
        else
            nk = ((u32)l0 * (u32)g.ng[1] + (u32)q1) * (u32)g.ng[2] + (u32)q2;
    }
    P.newkey[i] = nk;
    P.sortkey[i] = t;  // input order inside a voxel (chili/datastructures.rs:278-299)
    P.rank_arb[i] = atomicAdd(&P.count_new[nk], 1u);
}

// ordered candidate pairs of the current state
__global__ void k_count_pairs(const KParams P, u64* out) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 i = P.first + t;
    u64 c = 0;
    if (t < P.n) {
        const Grid& g = P.g;
        const int n1 = g.ng[1], n2 = g.ng[2];
        const u32 v = P.key[i];
        const int i2 = (int)(v % (u32)n2);
        const u32 tq = v / (u32)n2;
        const int i1 = (int)(tq % (u32)n1), i0 = (int)(tq / (u32)n1);
        const int z_lo = i2 > 0 ? i2 - 1 : 0, z_hi = i2 + 1 < n2 ? i2 + 1 : n2 - 1;
        for (int d0 = -1; d0 <= 1; ++d0) {
            const int X = i0 + d0;
            if (X < 0 || X >= g.L0) continue;
            for (int d1 = -1; d1 <= 1; ++d1) {
                const int Y = i1 + d1;
                if (Y < 0 || Y >= n1) continue;
                const u32 row = ((u32)X * (u32)n1 + (u32)Y) * (u32)n2;
                c += P.start[row + z_hi + 1] - P.start[row + z_lo];
            }
        }
        c -= 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// migration tap: (id, old plain, new plain) of every agent flagged as moved in the last step
__global__ void k_collect_migrations(const u64* __restrict__ sortkey, const u32* __restrict__ newkey,
                                     const u64* __restrict__ id_old, Grid g, u64 first, u64 n, u64 cap,
                                     u64* ids, u64* oldv, u64* newv, u64* counter) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u64 i = first + t;
    const u64 sk = sortkey[i];
    if (!(sk & CRB_SK_MOVED)) return;
    const u64 slot = atomicAdd(counter, 1ull);
    if (slot >= cap) return;
    ids[slot] = id_old[i];
    oldv[slot] = CRB_SK_VOXEL(sk);
    newv[slot] = (u64)newkey[i] + (u64)g.p0 * g.ng[1] * g.ng[2];
}

// ------------------------------------------------------------------------------------------------
// slab exchange (replaces PosInformation / SendCell messages of
// chili/update_mechanics.rs:12-61 between subdomains that live on different devices)
// ------------------------------------------------------------------------------------------------
// boundary plane [v0, v0 + plane) of the sorted state -> halo buffer.  The agent range of a plane is
// contiguous, so every column is a straight copy.
__global__ void k_halo_export(const State A, const u32* __restrict__ start, u32 v0, HaloLayout L,
                              unsigned char* __restrict__ buf, int* err) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 a = start[v0], b = start[v0 + (u32)L.plane];
    const u64 n = b - a;
    if (t == 0) {
        reinterpret_cast<u64*>(buf)[0] = n;
        reinterpret_cast<u64*>(buf)[1] = L.mode;
        if (n > L.gcap) latch_error(err, CRB_ERR_CAPACITY, n);
    }
    if (t <= L.plane) reinterpret_cast<u32*>(buf + L.off_rel)[t] = start[v0 + (u32)t] - a;
    if (t < n && t < L.gcap) {
        const u64 i = a + t;
        for (int c = 0; c < L.D; ++c) reinterpret_cast<double*>(buf + L.off_pos[c])[t] = A.pos[c][i];
        reinterpret_cast<float4*>(buf + L.off_posf)[t] = A.posf[i];
        reinterpret_cast<u32*>(buf + L.off_meta)[t] = A.meta[i];
        (buf + L.off_sp)[t] = A.species[i];
        for (int k = 0; k < L.n_ip; ++k) reinterpret_cast<double*>(buf + L.off_ip[k])[t] = A.ip[k][i];
    }
}

// halo buffer -> ghost plane.  lower ghost: agents end at `edge` (= first owned index); upper ghost:
// agents begin at `edge` (= one past the last owned index).  posf.w is re-based to the local voxel
// index of the ghost plane (the sender wrote its own local index).
__global__ void k_halo_import(State A, u32* __restrict__ start, u32 v0, u64 edge, int lower, HaloLayout L,
                              const unsigned char* __restrict__ buf, int* err, const u64* __restrict__ n_dev = nullptr) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev && !lower) edge += *n_dev;  // upper ghost plane begins behind the owned agents: first + n
    const u64 n = reinterpret_cast<const u64*>(buf)[0];
    if (n > L.gcap) {
        if (t == 0) latch_error(err, CRB_ERR_CAPACITY, n);
        return;
    }
    // both sides of a face must agree on uniform vs per-agent parameters (and on the uniform row): the columns a
    // uniform sender leaves out would be read as stale memory by a per-agent receiver
    if (reinterpret_cast<const u64*>(buf)[1] != L.mode) {
        if (t == 0) latch_error(err, CRB_ERR_STATE, reinterpret_cast<const u64*>(buf)[1]);
        return;
    }
    const u64 base = lower ? edge - n : edge;
    const u32* rel = reinterpret_cast<const u32*>(buf + L.off_rel);
    if (t <= L.plane) start[v0 + t] = (u32)(base + rel[t]);
    if (t < n) {
        const u64 i = base + t;
        for (int c = 0; c < L.D; ++c) A.pos[c][i] = reinterpret_cast<const double*>(buf + L.off_pos[c])[t];
        // voxel of agent t: the last v with rel[v] <= t (binary search over the plane)
        u32 lo = 0, hi = (u32)L.plane;
        while (hi - lo > 1) {
            const u32 mid = (lo + hi) >> 1;
            if (rel[mid] <= t) lo = mid; else hi = mid;
        }
        float4 pf = reinterpret_cast<const float4*>(buf + L.off_posf)[t];
        pf.w = __uint_as_float(v0 + lo);
        A.posf[i] = pf;
        A.meta[i] = reinterpret_cast<const u32*>(buf + L.off_meta)[t];
        A.species[i] = (buf + L.off_sp)[t];
        for (int k = 0; k < L.n_ip; ++k) A.ip[k][i] = reinterpret_cast<const double*>(buf + L.off_ip[k])[t];
    }
}

// agents whose new voxel lies in a ghost plane leave this slab: stream-compacted with their complete
// state into the migration buffer of that side (chili: SendCell(old_voxel, CellBox, Aux),
// update_mechanics.rs:457-470).  The order inside the buffer is arbitrary; the receiver's reorder
// ranks arrivals by (source voxel, in-voxel rank) exactly like local movers (:492-511).
struct MigExportParams {
    State A, B;  // A: static fields (pre-step order), B: integrated dynamic fields
    u32* newkey;
    const u64* sortkey;
    u64 first, n;
    u32 lo_end;    // local voxels [0, lo_end) belong to the lower ghost plane (0 if none)
    u32 hi_begin;  // local voxels [hi_begin, ...) belong to the upper ghost plane (0xffffffff if none)
    u64 vox_base;  // global plain index of local voxel 0
    MigLayout L;
    unsigned char* buf[2];
    u64* counter;  // [2]
    int* err;
    int uniform;
    const u64* n_dev;  // != nullptr: owned agents are counted on the device (crb_slab_step)
};
__global__ void k_mig_export(const MigExportParams M) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {  // parameter mode word of the headers (the counts at byte 0 are copied in by the host code)
        if (M.buf[0]) reinterpret_cast<u64*>(M.buf[0])[1] = M.L.mode;
        if (M.buf[1]) reinterpret_cast<u64*>(M.buf[1])[1] = M.L.mode;
    }
    if (t >= (M.n_dev ? *M.n_dev : M.n)) return;
    const u64 i = M.first + t;
    const u32 k = M.newkey[i];
    int side;
    if (k < M.lo_end) side = 0;
    else if (k >= M.hi_begin && k != CRB_KEY_GONE) side = 1;
    else return;
    const u64 slot = atomicAdd(&M.counter[side], 1ull);
    M.newkey[i] = CRB_KEY_GONE;
    if (slot >= M.L.mcap) {
        latch_error(M.err, CRB_ERR_CAPACITY, M.A.id[i]);
        return;
    }
    unsigned char* buf = M.buf[side];
    const MigLayout& L = M.L;
    const int D = L.D;
    for (int c = 0; c < D; ++c) {
        reinterpret_cast<double*>(buf + L.off_dyn[0 * D + c])[slot] = M.B.pos[c][i];
        reinterpret_cast<double*>(buf + L.off_dyn[1 * D + c])[slot] = M.B.vel[c][i];
        reinterpret_cast<double*>(buf + L.off_dyn[2 * D + c])[slot] = M.B.h1x[c][i];
        reinterpret_cast<double*>(buf + L.off_dyn[3 * D + c])[slot] = M.B.h1v[c][i];
        reinterpret_cast<double*>(buf + L.off_dyn[4 * D + c])[slot] = M.B.h0x[c][i];
        reinterpret_cast<double*>(buf + L.off_dyn[5 * D + c])[slot] = M.B.h0v[c][i];
    }
    if (!M.uniform) {
        for (int q = 0; q < L.n_mp; ++q) reinterpret_cast<double*>(buf + L.off_mp[q])[slot] = M.A.mp[q][i];
        for (int q = 0; q < L.n_ip; ++q) reinterpret_cast<double*>(buf + L.off_ip[q])[slot] = M.A.ip[q][i];
    }
    reinterpret_cast<u64*>(buf + L.off_id)[slot] = M.A.id[i];
    reinterpret_cast<u64*>(buf + L.off_sortkey)[slot] = M.sortkey[i];
    reinterpret_cast<u64*>(buf + L.off_gvox)[slot] = M.vox_base + k;
    (buf + L.off_hlen)[slot] = M.B.hlen[i];
    (buf + L.off_sp)[slot] = M.A.species[i];
}

struct MigImportParams {
    State A, B;
    u32 *newkey, *rank_arb, *count_new;
    u64* sortkey;
    u64 dst;       // index of the first arrival in the unsorted source range
    u64 n;         // arrivals
    u64 vox_base;  // global plain index of local voxel 0
    u64 nvox;      // local voxels
    MigLayout L;
    const unsigned char* buf;
    int* err;
    int uniform;
};
__global__ void k_mig_import(const MigImportParams M) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M.n) return;
    if (reinterpret_cast<const u64*>(M.buf)[1] != M.L.mode) {
        if (t == 0) latch_error(M.err, CRB_ERR_STATE, reinterpret_cast<const u64*>(M.buf)[1]);
        M.newkey[M.dst + t] = CRB_KEY_GONE;
        return;
    }
    const u64 i = M.dst + t;
    const MigLayout& L = M.L;
    const unsigned char* buf = M.buf;
    const int D = L.D;
    for (int c = 0; c < D; ++c) {
        M.B.pos[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[0 * D + c])[t];
        M.B.vel[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[1 * D + c])[t];
        M.B.h1x[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[2 * D + c])[t];
        M.B.h1v[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[3 * D + c])[t];
        M.B.h0x[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[4 * D + c])[t];
        M.B.h0v[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[5 * D + c])[t];
    }
    if (!M.uniform) {
        for (int q = 0; q < L.n_mp; ++q) M.A.mp[q][i] = reinterpret_cast<const double*>(buf + L.off_mp[q])[t];
        for (int q = 0; q < L.n_ip; ++q) M.A.ip[q][i] = reinterpret_cast<const double*>(buf + L.off_ip[q])[t];
    }
    M.A.id[i] = reinterpret_cast<const u64*>(buf + L.off_id)[t];
    M.sortkey[i] = reinterpret_cast<const u64*>(buf + L.off_sortkey)[t];
    M.B.hlen[i] = (buf + L.off_hlen)[t];
    M.A.species[i] = (buf + L.off_sp)[t];
    const u64 gv = reinterpret_cast<const u64*>(buf + L.off_gvox)[t];
    if (gv < M.vox_base || gv - M.vox_base >= M.nvox) {
        latch_error(M.err, CRB_ERR_INDEX, M.A.id[i]);
        M.newkey[i] = CRB_KEY_GONE;
        return;
    }
    const u32 nk = (u32)(gv - M.vox_base);
    M.newkey[i] = nk;
    M.rank_arb[i] = atomicAdd(&M.count_new[nk], 1u);
}

// ------------------------------------------------------------------------------------------------
// crb_slab_step: exchange over peer memory without the host.  A sender's export kernels store
// straight into the neighbour's import buffers (mapped peer memory: NVLink / NVSwitch, or the same device) and then
// raise a step counter in the NEIGHBOUR's memory; the receiver's stream holds a one-thread kernel that spins on its
// own, local counter.  Replaces the barrier between the exchange phases of chili
// (cellular_raza-core/src/backend/chili/simulation_flow.rs:350-423) for two neighbours instead of all threads.
// ------------------------------------------------------------------------------------------------
// stores the migration counts into the headers of the neighbours' buffers, then publishes `value` to the flags;
// runs after the export kernels of the same stream (their stores are complete: stream order), the system-scope
// fence orders them before the flag for an observer on another device
__global__ void k_slab_publish(const u64* __restrict__ counts, unsigned char* hdr_lo, unsigned char* hdr_hi,
                               u64* flag_lo, u64* flag_hi, u64 value, u64* counts_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (counts) {
        if (hdr_lo) *reinterpret_cast<u64*>(hdr_lo) = counts[0];
        if (hdr_hi) *reinterpret_cast<u64*>(hdr_hi) = counts[1];
        if (counts_out) {
            counts_out[0] = counts[0];
            counts_out[1] = counts[1];
        }
    }
    __threadfence_system();
    if (flag_lo) *reinterpret_cast<volatile u64*>(flag_lo) = value;
    if (flag_hi) *reinterpret_cast<volatile u64*>(flag_hi) = value;
}

// spins until both local flags reached `value` (a neighbour that died would hang the stream: after `timeout_ns`
// the kernel latches an error and lets the stream continue, the caller then fails at its next synchronisation)
__global__ void k_slab_wait(const u64* flag_a, const u64* flag_b, u64 value, unsigned long long timeout_ns, int* err) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int k = 0; k < 2; ++k) {
        const u64* f = k == 0 ? flag_a : flag_b;
        if (!f) continue;
        unsigned spins = 0;
        while (true) {
            u64 v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= value) break;
            if ((++spins & 0x3ff) == 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > timeout_ns) {
                    latch_error(err, CRB_ERR_TIMEOUT, value);
                    return;
                }
            }
            __nanosleep(64);
        }
    }
    __threadfence_system();
}

// arrivals of both faces in one launch (blockIdx.y = side): counts come from the buffer headers, the destination
// range behind the owned agents from the device-side count; thread (0, 0) of side 0 leaves the new number of sources
struct MigImportDevParams {
    MigImportParams M;        // buf / dst / n are filled in on the device
    const unsigned char* buf[2];
    u64 first;                // index of the first owned agent (G)
    u64 src_cap;              // capacity of the unsorted source range (owned capacity + 2 migration capacities)
    u64* cnt;                 // [0] owned agents, [1] out: sources = owned + arrivals, [2..3] out: arrivals per face
};
__global__ void k_mig_import_dev(const MigImportDevParams Q) {
    const int side = blockIdx.y;
    const unsigned char* buf = Q.buf[side];
    const u64 n_own = Q.cnt[0];
    const u64 n_lo = Q.buf[0] ? *reinterpret_cast<const u64*>(Q.buf[0]) : 0ull;
    const u64 n_hi = Q.buf[1] ? *reinterpret_cast<const u64*>(Q.buf[1]) : 0ull;
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool overflow = n_lo > Q.M.L.mcap || n_hi > Q.M.L.mcap || n_own + n_lo + n_hi > Q.src_cap;
    if (t == 0 && side == 0) {
        Q.cnt[1] = overflow ? n_own : n_own + n_lo + n_hi;
        Q.cnt[2] = n_lo;
        Q.cnt[3] = n_hi;
        if (overflow) latch_error(Q.M.err, CRB_ERR_CAPACITY, n_lo + n_hi);
    }
    if (!buf || overflow) return;
    const u64 n = side == 0 ? n_lo : n_hi;
    if (t >= n) return;
    if (reinterpret_cast<const u64*>(buf)[1] != Q.M.L.mode) {
        if (t == 0) latch_error(Q.M.err, CRB_ERR_STATE, reinterpret_cast<const u64*>(buf)[1]);
        Q.M.newkey[Q.first + n_own + (side ? n_lo : 0) + t] = CRB_KEY_GONE;
        return;
    }
    const MigImportParams& M = Q.M;
    const u64 i = Q.first + n_own + (side ? n_lo : 0) + t;
    const MigLayout& L = M.L;
    const int D = L.D;
    for (int c = 0; c < D; ++c) {
        M.B.pos[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[0 * D + c])[t];
        M.B.vel[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[1 * D + c])[t];
        M.B.h1x[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[2 * D + c])[t];
        M.B.h1v[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[3 * D + c])[t];
        M.B.h0x[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[4 * D + c])[t];
        M.B.h0v[c][i] = reinterpret_cast<const double*>(buf + L.off_dyn[5 * D + c])[t];
    }
    if (!M.uniform) {
        for (int q = 0; q < L.n_mp; ++q) M.A.mp[q][i] = reinterpret_cast<const double*>(buf + L.off_mp[q])[t];
        for (int q = 0; q < L.n_ip; ++q) M.A.ip[q][i] = reinterpret_cast<const double*>(buf + L.off_ip[q])[t];
    }
    M.A.id[i] = reinterpret_cast<const u64*>(buf + L.off_id)[t];
    M.sortkey[i] = reinterpret_cast<const u64*>(buf + L.off_sortkey)[t];
    M.B.hlen[i] = (buf + L.off_hlen)[t];
    M.A.species[i] = (buf + L.off_sp)[t];
    const u64 gv = reinterpret_cast<const u64*>(buf + L.off_gvox)[t];
    if (gv < M.vox_base || gv - M.vox_base >= M.nvox) {
        latch_error(M.err, CRB_ERR_INDEX, M.A.id[i]);
        M.newkey[i] = CRB_KEY_GONE;
        return;
    }
    const u32 nk = (u32)(gv - M.vox_base);
    M.newkey[i] = nk;
    M.rank_arb[i] = atomicAdd(&M.count_new[nk], 1u);
}

// crb_slab_step queues kernels only (no copy-engine operations): a memset or copy of ANOTHER slab's stream can sit
// behind one that waits for a spinning k_slab_wait in the copy engine's queue (observed: two slabs on one device
// serialised completely), so the small clears and the count upload are kernels as well
__global__ void k_zero_u32(u32* __restrict__ a, u64 n, u32* __restrict__ b, u64 nb, u64* __restrict__ c, u64 nc) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = 0u;
    if (i < nb) b[i] = 0u;
    if (i < nc) c[i] = 0ull;
}
// asynchronous API (crb_set_agents_async / crb_step_async / crb_snapshot_async): the small clears and device copies
// of a handle are kernels, not copy-engine operations — a memset queued by one handle otherwise sits in the copy
// engine's queue behind the multi-millisecond host transfer of the handle it alternates with, and holds up the
// kernels that wait for it (the e2e pipeline lost ~2 ms per step to that)
struct ZeroCols {
    double* col[16];
    int n_cols;
    u8* bytes_a;  // optional byte arrays of n entries
    u8* bytes_b;
};
__global__ void k_zero_cols(const ZeroCols Z, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < Z.n_cols; ++k) Z.col[k][i] = 0.0;
    if (Z.bytes_a) Z.bytes_a[i] = 0;
    if (Z.bytes_b) Z.bytes_b[i] = 0;
}
__global__ void k_copy_u8(u8* __restrict__ dst, const u8* __restrict__ src, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_copy_u64(u64* __restrict__ dst, const u64* __restrict__ src, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_set_u64(u64* a, u64 v) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *a = v;
}

// after the reorder: the owned agents of the next step = grand total of the scan; capacity check
__global__ void k_slab_set_count(const u32* __restrict__ start_new, u64 nvox, u64 first, u64 cap, u64* cnt, int* err) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const u64 n = (u64)start_new[nvox] - first;
    cnt[0] = n;
    if (n > cap) latch_error(err, CRB_ERR_CAPACITY, n);
}

// DFMA-only kernel: 8 independent chains per thread, 2 flop per fma.  Roofline denominator of the
// force kernel.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
           x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    const double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (r == 123.456) out[0] = r;
}

__global__ void k_rsqrt_test(const double* x, double* y, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = rsqrt_nr(x[i]);
}
__global__ void k_exp_test(const double* x, double* y, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = exp_bounded(x[i]);
}
__global__ void k_iota_u64(u64* a, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
__global__ void k_fill_f64(double* a, double v, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
__global__ void k_global_keys(const u32* key, u64* out, u64 base, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u64)key[i] + base;
}

#endif  // CRB_TILE_TU

}  // namespace crb
