// oracle/chili_common.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Shared pieces of the CPU restatement of cellular_raza's `chili` mechanics/interaction path:
// geometry of CartesianCuboid, chili's subdomain chunking, the Xapy helpers, the ring buffer of
// AuxStorageMechanics and the solver helper formulas.  Every function cites the reference
// file:line it follows (paths relative to /root/reference).  Nothing here is linked into or
// imported by the product library (cellular_raza_b200/); only tests/, bench.py's cpu_baseline /
// --impl reference legs and __graft_entry__.smoke() use it, as the checker.
//
// Parity status: the reference (Rust) cannot be compiled in this environment and ships no golden
// vectors for pair forces, cross-voxel weights or the AB wrapper recurrence, so those parts are
// "parity unpinned" (defined by the source text restated here).  What the reference's own tests
// DO pin (voxel counts, get_decomp_res table + invariants, ring-buffer order, solver helper
// error bounds, Xapy, nearest point, free motion, determinism) is checked in tests/test_oracle_*.py.
//
// Build with -O2 -ffp-contract=off: Rust never contracts a*b+c into an fma.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../include/crb.h"

namespace chili {

constexpr int MAXD = 3;

// parameter-row widths of the C ABI (include/crb.h); restated here so that the oracle does not
// link the product library.
inline uint32_t n_mech_params(uint32_t kind) {
    switch (kind) {
        case CRB_MECH_NEWTON_DAMPED: return 2;
        case CRB_MECH_BROWNIAN: return 2;
        case CRB_MECH_LANGEVIN: return 3;
        case CRB_MECH_ROD: return 8;
    }
    return 0;
}
inline uint32_t n_inter_params(uint32_t kind) {
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

// ---- num-traits ToPrimitive::to_usize for f64 -------------------------------------------------
// num-traits 0.2 `float_to_unsigned_int_impl`: Some(x as usize) iff -1.0 < x < 2^64, truncating
// toward zero; None otherwise (NaN included).  Call sites: cartesian_cuboid_n.rs:325, 478.
inline bool to_usize(double x, uint64_t* out) {
    if (x > -1.0 && x < 18446744073709551616.0) {
        *out = (x <= 0.0) ? 0 : (uint64_t)x;  // (-1,0) truncates to 0
        return true;
    }
    return false;
}

// ---- get_decomp_res (cellular_raza-building-blocks/src/domains/cartesian_cuboid_n.rs:19-59) ---
inline bool get_decomp_res(uint64_t n_voxel, uint64_t n_regions, uint64_t* n_out, uint64_t* m_out,
                           uint64_t* avg_out) {
    int64_t average_len = (int64_t)std::ceil((double)n_voxel / (double)n_regions);
    int64_t n = (int64_t)n_regions;
    int64_t m = 0;
    for (uint64_t it = 0; it < n_regions; ++it) {
        int64_t residue = (int64_t)n_voxel - average_len * n - (average_len - 1) * m;
        if (residue == 0) {
            *n_out = (uint64_t)n;
            *m_out = (uint64_t)m;
            *avg_out = (uint64_t)average_len;
            return true;
        } else if (residue > 0) {
            if (n == (int64_t)n_regions) {
                average_len += 1;
                n = (int64_t)n_regions;
                m = 0;
            }
        } else {
            n -= 1;
            m += 1;
        }
    }
    return false;
}

// ---- CartesianCuboid -------------------------------------------------------------------------
struct Domain {
    int dim = 3;
    double min[MAXD] = {0, 0, 0}, max[MAXD] = {0, 0, 0}, dx[MAXD] = {1, 1, 1};
    uint64_t n_voxels[MAXD] = {1, 1, 1};
    uint64_t n_total() const {
        uint64_t r = 1;
        for (int i = 0; i < dim; ++i) r *= n_voxels[i];
        return r;
    }
    // plain index = position in the enumeration of multi_cartesian_product, last axis fastest
    // (cartesian_cuboid_n.rs:260-270 + chili/datastructures.rs:113-121)
    uint64_t plain(const uint64_t* idx) const {
        uint64_t p = 0;
        for (int i = 0; i < dim; ++i) p = p * n_voxels[i] + idx[i];
        return p;
    }
    void unplain(uint64_t p, uint64_t* idx) const {
        for (int i = dim - 1; i >= 0; --i) {
            idx[i] = p % n_voxels[i];
            p /= n_voxels[i];
        }
    }
};

// CartesianCuboid::from_boundaries_and_interaction_range (cartesian_cuboid_n.rs:183-222)
inline bool domain_from_interaction_range(int dim, const double* min, const double* max,
                                          double range, Domain* d) {
    d->dim = dim;
    for (int i = 0; i < dim; ++i) {
        if (min[i] >= max[i]) return false;  // check_min_max :149-162
        double n = std::floor((max[i] - min[i]) / range);
        uint64_t m;
        if (!to_usize(n, &m)) return false;
        d->n_voxels[i] = m > 1 ? m : 1;
        d->dx[i] = (max[i] - min[i]) / n;
        d->min[i] = min[i];
        d->max[i] = max[i];
    }
    return true;
}

// CartesianCuboid::from_boundaries_and_n_voxels (cartesian_cuboid_n.rs:226-256)
inline bool domain_from_n_voxels(int dim, const double* min, const double* max,
                                 const uint64_t* n_voxels, Domain* d) {
    d->dim = dim;
    for (int i = 0; i < dim; ++i) {
        if (min[i] >= max[i]) return false;
        d->n_voxels[i] = n_voxels[i];
        d->dx[i] = (max[i] - min[i]) / (double)n_voxels[i];
        d->min[i] = min[i];
        d->max[i] = max[i];
    }
    return true;
}

// CartesianSubDomain::get_index_of (cartesian_cuboid_n.rs:469-490): no bounds check besides
// to_usize.  Returns false for the BoundaryError case.
inline bool get_index_of(const Domain& d, const double* pos, uint64_t* idx) {
    for (int i = 0; i < d.dim; ++i) {
        double n_vox = (pos[i] - d.min[i]) / d.dx[i];
        if (!to_usize(n_vox, &idx[i])) return false;
    }
    return true;
}

// CartesianCuboid::get_voxel_index_of_raw (cartesian_cuboid_n.rs:320-337): setup-time twin with
// the additional check_min_max(min, pos).
inline bool get_voxel_index_of_raw(const Domain& d, const double* pos, uint64_t* idx) {
    for (int i = 0; i < d.dim; ++i)
        if (d.min[i] >= pos[i]) return false;
    return get_index_of(d, pos, idx);
}

// true when idx is a key of voxel_index_to_plain_index (chili/datastructures.rs:113-121)
inline bool index_in_domain(const Domain& d, const uint64_t* idx) {
    for (int i = 0; i < d.dim; ++i)
        if (idx[i] >= d.n_voxels[i]) return false;
    return true;
}

// get_neighbor_voxel_indices (cartesian_cuboid_n.rs:641-662) mapped to plain indices and stored
// in a BTreeSet (chili/datastructures.rs:150-160) => ascending plain order, self excluded.
inline void neighbor_plain_indices(const Domain& d, uint64_t plain, std::vector<uint64_t>* out) {
    out->clear();
    uint64_t idx[MAXD] = {0, 0, 0};
    d.unplain(plain, idx);
    uint64_t lo[MAXD] = {0, 0, 0}, hi[MAXD] = {1, 1, 1};
    for (int i = 0; i < d.dim; ++i) {
        lo[i] = idx[i] > 0 ? idx[i] - 1 : 0;
        hi[i] = idx[i] + 2 < d.n_voxels[i] ? idx[i] + 2 : d.n_voxels[i];
    }
    uint64_t c[MAXD];
    for (c[0] = lo[0]; c[0] < hi[0]; ++c[0])
        for (c[1] = (d.dim > 1 ? lo[1] : 0); c[1] < (d.dim > 1 ? hi[1] : 1); ++c[1])
            for (c[2] = (d.dim > 2 ? lo[2] : 0); c[2] < (d.dim > 2 ? hi[2] : 1); ++c[2]) {
                uint64_t p = d.plain(c);
                if (p != plain) out->push_back(p);
            }
}

// chili's subdomain map: create_subdomains (cartesian_cuboid_n.rs:500-583) chunks the plain
// enumeration into n chunks of average_len followed by m chunks of average_len-1.
struct SubdomainMap {
    uint64_t n_sub = 1, n = 1, m = 0, avg = 1, switcher = 0;
    bool init(uint64_t n_voxel, uint64_t n_regions) {
        n_sub = n_regions;
        if (!get_decomp_res(n_voxel, n_regions, &n, &m, &avg)) return false;
        switcher = n * avg;
        return true;
    }
    uint64_t sub_of(uint64_t plain) const {
        if (plain < switcher) return plain / avg;
        uint64_t a = avg - 1 > 1 ? avg - 1 : 1;
        return (plain - switcher) / a + n;
    }
    // first plain index of subdomain s (s may be n_sub => one past the end)
    uint64_t first_voxel(uint64_t s) const {
        if (s <= n) return s * avg;
        return switcher + (s - n) * (avg - 1);
    }
};

// ---- Xapy (cellular_raza-concepts/src/reactions.rs:61-82): self * a + y, no fma ----------------
inline double xapy(double x, double a, double y) { return x * a + y; }
inline double xa(double x, double a) { return x * a; }

// ---- solver helpers (chili/solvers.rs:303-347) ------------------------------------------------
inline double helper_euler(double x, double dx, double dt, double dx_rand) {
    return xapy(xapy(dx, dt, x), 1.0, xa(dx_rand, dt));
}
inline double helper_ab2(double x, double dx0, double dx1, double dt, double dx_rand) {
    const double f0 = 3.0 / 2.0;
    const double f1 = -1.0 / 2.0;
    return xapy(xapy(xapy(dx0, f0, xa(dx1, f1)), dt, x), 1.0, xa(dx_rand, dt));
}
inline double helper_ab3(double x, double dx0, double dx1, double dx2, double dt, double dx_rand) {
    const double f0 = 23.0 / 12.0;
    const double f1 = -16.0 / 12.0;
    const double f2 = 5.0 / 12.0;
    return xapy(xapy(xapy(dx0, f0, xapy(dx1, f1, xa(dx2, f2))), dt, x), 1.0, xa(dx_rand, dt));
}

// ---- RingBuffer<T, N> of circ_buffer 0.1.8 as used by AuxStorageMechanics ----------------------
// (chili/aux_storage.rs:41-130).  push appends, dropping the oldest entry when full; iteration
// is oldest -> newest (pinned by the doctest at chili/aux_storage.rs:862-868).
template <int W>
struct Ring {
    double v[2][W];
    uint32_t len = 0;
    void push(const double* x, uint32_t cap) {
        if (cap == 0) return;
        if (len < cap) {
            std::memcpy(v[len], x, sizeof(double) * W);
            ++len;
        } else {
            for (uint32_t k = 1; k < cap; ++k) std::memcpy(v[k - 1], v[k], sizeof(double) * W);
            std::memcpy(v[cap - 1], x, sizeof(double) * W);
        }
    }
};

// Rust f64::signum: 1.0 for +0.0 and positives, -1.0 for -0.0 and negatives, NaN for NaN.
inline double rust_signum(double x) {
    if (std::isnan(x)) return x;
    return std::signbit(x) ? -1.0 : 1.0;
}

}  // namespace chili
