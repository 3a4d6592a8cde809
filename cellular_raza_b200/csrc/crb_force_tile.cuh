// crb_force_tile.cuh — shared-memory tiled neighbour-force kernel (uniform parameters), sm_100a.
//
// Replaces SubDomainBox::update_mechanics_interaction_step_1/2/3
// (cellular_raza-core/src/backend/chili/update_mechanics.rs:153-238, 285-382) +
// local_mechanics_update (:530-559) + apply_boundary (:387-406) for the closed model set.
//
// One CTA owns a pencil of voxels: one (a0, a1) row of the voxel grid and a segment [za, zb) of the
// fastest axis.  Because agents are sorted by VoxelPlainIndex, the candidates of the whole pencil are
// <= 9 (3-D) / 3 (2-D) CONTIGUOUS index ranges (one per neighbouring row, voxels za-1 .. zb), which
// are staged once in shared memory:
//     f64  sx, sy, sz        exact positions (used by the potential)
//     f32x4 sposf            {x - min_x, y - min_y, z - min_z, voxel index}  (used by the filter)
//     u8   ssp               species
// Work of one warp = 32 consecutive target agents of the pencil:
//   filter  every staged candidate of the (at most 9 x (zspan+2)) voxels around the warp's targets is
//           broadcast from shared memory (one LDS.128 per candidate per warp) and tested against the
//           32 targets at once in FP32: conservative distance test (cutoff inflated by the FP32
//           rounding bound of the coordinates) AND voxel adjacency (chili only pairs agents of
//           neighbouring voxels).  The result is one bit per (target, candidate) in a 32-candidate
//           mask word per lane.
//   drain   every lane walks the set bits of its own masks in ascending candidate order and evaluates
//           the potential in FP64 exactly as the per-agent kernel does (same coef(), same pair
//           weights), accumulating in registers.  A candidate the filter rejected has coefficient
//           exactly zero (every potential of the model set is zero beyond its cutoff), so skipping it
//           does not change the result; the summation order per target is ascending candidate
//           index, a pure function of the sorted state => bitwise run-to-run determinism.
//   fused epilogue: integrate_agent() (increment, AB/Euler, boundary, new key, histogram).
// If the candidates of a pencil exceed the staging capacity the pencil is processed in several
// z-windows; a single voxel whose 3^D neighbourhood alone exceeds it takes the per-agent path
// (force_from_global) inside the same kernel.
#pragma once
#include "crb_kernels.cuh"

namespace crb {

constexpr int kTileRows = 9;
constexpr int kTileQueue = 56;      // in-range candidate indices buffered per lane between two drains
constexpr int kTileThreads = 128;
constexpr int kTileSub = 4;         // lanes per candidate-window sub-group
constexpr int kTileRowPad = 8;      // sentinel entries staged after every row (filter over-read)

struct TileLaunch {
    int chunk_len;  // voxels of the fastest axis per chunk (bounds the window table)
    int nchunks;    // chunks per row
    int nseg;       // count-balanced segments per chunk (CTAs beyond the chunk's population exit)
    u32 tseg;       // target agents per segment
    int x_begin;    // first local plane whose agents are targets
    int x_count;    // number of such planes
    u32 cap;        // staged-candidate capacity (incl. sentinels)
    int kmax;       // max voxels of a window + 1 (row stride of the wst table)
    float cut2f;    // conservative FP32 cutoff^2 of the filter
    float cut2m;    // conservative cutoff^2 of the TF32 tensor-core filter (k_force_tile3)
    u32 inv_nseg;   // ceil(2^32 / nseg): blockIdx.x / nseg == __umulhi(blockIdx.x, inv_nseg) (k_force_tile2, set by its launcher)
};

// exact data of one staged candidate: one address computation and two LDS.128 per evaluated pair
struct __align__(16) TileCand {
    double x, y, z;
    u32 meta;  // species | (z voxel & 0xff) << 8
    u32 cv;    // local voxel index
};

__host__ __device__ inline size_t tile_smem_bytes(u32 cap, int kmax, int threads) {
    const size_t c = cap + 34;  // cap includes kTileRows * kTileRowPad sentinels; +34 filter over-read
    size_t b = 0;
    b += c * sizeof(TileCand);                                         // exact positions + meta + voxel
    b += c * sizeof(float4);                                           // FP32 filter copy (pair layout)
    b += (size_t)kTileRows * (kmax + 1) * sizeof(u32);                 // start[] of the 9 rows
    b += (size_t)(kTileQueue + 4) * threads * sizeof(unsigned short);  // per-lane FIFOs
    return b + 64;
}

// per-target state of the drain
template <int D>
struct TileTarget {
    double x[D];
    double acc[D];
    u32 my_v;       // local voxel index
    u32 subv_lo;    // local voxel range of the target's chili subdomain (len 0 => weights 1)
    u32 subv_len;
    u32 zw;         // voxel index of the target along the fastest axis, low 8 bits
    u8 sp;
    int calc_flag;
};

// The exact record of candidate j is two 16-byte halves {x, y} = ra[j << SHIFT] and {z, meta | voxel << 32} =
// rb[j << SHIFT]: SHIFT 1 with ra = sc, rb = sc + 16 B for the 32-byte TileCand records of k_force_tile /
// k_force_tile3, SHIFT 0 for the two arrays of k_force_tile2.
template <int D, int INTER, int NE, int SHIFT>
__device__ __forceinline__ void tile_eval_rec(TileTarget<D>& t, const Pot<INTER>& own, const double ext_radius,
                                              const double2* __restrict__ ra, const double2* __restrict__ rb,
                                              const u32 (&jj)[NE], const bool (&hh)[NE]) {
    // straight-line code for NE candidates (coef() is branch-free) so that the dependent FP64 chains
    // interleave; an empty slot points at the target itself (z == 0) and contributes +0.
    double z[NE][D], cf[NE];
    int flag[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const double2 c0 = ra[jj[e] << SHIFT], c1 = rb[jj[e] << SHIFT];
        z[e][0] = t.x[0] - c0.x;
        z[e][1] = t.x[1] - c0.y;
        if (D == 3) z[e][2] = t.x[2] - c1.x;
        double r2 = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) r2 = fma(z[e][c], z[e][c], r2);
        const u32 mj = (u32)__double2loint(c1.y);
        const u32 cvj = (u32)__double2hiint(c1.y);
        const bool same_voxel = cvj == t.my_v;
        const bool same_sub = (u32)(cvj - t.subv_lo) < t.subv_len;
        const bool keep = same_voxel | !same_sub;
        // chili only pairs agents of neighbouring voxels: the candidate rows are x/y-adjacent by
        // construction, the fastest axis is checked here (the FP32 filter tests the distance only)
        const bool adj = (((mj >> 8) - t.zw + 1u) & 0xffu) <= 2u;
        const bool same_species = Pot<INTER>::kUsesSpecies ? ((mj & 0xffu) == (u32)t.sp) : true;
        flag[e] = 0;
        const double cc = coef_tile<INTER>(own, r2, ext_radius, same_species, &flag[e]);
        const bool use = hh[e] & adj;
        const double w = keep ? 1.0 : 0.5;
        cf[e] = use ? cc * w : 0.0;
        if (!use) flag[e] = 0;
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        t.calc_flag |= flag[e];
#pragma unroll
        for (int c = 0; c < D; ++c) t.acc[c] = fma(z[e][c], cf[e], t.acc[c]);
    }
}

template <int D, int INTER, int NE>
__device__ __forceinline__ void tile_eval(TileTarget<D>& t, const Pot<INTER>& own, const double ext_radius,
                                          const TileCand* __restrict__ sc,
                                          const u32 (&jj)[NE], const bool (&hh)[NE]) {
    const double2* r = reinterpret_cast<const double2*>(sc);
    tile_eval_rec<D, INTER, NE, 1>(t, own, ext_radius, r, r + 1, jj, hh);
}
template <int D, int INTER, int NE>
__device__ __forceinline__ void tile_eval_soa(TileTarget<D>& t, const Pot<INTER>& own, const double ext_radius,
                                              const double2* __restrict__ sxyd, const double2* __restrict__ szm,
                                              const u32 (&jj)[NE], const bool (&hh)[NE]) {
    tile_eval_rec<D, INTER, NE, 0>(t, own, ext_radius, sxyd, szm, jj, hh);
}

// per-agent path over global memory (same arithmetic as k_force_integrate, uniform parameters)
template <int D, int INTER>
__device__ __forceinline__ void force_from_global(const KParams& P, const u64 i, const u32 v, const int i0,
                                               const int i1, const int i2, const double* xi,
                                               double* acc) {
    const Grid& g = P.g;
    const int n1 = g.ng[1], n2 = g.ng[2];
    Pot<INTER> own;
    own.init(P.u.ip);
    const u8 spi = Pot<INTER>::kUsesSpecies ? P.A.species[i] : 0;
    const u32 own_lo = P.start[v], own_hi = P.start[v + 1];
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
    int calc_flag = 0;
    const int z_lo = i2 > 0 ? i2 - 1 : 0;
    const int z_hi = i2 + 1 < n2 ? i2 + 1 : n2 - 1;
    for (int d0 = -1; d0 <= 1; ++d0) {
        const int X = i0 + d0;
        if (X < 0 || X >= g.L0) continue;
        for (int d1 = -1; d1 <= 1; ++d1) {
            const int Y = i1 + d1;
            if (Y < 0 || Y >= n1) continue;
            const u32 row = ((u32)X * (u32)n1 + (u32)Y) * (u32)n2;
            const u32 jb = P.start[row + z_lo], je = P.start[row + z_hi + 1];
            for (u32 j = jb; j < je; ++j) {
                if (j == (u32)i) continue;
                double z[D];
                double r2 = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    z[c] = xi[c] - __ldg(&P.A.pos[c][j]);
                    r2 = fma(z[c], z[c], r2);
                }
                const bool same_voxel = (j >= own_lo) & (j < own_hi);
                const bool same_sub = (j >= sub_lo) & (j < sub_hi);
                const bool keep = same_voxel | !same_sub;
                const u8 spj = Pot<INTER>::kUsesSpecies ? __ldg(&P.A.species[j]) : 0;
                double cf = coef_uniform<INTER>(own, r2, P.u.ip[0], spi == spj, &calc_flag);
                cf = keep ? cf : 0.5 * cf;
#pragma unroll
                for (int c = 0; c < D; ++c) acc[c] = fma(z[c], cf, acc[c]);
            }
        }
    }
    if (calc_flag) latch_error(P.err, CRB_ERR_CALC, P.A.id[i]);
}

__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}

template <int D, int INTER, bool FORCE_ONLY>
__global__ void __launch_bounds__(kTileThreads, 3) k_force_tile(const KParams P, const TileLaunch T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int nthreads = kTileThreads;  // launch contract: blockDim.x == kTileThreads
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = nthreads >> 5;
    constexpr int NE = 4;  // candidates evaluated per drain iteration and lane
    const Grid& g = P.g;
    const int n1 = g.ng[1], n2 = g.ng[2];

    // ---- carve shared memory -----------------------------------------------------------------------
    const u32 capp = T.cap + 34;
    TileCand* sc = reinterpret_cast<TileCand*>(smem_raw);
    // FP32 filter copy, two candidates per 2 x 16 B so that one LDS.128 feeds a packed f32x2 op:
    //   sf[2p] = {x[2p], x[2p+1], y[2p], y[2p+1]},  sf[2p+1] = {z[2p], z[2p+1], cv[2p], cv[2p+1]}
    // (staged as one float4 {x, y, z, cv} per candidate by cp.async, then transposed pairwise in place)
    float4* sf = reinterpret_cast<float4*>(sc + capp);
    u32* sst = reinterpret_cast<u32*>(sf + capp);                // [kTileRows][kmax + 1] copy of start[]
    unsigned short* q = reinterpret_cast<unsigned short*>(sst + kTileRows * (T.kmax + 1));  // [kTileQueue + NE][nthreads]
    __shared__ u32 s_rowvox[kTileRows];   // first voxel (local plain index) of every neighbour row
    __shared__ u32 s_g0[kTileRows];       // first global agent index staged per row
    __shared__ u32 s_rs[kTileRows + 1];   // staged index of the first candidate of every row
    __shared__ u32 s_rn[kTileRows];       // candidates staged per row (without sentinels)
    __shared__ int s_valid[kTileRows];
    __shared__ int s_kc, s_za, s_zb;

    // ---- which pencil: row, chunk of the fastest axis, count-balanced segment of the chunk -----------
    const int seg = blockIdx.x % T.nseg;
    const int chunk = (blockIdx.x / T.nseg) % T.nchunks;
    const int rowid = blockIdx.x / (T.nseg * T.nchunks);
    const int Y = rowid % n1;
    const int X = T.x_begin + rowid / n1;
    const int ca = chunk * T.chunk_len;
    const int cb = min(ca + T.chunk_len, n2);
    {
        // CTAs beyond the segments this row needs leave before anything is staged (two broadcast loads)
        const u32 rvc = ((u32)X * (u32)n1 + (u32)Y) * (u32)n2;
        const u32 R = __ldg(P.start + rvc + cb) - __ldg(P.start + rvc + ca);
        if ((u32)seg * T.tseg >= R) return;  // seg >= ceil(R / tseg); uniform
    }
    if (tid < kTileRows) {
        const int d0 = tid / 3 - 1, d1 = tid % 3 - 1;
        const int XX = X + d0, YY = Y + d1;
        const bool ok = XX >= 0 && XX < g.L0 && YY >= 0 && YY < n1 && (D == 3 || d1 == 0);
        s_valid[tid] = ok ? 1 : 0;
        s_rowvox[tid] = ok ? ((u32)XX * (u32)n1 + (u32)YY) * (u32)n2 : 0u;
    }
    if (tid == 0) {
        s_za = cb;
        s_zb = cb;
    }
    __syncthreads();
    const u32 rv_c = s_rowvox[4];
    // ---- ONE round of global loads: start[] of the 9 rows over the chunk +- 1 voxel -> shared -------
    const int zb0 = max(ca - 1, 0), ztop = min(cb + 1, n2);
    const int sstride = T.kmax + 1;  // >= ztop - zb0 + 1
    {
        const int ns = ztop - zb0 + 1;
        for (int idx = tid; idx < kTileRows * ns; idx += nthreads) {
            const int r = idx / ns, k = idx - r * ns;
            sst[r * sstride + k] = s_valid[r] ? P.start[s_rowvox[r] + (u32)(zb0 + k)] : 0u;
        }
    }
    __syncthreads();
#define SS(r, zz) sst[(r) * sstride + ((zz) - zb0)]
    {
        // segment k owns the voxels whose first agent has in-chunk index in [k, k+1) * tsr, balanced per
        // row (ceil(R / tseg) segments of equal size): every CTA gets <= ~tseg targets whatever the
        // local density
        const u32 r0 = SS(4, ca);
        const u32 R = SS(4, cb) - r0;
        const u32 nsr = min((R + T.tseg - 1) / T.tseg, (u32)T.nseg);
        if ((u32)seg >= nsr) return;  // uniform: no barrier is pending
        const u32 tsr = (R + nsr - 1) / nsr;
        const u32 lo_cnt = (u32)seg * tsr;
        const u32 hi_cnt = (u32)seg == nsr - 1 ? 0xffffffffu : lo_cnt + tsr;
        for (int zz = ca + tid; zz < cb; zz += nthreads) {
            const u32 c = SS(4, zz) - r0;
            if (c >= lo_cnt) atomicMin(&s_za, zz);
            if (c >= hi_cnt) atomicMin(&s_zb, zz);
        }
    }
    __syncthreads();
    const int za = s_za, zb = s_zb;
    if (za >= zb) return;
    if (SS(4, zb) == SS(4, za)) return;  // no targets

    Pot<INTER> own;
    own.init(P.u.ip);
    const double ext_radius = P.u.ip[0];

    int z = za;
    while (z < zb) {
        // ---- window size: largest kc with all candidates of voxels [z-1, z+kc] staged ---------------
        __syncthreads();
        if (tid == 0) s_kc = 0;
        __syncthreads();
        {
            int found = 0;
            for (int kk = zb - z - tid; kk >= 1 && !found; kk -= nthreads) {
                const int wa = max(z - 1, 0), wb = min(z + kk + 1, n2);
                u32 total = 0;
#pragma unroll
                for (int r = 0; r < kTileRows; ++r) total += SS(r, wb) - SS(r, wa);
                if (total + kTileRows * kTileRowPad <= T.cap) {
                    atomicMax(&s_kc, kk);
                    found = 1;
                }
            }
        }
        __syncthreads();
        const int kc = s_kc;
        if (kc == 0) {
            // the neighbourhood of voxel z alone exceeds the staging capacity: per-agent path
            const u32 v = rv_c + (u32)z;
            const u32 t0 = SS(4, z), t1 = SS(4, z + 1);
            for (u32 i = t0 + tid; i < t1; i += nthreads) {
                double xi[D], acc[D];
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    xi[c] = P.A.pos[c][i];
                    acc[c] = 0.0;
                }
                force_from_global<D, INTER>(P, i, v, X, Y, z, xi, acc);
                if constexpr (FORCE_ONLY) {
                    add_domain_force<D>(P, xi, acc);
#pragma unroll
                    for (int c = 0; c < D; ++c) P.force_out[(u64)c * P.cap + i] = acc[c];
                } else {
                    integrate_agent<D, true>(P, i, v, t0, X, Y, z, xi, acc, load_agent_dyn<D>(P, i));
                }
            }
            z += 1;
            continue;
        }
        const int wa = max(z - 1, 0), wb = min(z + kc + 1, n2);

        // ---- row ranges -------------------------------------------------------------------------------
        if (warp == 0) {
            u32 a = 0, n = 0;
            if (lane < kTileRows) {
                a = SS(lane, wa);
                n = SS(lane, wb) - a;
            }
            const u32 np = lane < kTileRows ? n + kTileRowPad : 0u;  // row + sentinels
            u32 incl = np;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane < kTileRows) {
                s_g0[lane] = a;
                s_rs[lane] = incl - np;
                s_rn[lane] = n;
            }
            if (lane == kTileRows - 1) s_rs[kTileRows] = incl;
        }
        __syncthreads();

        // ---- stage the candidates: asynchronous copies, nothing waits until all are in flight ----------
        {
            const u32 total = s_rs[kTileRows];
            const u32 tend = (total + 33) & ~1u;  // even, >= total + 32: whole pairs of sentinels
            const float qn_ = __uint_as_float(0x7fc00000u);
            int r = 0;
            for (u32 k = tid; k < tend; k += nthreads) {
                bool real = false;
                if (k < total) {
                    while (k >= s_rs[r + 1]) ++r;
                    const u32 o = k - s_rs[r];
                    if (o < s_rn[r]) {
                        const u32 gi = s_g0[r] + o;
                        cp_async_8(&sc[k].x, P.A.pos[0] + gi);
                        cp_async_8(&sc[k].y, P.A.pos[1] + gi);
                        if (D == 3) cp_async_8(&sc[k].z, P.A.pos[2] + gi);
                        cp_async_16(sf + k, P.A.posf + gi);
                        cp_async_4(&sc[k].meta, P.A.meta + gi);
                        real = true;
                    }
                }
                if (!real) {  // sentinel: NaN coordinates never pass the distance test
                    sf[k] = make_float4(qn_, qn_, qn_, __uint_as_float(0xffffffffu));
                    sc[k].meta = 0xffffu;
                }
            }
            cp_async_wait_all();
            __syncthreads();
            // pairwise in-place transpose {x,y,z,cv} x 2 -> {x0,x1,y0,y1}, {z0,z1,cv0,cv1}
            for (u32 p = tid; 2 * p + 1 < tend; p += nthreads) {
                const float4 a = sf[2 * p], b = sf[2 * p + 1];
                sf[2 * p] = make_float4(a.x, b.x, a.y, b.y);
                sf[2 * p + 1] = make_float4(a.z, b.z, a.w, b.w);
                sc[2 * p].cv = __float_as_uint(a.w);
                sc[2 * p + 1].cv = __float_as_uint(b.w);
            }
        }
        __syncthreads();

        // ---- targets: groups of 32 consecutive agents, one warp each ---------------------------------------
        const u32 t0 = SS(4, z), t1 = SS(4, z + kc);
        const u32 w_end = t1;
        const u32 limit = (s_rs[kTileRows] + 24) & ~1u;  // last safe (even) start of an 8-candidate block
        for (u32 gb = t0 + 32u * (u32)warp; gb < w_end; gb += 32u * nwarps) {
            const u32 i_raw = gb + (u32)lane;
            const bool valid = i_raw < w_end;
            const u32 i = valid ? i_raw : w_end - 1;
            // integrator inputs: issued now, consumed after the pair loop
            AgentDyn<D> dyn;
            if constexpr (!FORCE_ONLY) dyn = load_agent_dyn<D>(P, i);
            const u32 sj_self = s_rs[4] + (i - s_g0[4]);
            float4 pf;
            {
                const float* fa = reinterpret_cast<const float*>(sf + (sj_self & ~1u)) + (sj_self & 1);
                const float* fb = reinterpret_cast<const float*>(sf + (sj_self | 1u)) + (sj_self & 1);
                pf = make_float4(fa[0], fa[2], fb[0], fb[2]);
            }
            TileTarget<D> t;
            t.x[0] = sc[sj_self].x;
            t.x[1] = sc[sj_self].y;
            if (D == 3) t.x[2] = sc[sj_self].z;
#pragma unroll
            for (int c = 0; c < D; ++c) t.acc[c] = 0.0;
            t.my_v = __float_as_uint(pf.w);
            const u32 mself = sc[sj_self].meta;
            t.sp = (u8)(mself & 0xffu);
            t.zw = mself >> 8;
            t.calc_flag = 0;
            t.subv_lo = 0;
            t.subv_len = 0;
            const int zt = (int)(t.my_v - rv_c);
            if (P.weight_mode == CRB_WEIGHT_CHILI && g.sm_nsub <= 1) {
                t.subv_len = (u32)g.nvox_local();  // one chili subdomain: everything is "local"
            } else if (P.weight_mode == CRB_WEIGHT_CHILI) {
                const u64 plane = (u64)n1 * n2;
                const u64 gp = ((u64)(X + g.p0) * n1 + Y) * n2 + zt;
                const u64 s = g.sub_of(gp);
                const u64 base = (u64)g.p0 * plane, nloc = g.nvox_local();
                u64 f = g.sub_first_voxel(s), l = g.sub_first_voxel(s + 1);
                f = f < base ? 0 : (f - base > nloc ? nloc : f - base);
                l = l < base ? 0 : (l - base > nloc ? nloc : l - base);
                t.subv_lo = (u32)f;
                t.subv_len = (u32)(l - f);
            }
            // filter coordinates; lanes without a target never pass
            const float qnan = __uint_as_float(0x7fc00000u);
            const float xf = valid ? pf.x : qnan, yf = pf.y, zf = pf.z;
            // every sub-group of kTileSub consecutive lanes streams its OWN candidate window (the
            // voxels zmin-1 .. zmax+1 of its targets): lanes of different sub-groups read different
            // shared-memory addresses in the same LDS.128 (still one wavefront), and the number of
            // (target, candidate) tests drops with the z-span of the sub-group instead of the warp's.
            const int zmin = __shfl_sync(0xffffffffu, zt, 0, kTileSub);
            const int zmax = __shfl_sync(0xffffffffu, zt, kTileSub - 1, kTileSub);
            const int zlo = max(zmin - 1, wa), zhi = min(zmax + 2, wb);

            // per-lane FIFO of in-range candidate indices (ascending): q[k][tid]
            // (32-bit shared-window address kept in a register: one predicated STS + one predicated
            // IADD per accepted candidate)
            unsigned short* const q0 = q + tid;
            const u32 qa0 = (u32)__cvta_generic_to_shared(q0);
            u32 qa = qa0;
            auto drain = [&]() {
                const u32 nq = (qa - qa0) / (2u * nthreads);
                const u32 nqmax = __reduce_max_sync(0xffffffffu, nq);
                for (u32 k = 0; k < nqmax; k += NE) {
                    u32 jj[NE];
                    bool hh[NE];
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        const u32 ent = q0[(k + e) * nthreads];
                        hh[e] = (k + e < nq) & (ent != sj_self);
                        jj[e] = hh[e] ? ent : sj_self;
                    }
                    tile_eval<D, INTER, NE>(t, own, ext_radius, sc, jj, hh);
                }
                qa = qa0;
            };
            const u32 qfull = qa0 + 2u * (kTileQueue - 8) * nthreads;
            const float2 xf2 = make_float2(-xf, -xf), yf2 = make_float2(-yf, -yf), zf2 = make_float2(-zf, -zf);

            for (int r = 0; r < kTileRows; ++r) {
                const u32 off = s_rs[r] - s_g0[r];
                const u32 lo = off + SS(r, zlo), hi = off + SS(r, zhi);
                const u32 lo2 = lo & ~1u;  // blocks start at an even candidate (pair layout); the extra
                                           // leading candidate is voxel zmin-2 or a sentinel: rejected
                                           // by the distance test or by the adjacency test of the drain
                const u32 n = hi - lo2;
                const u32 nmax = __reduce_max_sync(0xffffffffu, n);
                if (nmax == 0) continue;  // warp-uniform
                for (u32 k0 = 0; k0 < nmax; k0 += 8) {
                    if (__any_sync(0xffffffffu, qa > qfull)) drain();
                    // a block past the lane's range is disabled through the threshold; a block that
                    // starts inside may over-read <= 7 entries (later voxels of the row or sentinels)
                    const float cut = k0 < n ? T.cut2f : -1.0f;
                    const u32 b8 = min(lo2 + k0, limit);
                    float4 ca4[4], cb4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {  // all loads first: the queue stores below alias smem
                        ca4[u] = sf[b8 + 2 * u];
                        if (D == 3) cb4[u] = sf[b8 + 2 * u + 1];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float2 dx = __fadd2_rn(make_float2(ca4[u].x, ca4[u].y), xf2);
                        const float2 dy = __fadd2_rn(make_float2(ca4[u].z, ca4[u].w), yf2);
                        float2 d2 = __fmul2_rn(dx, dx);
                        d2 = __ffma2_rn(dy, dy, d2);
                        if (D == 3) {
                            const float2 dz = __fadd2_rn(make_float2(cb4[u].x, cb4[u].y), zf2);
                            d2 = __ffma2_rn(dz, dz, d2);
                        }
                        if (d2.x <= cut) {
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "h"((unsigned short)(b8 + 2 * u)) : "memory");
                            qa += 2u * nthreads;
                        }
                        if (d2.y <= cut) {
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "h"((unsigned short)(b8 + 2 * u + 1)) : "memory");
                            qa += 2u * nthreads;
                        }
                    }
                }
            }
            drain();

            if (valid) {
                if (t.calc_flag) latch_error(P.err, CRB_ERR_CALC, P.A.id[i]);
                if constexpr (FORCE_ONLY) {
                    add_domain_force<D>(P, t.x, t.acc);
#pragma unroll
                    for (int c = 0; c < D; ++c) P.force_out[(u64)c * P.cap + i] = t.acc[c];
                } else {
                    integrate_agent<D, true>(P, i, t.my_v, SS(4, zt), X, Y, zt, t.x, t.acc, dyn);
                }
            }
        }
        z += kc;
    }
#undef SS
}

}  // namespace crb
