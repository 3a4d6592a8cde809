// crb_force_tile3.cuh — 2 x 2-row tiled neighbour-force kernel whose distance filter runs on the tensor
// cores (uniform parameters), sm_100a.
//
// Same job, same tile geometry and same FP64 arithmetic as k_force_tile2 (crb_force_tile2.cuh; replaces
// SubDomainBox::update_mechanics_interaction_step_1/2/3, local_mechanics_update and apply_boundary,
// cellular_raza-core/src/backend/chili/update_mechanics.rs:153-238, 285-382, 387-406, 530-559).  What
// changes is the conservative pre-test "is candidate j within the cutoff of target i?", which was 36 % of
// the instructions of k_force_tile2 (nine FP32 operations, two compares and four predicated queue
// operations per pair of tests and lane) and half of its shared-memory traffic.
//
//   |a - b|^2 - cut^2  =  -2 a.b + |b|^2 + (|a|^2 - cut^2)
//
// is ONE m16n8k8 TF32 tensor-core instruction for 16 targets x 8 candidates:
//     A row (target)    = [-2ax, -2ay, -2az, 1, 1, 0, 0, 0]
//     B column (cand.)  = [  bx,   by,   bz, nh, nl, *, *, *]      nh + nl = |b|^2 split into two TF32
//     C (per row)       = |a|^2 - cut^2                            full FP32
// with coordinates relative to the centre of the staged window and rounded to TF32 ONCE at staging
// (targets and candidates share the rounded values, so the instruction computes the exact squared
// distance of the ROUNDED points up to FP32 accumulation).  cut^2 is inflated on the host by the TF32
// rounding of the coordinates (<= 2^-11 relative per axis), the FP32 rounding of the filter copy and an
// accumulation allowance (plan_tile3), so that no pair inside the true cutoff is ever rejected; a pair
// that is accepted needlessly contributes exactly zero (every potential of the model set is zero beyond
// its cutoff).  The result of the step therefore has the same bits as k_force_tile2 / k_force_tile /
// k_force_integrate (tested).
//
// A warp owns 32 target slots = two halves of 16 targets; a half lies in one target row and streams its
// own candidate window (voxels zmin-1 .. zmax+1 of its targets) through the 9 neighbour rows in groups of
// 32 candidates = 4 MMAs.  The sign bits of the 4 results per lane are gathered with one PRMT (byte
// sign replication) + one LOP3 per pair, OR-reduced over the quad by three shuffles, and the 32-bit
// acceptance mask of (target, 32 candidates) goes to a per-lane FIFO of (first candidate, mask) entries.
// The drain walks the set bits in ascending order — ascending candidate index of the voxel-sorted state,
// the summation order of every other force kernel — four candidates per iteration through tile_eval().
//
// Filter copy in shared memory: five FP32 arrays x[], y[], z[], nh[], nl[] (20 B per candidate) whose
// bases are offset by {0, 4, 16, 20, 0} words modulo 32, which makes the B-fragment loads of a step
// (lane (g, t) reads array t at candidates {0..3, 8..11} + step offset) conflict free.  No sentinels: a
// group that reaches past its window reads whatever follows (finite staged values or stale shared memory
// that only reaches columns whose bits are masked off).
#pragma once
#include "crb_force_tile2.cuh"

namespace crb {

constexpr int kT3Queue = 20;  // FIFO entries (groups of 32 candidates with >= 1 accepted) per lane (+ 1 terminator)

__host__ __device__ inline u32 tile3_c32(u32 cap) { return (cap + 32u + 31u) & ~31u; }

__host__ __device__ inline size_t tile3_smem_bytes(u32 cap, int kmax, int threads) {
    const size_t c32 = tile3_c32(cap);
    size_t b = 0;
    b += c32 * sizeof(TileCand);                                   // exact positions + meta + voxel
    b += (5 * c32 + 32) * sizeof(float);                           // x, y, z, nh, nl (bank-staggered)
    b += (size_t)kT2Rows * (kmax + 1) * sizeof(u32);               // start[] of the 16 rows
    b += (size_t)(kT3Queue + 1) * threads * (sizeof(u32) + sizeof(unsigned short));  // per-lane FIFOs
    return b + 64;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ float to_tf32(float x) {
    u32 r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// D = A (16 x 8, row) * B (8 x 8, col) + C, TF32 inputs, FP32 accumulate.  Fragment layout (PTX ISA, m16n8k8
// .tf32; g = lane / 4, t = lane % 4):  a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4);
// b0 (k = t, n = g)  b1 (k = t + 4, n = g);  c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1).
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const u32 (&a)[4], u32 b0, u32 b1,
                                                const float (&c)[4]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%10, %11, %12, %13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

// an opaque copy: the compiler cannot merge it with its source, so a fragment whose registers hold equal values
// (a2 == a3, c0 == c1) stays resident as four registers instead of being re-assembled by moves before every MMA
__device__ __forceinline__ u32 opaque(u32 x) {
    u32 y;
    asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ float opaque(float x) { return __uint_as_float(opaque(__float_as_uint(x))); }

// bytes {sign(x), sign(y), sign(x), sign(y)} (0x00 / 0xff): PRMT with sign replication of the top bytes
__device__ __forceinline__ u32 sign_bytes(float x, float y) {
    u32 w;
    asm("prmt.b32 %0, %1, %2, 0xFBFB;" : "=r"(w) : "r"(__float_as_uint(x)), "r"(__float_as_uint(y)));
    return w;
}

template <int D, int INTER, bool FORCE_ONLY, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_force_tile3(const KParams P, const TileLaunch T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int nthreads = NT;  // launch contract: blockDim.x == NT (128: 4 CTAs/SM, 256: 2 CTAs/SM)
    constexpr u32 FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = nthreads >> 5;
    constexpr int NE = 4;  // candidates evaluated per drain iteration and lane
    const Grid& g = P.g;
    const int n1 = g.ng[1], n2 = g.ng[2];

    // ---- carve shared memory -----------------------------------------------------------------------
    const u32 c32 = tile3_c32(T.cap);
    TileCand* sc = reinterpret_cast<TileCand*>(smem_raw);
    float* fbase = reinterpret_cast<float*>(sc + c32);  // 128-byte aligned (c32 * 32 B)
    float* const FX = fbase;
    float* const FY = fbase + c32 + 4;
    float* const FZ = fbase + 2 * c32 + 16;
    float* const FNH = fbase + 3 * c32 + 20;
    float* const FNL = fbase + 4 * c32 + 32;
    u32* sst = reinterpret_cast<u32*>(fbase + 5 * c32 + 32);  // [16][kmax + 1] copy of start[]
    u32* qm = sst + kT2Rows * (T.kmax + 1);                    // [kT3Queue + 1][NT] acceptance masks
    unsigned short* qb = reinterpret_cast<unsigned short*>(qm + (kT3Queue + 1) * nthreads);  // first candidate
    __shared__ u32 s_rowvox[kT2Rows];   // first voxel (local plain index) of every staged row
    __shared__ u32 s_g0[kT2Rows];       // first global agent index staged per row
    __shared__ u32 s_rs[kT2Rows + 1];   // staged index of the first candidate of every row
    __shared__ int s_valid[kT2Rows];
    __shared__ u32 s_tpre[5];           // first slot of every target row (multiples of 16)
    __shared__ u32 s_tn[4];             // targets of every target row in the window
    __shared__ u32 s_t0[4];             // global index of the first of them
    __shared__ int s_tlast;             // a target row that holds targets (for lanes beyond the slots)
    __shared__ int s_kc, s_za, s_zb;

    // ---- which tile: 2 x 2 rows, chunk of the fastest axis, count-balanced segment of the chunk ------
    const int ny2 = (n1 + 1) >> 1;
    const int seg = blockIdx.x % T.nseg;
    const int chunk = (blockIdx.x / T.nseg) % T.nchunks;
    const int tileid = blockIdx.x / (T.nseg * T.nchunks);
    const int Y = (tileid % ny2) * 2;
    const int X = T.x_begin + (tileid / ny2) * 2;
    const bool va1 = X + 1 < T.x_begin + T.x_count;  // second target plane exists
    const bool vb1 = Y + 1 < n1;                     // second target row exists
    const int ca = chunk * T.chunk_len;
    const int cb = min(ca + T.chunk_len, n2);
    {
        // CTAs beyond the segments this tile needs leave before anything is staged (broadcast loads)
        u32 R = 0;
#pragma unroll
        for (int tr = 0; tr < 4; ++tr) {
            const int a = tr >> 1, b = tr & 1;
            if ((a == 0 || va1) && (b == 0 || vb1)) {
                const u32 rv = ((u32)(X + a) * (u32)n1 + (u32)(Y + b)) * (u32)n2;
                R += __ldg(P.start + rv + cb) - __ldg(P.start + rv + ca);
            }
        }
        if ((u32)seg * T.tseg >= R) return;  // seg >= ceil(R / tseg); uniform
    }
    if (tid < kT2Rows) {
        const int p = tid >> 2, qq = tid & 3;
        const int XX = X - 1 + p, YY = Y - 1 + qq;
        // a staged row must lie in the slab and touch a target row of this tile
        const bool ok = XX >= 0 && XX < g.L0 && YY >= 0 && YY < n1 && (p <= 2 || va1) && (qq <= 2 || vb1);
        s_valid[tid] = ok ? 1 : 0;
        s_rowvox[tid] = ok ? ((u32)XX * (u32)n1 + (u32)YY) * (u32)n2 : 0u;
    }
    if (tid == 0) {
        s_za = cb;
        s_zb = cb;
    }
    __syncthreads();
    // ---- ONE round of global loads: start[] of the 16 rows over the chunk +- 1 voxel -> shared ------
    const int zb0 = max(ca - 1, 0), ztop = min(cb + 1, n2);
    const int sstride = T.kmax + 1;  // >= ztop - zb0 + 1
    {
        const int ns = ztop - zb0 + 1;
        for (int r = warp; r < kT2Rows; r += nwarps) {
            const bool ok = s_valid[r] != 0;
            const u32 rv = s_rowvox[r];
            for (int k = lane; k < ns; k += 32) sst[r * sstride + k] = ok ? __ldg(P.start + rv + (u32)(zb0 + k)) : 0u;
        }
    }
    __syncthreads();
#define SS(r, zz) sst[(r) * sstride + ((zz) - zb0)]
    // staged row of target row tr = (a, b): (a + 1) * 4 + (b + 1) = 5, 6, 9, 10
    const int c0 = 5, c1 = 6, c2 = 9, c3 = 10;
    const bool tv1 = vb1, tv2 = va1, tv3 = va1 && vb1;
    auto tcount = [&](int zlo_, int zhi_) -> u32 {  // targets of the tile in voxels [zlo_, zhi_)
        u32 c = SS(c0, zhi_) - SS(c0, zlo_);
        if (tv1) c += SS(c1, zhi_) - SS(c1, zlo_);
        if (tv2) c += SS(c2, zhi_) - SS(c2, zlo_);
        if (tv3) c += SS(c3, zhi_) - SS(c3, zlo_);
        return c;
    };
    {
        const u32 R = tcount(ca, cb);
        const u32 nsr = min((R + T.tseg - 1) / T.tseg, (u32)T.nseg);
        if ((u32)seg >= nsr) return;  // uniform: no barrier is pending
        const u32 tsr = (R + nsr - 1) / nsr;
        const u32 lo_cnt = (u32)seg * tsr;
        const u32 hi_cnt = (u32)seg == nsr - 1 ? 0xffffffffu : lo_cnt + tsr;
        for (int zz = ca + tid; zz < cb; zz += nthreads) {
            const u32 c = tcount(ca, zz);
            if (c >= lo_cnt) atomicMin(&s_za, zz);
            if (c >= hi_cnt) atomicMin(&s_zb, zz);
        }
    }
    __syncthreads();
    const int za = s_za, zb = s_zb;
    if (za >= zb) return;
    if (tcount(za, zb) == 0) return;  // no targets

    // lane (g, t) of the MMA fragments; its own target is row (t & 1 ? g + 8 : g) of half (t >> 1)
    const int fg = lane >> 2, ft = lane & 3;
    const u32 slot_in_warp = 16u * (u32)(ft >> 1) + 8u * (u32)(ft & 1) + (u32)fg;
    // B fragment: array t of the filter copy, candidates {0..3, 8..11} + step offset of a group
    const u32* const Fb0 = reinterpret_cast<const u32*>(ft == 0 ? FX : ft == 1 ? FY : ft == 2 ? FZ : FNH) +
                           (fg >> 1) + 8 * (fg & 1);
    const u32* const Fb1 = reinterpret_cast<const u32*>(FNL) + (fg >> 1) + 8 * (fg & 1);
    // result bits of step s: c0 -> candidate {0, 4, 16, 20}[s] + t (bytes 0 / 2), c1 -> that + 8 (bytes 1 / 3)
    const u32 kbit = (1u << ft) | (1u << (8 + ft));

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
                for (int r = 0; r < kT2Rows; ++r) total += SS(r, wb) - SS(r, wa);
                if (total <= T.cap) {
                    atomicMax(&s_kc, kk);
                    found = 1;
                }
            }
        }
        __syncthreads();
        const int kc = s_kc;
        if (kc == 0) {
            // the neighbourhood of voxel column z alone exceeds the staging capacity: per-agent path
#pragma unroll
            for (int tr = 0; tr < 4; ++tr) {
                const int a = tr >> 1, b = tr & 1;
                if (!((a == 0 || va1) && (b == 0 || vb1))) continue;
                const int sr = (a + 1) * 4 + (b + 1);
                const u32 v = s_rowvox[sr] + (u32)z;
                const u32 t0 = SS(sr, z), t1 = SS(sr, z + 1);
                for (u32 i = t0 + tid; i < t1; i += nthreads) {
                    double xi[D], acc[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        xi[c] = P.A.pos[c][i];
                        acc[c] = 0.0;
                    }
                    force_from_global<D, INTER>(P, i, v, X + a, Y + b, z, xi, acc);
                    if constexpr (FORCE_ONLY) {
                        add_domain_force<D>(P, xi, acc);
#pragma unroll
                        for (int c = 0; c < D; ++c) P.force_out[(u64)c * P.cap + i] = acc[c];
                    } else {
                        integrate_agent<D, true>(P, i, v, t0, X + a, Y + b, z, xi, acc, load_agent_dyn<D>(P, i));
                    }
                }
            }
            z += 1;
            continue;
        }
        const int wa = max(z - 1, 0), wb = min(z + kc + 1, n2);

        // ---- row ranges (warp 0) and target slots (warp 1) ---------------------------------------------
        if (warp == 0) {
            u32 a = 0, n = 0;
            if (lane < kT2Rows) {
                a = SS(lane, wa);
                n = SS(lane, wb) - a;
            }
            const u32 np = (lane < kT2Rows && s_valid[lane & 15]) ? n : 0u;
            u32 incl = np;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const u32 y = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane < kT2Rows) {
                s_g0[lane] = a;
                s_rs[lane] = incl - np;
            }
            if (lane == kT2Rows - 1) s_rs[kT2Rows] = incl;
        } else if (warp == 1) {
            u32 n = 0, t0 = 0;
            if (lane < 4) {
                const int a = lane >> 1, b = lane & 1;
                const int sr = (a + 1) * 4 + (b + 1);
                if ((a == 0 || va1) && (b == 0 || vb1)) {
                    t0 = SS(sr, z);
                    n = SS(sr, z + kc) - t0;
                }
            }
            const u32 m = (n + 15u) & ~15u;  // a half-warp of 16 slots lies in ONE target row
            u32 incl = m;
#pragma unroll
            for (int o = 1; o < 4; o <<= 1) {
                const u32 y = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane < 4) {
                s_tpre[lane] = incl - m;
                s_tn[lane] = n;
                s_t0[lane] = t0;
            }
            if (lane == 3) s_tpre[4] = incl;
            const u32 has = __ballot_sync(FULL, n > 0) & 0xfu;
            if (lane == 0) s_tlast = has ? 31 - __clz((int)has) : 0;
        }
        __syncthreads();

        // ---- stage the candidates: every warp takes whole rows -------------------------------------------
        // origin of the filter coordinates (units of posf = position - domain min): centre of the 4 x 4 rows
        // and of the window along the fastest axis
        float o0, o1, o2;
        float cut2w;  // conservative cutoff^2 of THIS window (plan_tile bounds the largest window; most are shorter)
        {
            // |c0|, |c1| <= 2 voxels, |c2| <= half the window (+ 1 % for agents on a face); TF32 rounding moves a
            // coordinate by <= 2^-11 of its magnitude; T.cut2f is the FP32 filter's cutoff^2 = (cutoff + FP32
            // allowance)^2, to which the TF32 allowance and the accumulation allowance 2^-17 r^2_max are added
            const float m0 = 2.02f * (float)g.dx[0];
            const float m1 = D == 3 ? 2.02f * (float)g.dx[1] : 0.0f;
            const float m2 = (0.5f * (float)(wb - wa) + 0.02f) * (float)g.dx[D - 1];
            const float r2max = m0 * m0 + m1 * m1 + m2 * m2;
            const float delta = sqrtf(r2max) * 4.8828125e-4f * 1.0001f;  // 2^-11
            const float c = sqrtf(T.cut2f) * 1.000001f + 2.0f * delta;
            cut2w = fminf(c * c * 1.000001f + r2max * 7.62939453125e-6f, T.cut2m);  // 2^-17
            const float zc = 0.5f * (float)(wa + wb);
            o0 = (float)((double)(X + g.p0 + 1) * g.dx[0]);
            if (D == 3) {
                o1 = (float)((double)(Y + 1) * g.dx[1]);
                o2 = zc * (float)g.dx[2];
            } else {
                o1 = zc * (float)g.dx[1];
                o2 = 0.0f;
            }
        }
        {
            for (int r = warp; r < kT2Rows; r += nwarps) {
                if (!s_valid[r]) continue;
                const u32 rs = s_rs[r], g0 = s_g0[r];
                const u32 n = s_rs[r + 1] - rs;
                for (u32 o = lane; o < n; o += 32) {
                    const u32 k = rs + o, gi = g0 + o;
                    const float4 a = __ldg(P.A.posf + gi);
                    cp_async_8(&sc[k].x, P.A.pos[0] + gi);
                    cp_async_8(&sc[k].y, P.A.pos[1] + gi);
                    if (D == 3) cp_async_8(&sc[k].z, P.A.pos[2] + gi);
                    cp_async_4(&sc[k].meta, P.A.meta + gi);
                    const float bx = to_tf32(a.x - o0), by = to_tf32(a.y - o1), bz = to_tf32(a.z - o2);
                    const float nb = fmaf(bz, bz, fmaf(by, by, bx * bx));
                    const float nh = to_tf32(nb);
                    FX[k] = bx;
                    FY[k] = by;
                    FZ[k] = bz;
                    FNH[k] = nh;
                    FNL[k] = to_tf32(nb - nh);
                    sc[k].cv = __float_as_uint(a.w);
                }
            }
            cp_async_wait_all();
        }
        __syncthreads();

        // ---- targets: 32 slots per warp = two halves of 16 -------------------------------------------------
        const u32 S = s_tpre[4];
        const u32 total = s_rs[kT2Rows];
        for (u32 gb = 32u * (u32)warp; gb < S; gb += 32u * nwarps) {
            const u32 sl = gb + slot_in_warp;
            int tr = (sl >= s_tpre[1]) + (sl >= s_tpre[2]) + (sl >= s_tpre[3]);
            if (sl >= S) tr = s_tlast;
            const u32 ntr = s_tn[tr];  // > 0: a row without targets has no slots
            const u32 o_raw = sl - s_tpre[tr];
            const bool valid = sl < S && o_raw < ntr;
            const u32 i = s_t0[tr] + min(o_raw, ntr - 1u);
            const int ta = tr >> 1, tb = tr & 1;
            const int Xt = X + ta, Yt = Y + tb;
            const int src = ta * 4 + tb + 5;  // staged row of the target's own row
            const u32 sj_self = s_rs[src] + (i - s_g0[src]);
            // the filter phase keeps only a handful of per-target values in registers (the tensor-core loop wants
            // its fragments resident): the exact position, species and voxel of the target are re-read from its
            // staged record by every drain and by the epilogue, the integrator inputs are only prefetched here
            if constexpr (!FORCE_ONLY) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    if (P.mech != CRB_MECH_BROWNIAN) prefetch_l1(&P.A.vel[c][i]);
                    prefetch_l1(&P.A.h1x[c][i]);
                    prefetch_l1(&P.A.h1v[c][i]);
                }
            }
            double facc[D];
#pragma unroll
            for (int c = 0; c < D; ++c) facc[c] = 0.0;
            int calc_flag = 0;
            u32 subv_lo = 0, subv_len = 0;
            const int zt = (int)(sc[sj_self].cv - s_rowvox[src]);
            if (P.weight_mode == CRB_WEIGHT_CHILI && g.sm_nsub <= 1) {
                subv_len = (u32)g.nvox_local();  // one chili subdomain: everything is "local"
            } else if (P.weight_mode == CRB_WEIGHT_CHILI) {
                const u64 plane = (u64)n1 * n2;
                const u64 gp = ((u64)(Xt + g.p0) * n1 + Yt) * n2 + zt;
                const u64 s = g.sub_of(gp);
                const u64 base = (u64)g.p0 * plane, nloc = g.nvox_local();
                u64 f = g.sub_first_voxel(s), l = g.sub_first_voxel(s + 1);
                f = f < base ? 0 : (f - base > nloc ? nloc : f - base);
                l = l < base ? 0 : (l - base > nloc ? nloc : l - base);
                subv_lo = (u32)f;
                subv_len = (u32)(l - f);
            }
            // ---- the two halves: target row and candidate window (voxels zmin-1 .. zmax+1) of each --------
            // half h = slots gb + 16 h .. + 15: row 0 sits in lane 2h, row 15 in lane 29 + 2h
            const int trA = __shfl_sync(FULL, tr, 0), trB = __shfl_sync(FULL, tr, 2);
            const int zminA = __shfl_sync(FULL, zt, 0), zmaxA = __shfl_sync(FULL, zt, 29);
            const int zminB = __shfl_sync(FULL, zt, 2), zmaxB = __shfl_sync(FULL, zt, 31);
            const bool hvB = gb + 16u < S;  // half A always holds a target (gb < S)
            const int zloA = max(zminA - 1, wa), zhiA = min(zmaxA + 2, wb);
            const int zloB = max(zminB - 1, wa), zhiB = min(zmaxB + 2, wb);
            const int sr0A = (trA >> 1) * 4 + (trA & 1), sr0B = (trB >> 1) * 4 + (trB & 1);
            // filter coordinates of the own target (the staged, rounded values) and |a|^2 - cut^2
            const float ax = FX[sj_self], ay = FY[sj_self], az = FZ[sj_self];
            const float sown = fmaf(az, az, fmaf(ay, ay, ax * ax)) - cut2w;
            const float m2x = -2.0f * ax, m2y = -2.0f * ay, m2z = -2.0f * az;

            u32* const qm0 = qm + tid;
            unsigned short* const qb0 = qb + tid;
            // candidate window of every (half, neighbour row): lane 9 h + r holds its first staged index and its
            // length (computed once per pass; the row loop only shuffles them out)
            u32 rw_lo = 0;
            int rw_n = 0;
            if (lane < 18) {
                const int h = lane >= 9, rr = lane - 9 * h;
                const int sr = (h ? sr0B : sr0A) + (rr / 3) * 4 + (rr % 3);
                const int zlo = h ? zloB : zloA, zhi = h ? zhiB : zhiA;
                const u32 first = SS(sr, zlo);
                rw_lo = s_rs[sr] - s_g0[sr] + first;
                rw_n = (h && !hvB) ? 0 : (int)(SS(sr, zhi) - first);
            }
            const bool odd = (ft & 1) != 0, hi = (ft & 2) != 0;
            int r = 0;
            u32 gi = 0;
            do {
                // ---- filter phase: rows r.. until done or a FIFO is full ------------------------------------
                u32 qn = 0;
                {
                    // A fragments of both halves (rebuilt after every drain so that they do not live across it)
                    u32 aA[4], aB[4];
                    float cA[4], cB[4];
                    {
                        const int q0 = lane & ~3;
                        float v[4][3];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[j][0] = __shfl_sync(FULL, m2x, q0 + j);
                            v[j][1] = __shfl_sync(FULL, m2y, q0 + j);
                            v[j][2] = __shfl_sync(FULL, m2z, q0 + j);
                        }
                        auto pick = [&](int j) -> u32 {
                            const float x = ft == 0 ? v[j][0] : ft == 1 ? v[j][1] : ft == 2 ? v[j][2] : 1.0f;
                            return __float_as_uint(x);
                        };
                        const u32 one0 = ft == 0 ? 0x3f800000u : 0u;  // column 4 carries nl, columns 5..7 are unused
                        aA[0] = pick(0);
                        aA[1] = pick(1);
                        aA[2] = one0;
                        aA[3] = opaque(one0);
                        aB[0] = pick(2);
                        aB[1] = pick(3);
                        aB[2] = opaque(one0);
                        aB[3] = opaque(one0);
                        cA[0] = __shfl_sync(FULL, sown, q0);
                        cA[1] = opaque(cA[0]);
                        cA[2] = __shfl_sync(FULL, sown, q0 + 1);
                        cA[3] = opaque(cA[2]);
                        cB[0] = __shfl_sync(FULL, sown, q0 + 2);
                        cB[1] = opaque(cB[0]);
                        cB[2] = __shfl_sync(FULL, sown, q0 + 3);
                        cB[3] = opaque(cB[2]);
                    }
                    const u32 ks0 = kbit, ks1 = opaque(kbit << 4);
                    bool full = false;
#pragma unroll 1
                    while (r < 9 && !full) {
                        const u32 loA = __shfl_sync(FULL, rw_lo, r), loB = __shfl_sync(FULL, rw_lo, 9 + r);
                        const int nA = __shfl_sync(FULL, rw_n, r), nB = __shfl_sync(FULL, rw_n, 9 + r);
                        const u32 ng = (u32)(max(nA, nB) + 31) >> 5;
                        // room for the whole row in every FIFO?  (a row longer than the FIFO is checked per group)
                        const bool each = ng > (u32)kT3Queue;
                        if (!each && gi == 0 && __any_sync(FULL, qn + ng > (u32)kT3Queue)) {
                            full = true;
                            break;
                        }
                        // the target's own staged record sits in row 4 of its 9 neighbour rows: its bit is cleared here
                        // so that the drain need not compare every candidate with the target
                        const u32 self_off = r == 4 ? sj_self - (hi ? loB : loA) : 0xffffffffu;
#pragma unroll 1
                        for (; gi < ng; ++gi) {
                            const int remA = nA - (int)(32u * gi), remB = nB - (int)(32u * gi);
                            const u32 baseA = min(loA + 32u * gi, total), baseB = min(loB + 32u * gi, total);
                            u32 accA0 = 0, accA1 = 0, accB0 = 0, accB1 = 0;
                            const int nit = max(remA, remB) > 16 ? 2 : 1;
                            const u32* pA0 = Fb0 + baseA;
                            const u32* pA1 = Fb1 + baseA;
                            const u32* pB0 = Fb0 + baseB;
                            const u32* pB1 = Fb1 + baseB;
#pragma unroll 1
                            for (int it = 0; it < nit; ++it) {  // candidates {0..3, 8..11}, {4..7, 12..15} (+ 16 it)
                                const int sh = 16 * it;
                                const u32 k0 = ks0 << sh, k1 = ks1 << sh;
                                float d[4], e4[4];
                                if (remA > 0) {
                                    mma_tf32_16x8x8(d, aA, pA0[0], pA1[0], cA);
                                    mma_tf32_16x8x8(e4, aA, pA0[4], pA1[4], cA);
                                    accA0 |= (sign_bytes(d[0], d[1]) & k0) | (sign_bytes(e4[0], e4[1]) & k1);
                                    accA1 |= (sign_bytes(d[2], d[3]) & k0) | (sign_bytes(e4[2], e4[3]) & k1);
                                }
                                if (remB > 0) {
                                    mma_tf32_16x8x8(d, aB, pB0[0], pB1[0], cB);
                                    mma_tf32_16x8x8(e4, aB, pB0[4], pB1[4], cB);
                                    accB0 |= (sign_bytes(d[0], d[1]) & k0) | (sign_bytes(e4[0], e4[1]) & k1);
                                    accB1 |= (sign_bytes(d[2], d[3]) & k0) | (sign_bytes(e4[2], e4[3]) & k1);
                                }
                                pA0 += 16;
                                pA1 += 16;
                                pB0 += 16;
                                pB1 += 16;
                            }
                            // OR over the quad; lane t keeps {A row g, A row g + 8, B row g, B row g + 8}[t]
                            u32 keepA = odd ? accA1 : accA0, keepB = odd ? accB1 : accB0;
                            keepA |= __shfl_xor_sync(FULL, odd ? accA0 : accA1, 1);
                            keepB |= __shfl_xor_sync(FULL, odd ? accB0 : accB1, 1);
                            u32 m = hi ? keepB : keepA;
                            m |= __shfl_xor_sync(FULL, hi ? keepA : keepB, 2);
                            const int rem = hi ? remB : remA;
                            u32 vm = rem >= 32 ? FULL : ~(FULL << max(rem, 0));
                            const u32 so = self_off - 32u * gi;
                            if (so < 32u) vm &= ~(1u << so);
                            m = valid ? (m & vm) : 0u;
                            if (m) {
                                qm0[qn * nthreads] = m;
                                qb0[qn * nthreads] = (unsigned short)(hi ? baseB : baseA);
                                ++qn;
                            }
                            if (each && __any_sync(FULL, qn >= (u32)kT3Queue)) {
                                full = true;
                                ++gi;
                                break;
                            }
                        }
                        if (!full) {
                            ++r;
                            gi = 0;
                        }
                    }
                }
                // ---- drain (FP64, exact): set bits in ascending candidate order, NE per iteration -------------
                {
                    // (the potential's constants are derived here, not at kernel entry: twelve registers that the
                    // filter phase needs for its fragments)
                    Pot<INTER> own;
                    own.init(P.u.ip);
                    const double ext_radius = P.u.ip[0];
                    TileTarget<D> t;
                    {
                        const TileCand self = sc[sj_self];
                        t.x[0] = self.x;
                        t.x[1] = self.y;
                        if (D == 3) t.x[2] = self.z;
                        t.my_v = self.cv;
                        t.sp = (u8)(self.meta & 0xffu);
                        t.zw = self.meta >> 8;
                    }
#pragma unroll
                    for (int c = 0; c < D; ++c) t.acc[c] = facc[c];
                    t.calc_flag = 0;
                    t.subv_lo = subv_lo;
                    t.subv_len = subv_len;
                    // terminator entry: the walk below loads the next entry whenever a mask runs empty
                    qm0[qn * nthreads] = 0u;
                    const u32* pm = qm0;
                    const unsigned short* pb = qb0;
                    u32 m = *pm;
                    u32 base = *pb;
                    while (__any_sync(FULL, m != 0u)) {
                        u32 jj[NE];
                        bool hh[NE];
#pragma unroll
                        for (int x = 0; x < NE; ++x) {
                            hh[x] = m != 0u;
                            const u32 j = base + (u32)__clz((int)__brev(m));  // lowest set bit (32 when m == 0)
                            jj[x] = hh[x] ? j : sj_self;
                            m &= m - 1u;
                            if (hh[x] && m == 0u) {  // entry exhausted: the next one (the terminator is 0)
                                pm += nthreads;
                                pb += nthreads;
                                m = *pm;
                                base = *pb;
                            }
                        }
                        tile_eval<D, INTER, NE>(t, own, ext_radius, sc, jj, hh);
                    }
#pragma unroll
                    for (int c = 0; c < D; ++c) facc[c] = t.acc[c];
                    calc_flag |= t.calc_flag;
                }
            } while (r < 9);

            if (valid) {
                if (calc_flag) latch_error(P.err, CRB_ERR_CALC, P.A.id[i]);
                const TileCand self = sc[sj_self];
                double xs[D];
                xs[0] = self.x;
                xs[1] = self.y;
                if (D == 3) xs[2] = self.z;
                if constexpr (FORCE_ONLY) {
                    add_domain_force<D>(P, xs, facc);
#pragma unroll
                    for (int c = 0; c < D; ++c) P.force_out[(u64)c * P.cap + i] = facc[c];
                } else {
                    integrate_agent<D, true>(P, i, self.cv, SS(src, zt), Xt, Yt, zt, xs, facc, load_agent_dyn<D>(P, i));
                }
            }
        }
        z += kc;
    }
#undef SS
}

}  // namespace crb
