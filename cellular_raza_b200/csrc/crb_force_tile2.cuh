// crb_force_tile2.cuh — 2 x 2-row tiled neighbour-force kernel (uniform parameters), sm_100a.
//
// Same job and same arithmetic as k_force_tile (crb_force_tile.cuh; replaces
// SubDomainBox::update_mechanics_interaction_step_1/2/3, local_mechanics_update and apply_boundary,
// cellular_raza-core/src/backend/chili/update_mechanics.rs:153-238, 285-382, 387-406, 530-559) with a
// tile that shares its staged candidates between FOUR target rows:
//
//   one CTA owns the rows (X, Y), (X, Y+1), (X+1, Y), (X+1, Y+1) of the voxel grid over a segment
//   [za, zb) of the fastest axis.  Their 3^D neighbourhoods are the 4 x 4 rows (X-1..X+2, Y-1..Y+2)
//   over [za-1, zb], i.e. 16 CONTIGUOUS index ranges of the voxel-sorted state, staged once.  Per
//   target 16/4 = 4 rows' worth of candidates are staged instead of 9 (k_force_tile), so the same
//   shared memory feeds ~1.6 x the targets, every warp of the CTA is full and four CTAs fit one SM.
//
// Targets are laid out in slots: the targets of every target row padded to a multiple of 4, rows
// concatenated; a warp takes 32 consecutive slots, so each sub-group of 4 lanes lies in ONE row and
// streams one candidate window per neighbour row (as in k_force_tile).  The 9 neighbour rows of a
// lane are visited in ascending (x, y) order and every row in ascending z, i.e. ascending index of
// the sorted state: the per-target summation order — and therefore every bit of the result — is the
// one of k_force_tile and of the per-agent kernel.
//
// Staging: the exact data of a candidate is two 16-byte records {x, y} and {z, meta, voxel} in two arrays
// filled by cp.async (16-byte records spread the drain's random reads over all 32 banks; one 32-byte
// record per candidate used only half of them: 45 % of the shared-memory wavefronts of the round-1 kernel
// were bank conflicts); the FP32 filter copy is loaded as pairs and stored transposed ({x0,x1,y0,y1},
// {z0,z1}) so that one LDS.128 + one LDS.64 feed the packed f32x2 distance test of two candidates.  Every
// row starts at an even staged index and is followed by >= 8 sentinels (NaN: never pass the test).
//
// Round 2 (ncu source page, profiles/r2_ncu_k_force_tile2_8M.txt): 30 of the 96 instructions of the filter
// block loop re-derived the shared-memory carve-up (T.cap from the constant bank -> four base addresses)
// in EVERY iteration — the compiler rematerialises what it considers cheap instead of holding three
// registers — and 211 instructions of integer division ran in every warp of every CTA before the
// early-exit test.  The bases are now opaque 32-bit shared addresses (asm volatile: nothing to
// rematerialise), the hot loads are ld.shared with immediate offsets, the grid is three-dimensional
// (segment x chunk, tile row, tile plane) and the integrator inputs are prefetched instead of parked in 18
// registers across the pair loop.
#pragma once
#include "crb_force_tile.cuh"

namespace crb {

constexpr int kT2Rows = 16;  // staged rows (4 x 4)
constexpr int kT2Pad = 8;    // sentinels after every staged row

// second half of the exact record of a staged candidate ({x, y} is a plain double2)
struct __align__(16) TileCandZ {
    double z;
    u32 meta;  // species | (z voxel & 0xff) << 8
    u32 cv;    // local voxel index
};

__host__ __device__ inline size_t tile2_smem_bytes(u32 cap, int kmax, int threads) {
    const size_t c = cap + 34;  // cap includes the rows' sentinels; +34 filter over-read (even)
    size_t b = 0;
    b += c * (sizeof(double2) + sizeof(TileCandZ));                    // exact positions + meta + voxel
    b += (c / 2) * sizeof(float4);                                     // {x0, x1, y0, y1} per pair
    b += (c / 2) * sizeof(float2);                                     // {z0, z1} per pair
    b += (size_t)kT2Rows * (kmax + 1) * sizeof(u32);                   // start[] of the 16 rows
    b += (size_t)(kTileQueue + 4) * threads * sizeof(unsigned short);  // per-lane FIFOs
    return b + 64;
}

// 32-bit shared-window address that the compiler cannot re-derive (see the header: it would otherwise recompute the
// carve-up of the dynamic shared memory inside the hot loops)
__device__ __forceinline__ u32 smem_addr_opaque(const void* p) {
    u32 a = (u32)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ float4 lds_f4(u32 a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds_f2(u32 a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ u32 lds_u32(u32 a) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void prefetch_l1_t2(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int D, int INTER, bool FORCE_ONLY, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_force_tile2(const KParams P, const TileLaunch T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int nthreads = NT;  // launch contract: blockDim.x == NT (128: 4 CTAs/SM, 256: 2 CTAs/SM)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = nthreads >> 5;
    // candidates evaluated per drain iteration and lane (interleaved FP64 chains).  Same-box sweep, round 2: 2 / 3 / 4
    // are within 0.2 % of each other for the species potential; Morse (a longer chain per pair, 128 registers) is
    // 1.8 % faster with 3
    constexpr int NE = INTER == CRB_INT_MORSE ? 3 : 4;
    const Grid& g = P.g;
    const int n1 = g.ng[1], n2 = g.ng[2];

    // ---- which tile: grid = (segment x chunk, tile row, tile plane); CTAs beyond the segments a tile needs
    // leave before anything else happens (broadcast loads) -------------------------------------------------
    const int chunk = T.nseg > 1 ? (int)__umulhi(blockIdx.x, T.inv_nseg) : (int)blockIdx.x;  // blockIdx.x / nseg (exact: < 2^16)
    const int seg = (int)blockIdx.x - chunk * T.nseg;
    const int Y = (int)blockIdx.y * 2;
    const int X = T.x_begin + (int)blockIdx.z * 2;
    const bool va1 = X + 1 < T.x_begin + T.x_count;  // second target plane exists
    const bool vb1 = Y + 1 < n1;                     // second target row exists
    const int ca = chunk * T.chunk_len;
    const int cb = min(ca + T.chunk_len, n2);
    {
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

    // ---- carve shared memory -----------------------------------------------------------------------
    const u32 capp = T.cap + 34;  // even
    double2* sxyd = reinterpret_cast<double2*>(smem_raw);                 // [capp] exact {x, y}
    TileCandZ* szm = reinterpret_cast<TileCandZ*>(sxyd + capp);           // [capp] exact {z, meta, voxel}
    float4* sxy = reinterpret_cast<float4*>(szm + capp);                 // [capp / 2]
    float2* szz = reinterpret_cast<float2*>(sxy + capp / 2);             // [capp / 2]
    u32* sst = reinterpret_cast<u32*>(szz + capp / 2);                   // [16][kmax + 1] copy of start[]
    unsigned short* q = reinterpret_cast<unsigned short*>(sst + kT2Rows * (T.kmax + 1));
    // the addresses of the hot loops, held in registers
    const u32 a_sxy = smem_addr_opaque(sxy), a_szz = smem_addr_opaque(szz), a_sst = smem_addr_opaque(sst);
    const u32 qa0 = smem_addr_opaque(q + tid);
    __shared__ u32 s_rowvox[kT2Rows];   // first voxel (local plain index) of every staged row
    __shared__ u32 s_g0[kT2Rows];       // first global agent index staged per row
    __shared__ u32 s_rs[kT2Rows + 1];   // staged index of the first candidate of every row (even)
    __shared__ u32 s_rn[kT2Rows];       // candidates staged per row (without sentinels)
    __shared__ u32 s_off[kT2Rows];      // s_rs - s_g0: staged index = global index + s_off (wraps; even when unstaged)
    __shared__ int s_valid[kT2Rows];
    __shared__ u32 s_tpre[5];           // first slot of every target row (multiples of 4)
    __shared__ u32 s_tn[4];             // targets of every target row in the window
    __shared__ u32 s_t0[4];             // global index of the first of them
    __shared__ int s_tlast;             // a target row that holds targets (for lanes beyond the slots)
    __shared__ int s_kc, s_za, s_zb;

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
    // staged row of target row tr = (a, b): (a + 1) * 4 + (b + 1) = 5, 6, 9, 10; a row that is not a
    // target row of this tile counts zero targets
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
        // segment k owns the voxel columns whose first target has in-chunk rank in [k, k+1) * tsr
        // (ceil(R / tseg) segments of equal size): every CTA gets <= ~tseg targets whatever the density
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
                for (int r = 0; r < kT2Rows; ++r) total += ((SS(r, wb) - SS(r, wa)) + 1u) & ~1u;
                if (total + kT2Rows * kT2Pad <= T.cap) {
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
            const u32 np = (lane < kT2Rows && s_valid[lane & 15]) ? ((n + 1u) & ~1u) + kT2Pad : 0u;  // even
            u32 incl = np;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane < kT2Rows) {
                s_g0[lane] = a;
                s_rs[lane] = incl - np;
                s_rn[lane] = n;
                s_off[lane] = incl - np - a;
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
            const u32 m = (n + 3u) & ~3u;
            u32 incl = m;
#pragma unroll
            for (int o = 1; o < 4; o <<= 1) {
                const u32 y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane < 4) {
                s_tpre[lane] = incl - m;
                s_tn[lane] = n;
                s_t0[lane] = t0;
            }
            if (lane == 3) s_tpre[4] = incl;
            const u32 has = __ballot_sync(0xffffffffu, n > 0) & 0xfu;
            if (lane == 0) s_tlast = has ? 31 - __clz((int)has) : 0;
        }
        __syncthreads();

        // ---- stage the candidates: every warp takes whole rows -------------------------------------------
        {
            const float qn_ = __uint_as_float(0x7fc00000u);
            const float4 sent = make_float4(qn_, qn_, qn_, __uint_as_float(0xffffffffu));
            for (int r = warp; r < kT2Rows; r += nwarps) {
                if (!s_valid[r]) continue;
                const u32 n = s_rn[r], rs = s_rs[r], g0 = s_g0[r];
                const u32 npair = (((n + 1u) & ~1u) + kT2Pad) >> 1;
                for (u32 pi = lane; pi < npair; pi += 32) {
                    const u32 o = 2 * pi, k = rs + o;
                    float4 a = sent, b = sent;
                    if (o < n) {
                        const u32 gi = g0 + o;
                        a = __ldg(P.A.posf + gi);
                        cp_async_8(&sxyd[k].x, P.A.pos[0] + gi);
                        cp_async_8(&sxyd[k].y, P.A.pos[1] + gi);
                        if (D == 3) cp_async_8(&szm[k].z, P.A.pos[2] + gi);
                        cp_async_4(&szm[k].meta, P.A.meta + gi);
                    } else {
                        szm[k].meta = 0xffffu;
                    }
                    if (o + 1 < n) {
                        const u32 gi = g0 + o + 1;
                        b = __ldg(P.A.posf + gi);
                        cp_async_8(&sxyd[k + 1].x, P.A.pos[0] + gi);
                        cp_async_8(&sxyd[k + 1].y, P.A.pos[1] + gi);
                        if (D == 3) cp_async_8(&szm[k + 1].z, P.A.pos[2] + gi);
                        cp_async_4(&szm[k + 1].meta, P.A.meta + gi);
                    } else {
                        szm[k + 1].meta = 0xffffu;
                    }
                    sxy[k >> 1] = make_float4(a.x, b.x, a.y, b.y);
                    szz[k >> 1] = make_float2(a.z, b.z);
                    szm[k].cv = __float_as_uint(a.w);
                    szm[k + 1].cv = __float_as_uint(b.w);
                }
            }
            // over-read region behind the last row
            const u32 total = s_rs[kT2Rows];  // even
            for (u32 pi = tid; pi < 17; pi += nthreads) {
                const u32 k = total + 2 * pi;
                sxy[k >> 1] = make_float4(qn_, qn_, qn_, qn_);
                szz[k >> 1] = make_float2(qn_, qn_);
            }
            cp_async_wait_all();
        }
        __syncthreads();

        // ---- targets: 32 consecutive slots per warp ------------------------------------------------------
        const u32 S = s_tpre[4];
        const u32 limit = (s_rs[kT2Rows] + 24) & ~1u;  // last safe (even) start of an 8-candidate block
        for (u32 gb = 32u * (u32)warp; gb < S; gb += 32u * nwarps) {
            const u32 sl = gb + (u32)lane;
            int tr = (sl >= s_tpre[1]) + (sl >= s_tpre[2]) + (sl >= s_tpre[3]);
            if (sl >= S) tr = s_tlast;
            const u32 ntr = s_tn[tr];  // > 0: a row without targets has no slots
            const u32 o_raw = sl - s_tpre[tr];
            const bool valid = sl < S && o_raw < ntr;
            const u32 i = s_t0[tr] + min(o_raw, ntr - 1u);
            const int ta = tr >> 1, tb = tr & 1;
            const int Xt = X + ta, Yt = Y + tb;
            const int sr0 = ta * 4 + tb;  // staged row of the (-1, -1) neighbour row
            const int src = sr0 + 5;      // staged row of the target's own row
            // integrator inputs: only prefetched here (18 registers would otherwise be parked across the pair loop)
            if constexpr (!FORCE_ONLY) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    if (P.mech != CRB_MECH_BROWNIAN) prefetch_l1_t2(&P.A.vel[c][i]);
                    prefetch_l1_t2(&P.A.h1x[c][i]);
                    prefetch_l1_t2(&P.A.h1v[c][i]);
                }
            }
            const u32 sj_self = s_rs[src] + (i - s_g0[src]);
            TileTarget<D> t;
            {
                const double2 sa = sxyd[sj_self];
                const TileCandZ sb = szm[sj_self];
                t.x[0] = sa.x;
                t.x[1] = sa.y;
                if (D == 3) t.x[2] = sb.z;
                t.my_v = sb.cv;
                t.sp = (u8)(sb.meta & 0xffu);
                t.zw = sb.meta >> 8;
            }
#pragma unroll
            for (int c = 0; c < D; ++c) t.acc[c] = 0.0;
            t.calc_flag = 0;
            t.subv_lo = 0;
            t.subv_len = 0;
            const int zt = (int)(t.my_v - s_rowvox[src]);
            if (P.weight_mode == CRB_WEIGHT_CHILI && g.sm_nsub <= 1) {
                t.subv_len = (u32)g.nvox_local();  // one chili subdomain: everything is "local"
            } else if (P.weight_mode == CRB_WEIGHT_CHILI) {
                const u64 plane = (u64)n1 * n2;
                const u64 gp = ((u64)(Xt + g.p0) * n1 + Yt) * n2 + zt;
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
            float xf, yf, zf;
            {
                const float* fxy = reinterpret_cast<const float*>(sxy + (sj_self >> 1)) + (sj_self & 1);
                const float* fz = reinterpret_cast<const float*>(szz + (sj_self >> 1)) + (sj_self & 1);
                xf = valid ? fxy[0] : qnan;
                yf = fxy[2];
                zf = fz[0];
            }
            // every sub-group of kTileSub consecutive lanes (one target row, ascending z) streams its OWN
            // candidate window: the voxels zmin-1 .. zmax+1 of its targets
            const int zmin = __shfl_sync(0xffffffffu, zt, 0, kTileSub);
            const int zmax = __shfl_sync(0xffffffffu, zt, kTileSub - 1, kTileSub);
            const int zlo = max(zmin - 1, wa), zhi = min(zmax + 2, wb);

            // per-lane FIFO of in-range candidate indices (ascending): q[k][tid]
            u32 qa = qa0;
            // (runs for whichever lanes are converged at the call: all of them at the end of the pass, only the
            // lanes whose FIFO is nearly full inside the filter loop — a rare, divergent call instead of a vote per block)
            auto drain = [&]() {
                const u32 nq = (qa - qa0) / (2u * nthreads);
                const u32 nqmax = __reduce_max_sync(__activemask(), nq);
                for (u32 k = 0; k < nqmax; k += NE) {
                    u32 jj[NE];
                    bool hh[NE];
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        u32 ent;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ent) : "r"(qa0 + (k + e) * 2u * nthreads));
                        hh[e] = (k + e < nq) & (ent != sj_self);
                        jj[e] = hh[e] ? ent : sj_self;
                    }
                    tile_eval_soa<D, INTER, NE>(t, own, ext_radius, sxyd, reinterpret_cast<const double2*>(szm), jj, hh);
                }
                qa = qa0;
            };
            u32 qfull = qa0 + 2u * (kTileQueue - 8) * nthreads;
            asm volatile("" : "+r"(qfull));  // held in a register: one compare per block instead of add + compare
            const float2 xf2 = make_float2(-xf, -xf), yf2 = make_float2(-yf, -yf), zf2 = make_float2(-zf, -zf);
            // start[] entries of this sub-group's window inside row sr: a_sst + 4 (sr sstride + z - zb0)
            const u32 a_wlo = a_sst + 4u * (u32)(zlo - zb0), a_whi = a_sst + 4u * (u32)(zhi - zb0);
            const u32 rstride = 4u * (u32)sstride;

#pragma unroll 1
            for (int ra = 0; ra < 3; ++ra) {
#pragma unroll 1
                for (int rb = 0; rb < 3; ++rb) {
                    const int sr = sr0 + ra * 4 + rb;
                    const u32 off = s_off[sr];  // even for a row that is not staged
                    const u32 lo = off + lds_u32(a_wlo + (u32)sr * rstride), hi = off + lds_u32(a_whi + (u32)sr * rstride);
                    const u32 lo2 = lo & ~1u;  // blocks start at an even candidate (pair layout); the extra
                                               // leading candidate belongs to voxel zlo-1 of the same row:
                                               // rejected by the distance test or by the adjacency test
                    const u32 n = hi - lo2;
                    const u32 nmax = __reduce_max_sync(0xffffffffu, n);
                    if (nmax == 0) continue;  // warp-uniform
                    // x = first candidate of the lane's current block; a lane whose own range [lo2, hi) has ended reads
                    // the block of sentinels at `limit` (never pass) while the longest sub-group of the warp finishes
                    const u32 xend = lo2 + nmax;
                    for (u32 x = lo2; x < xend; x += 8) {
                        if (qa > qfull) drain();
                        // a block that starts inside the range may over-read <= 7 entries (later voxels of the row or
                        // sentinels)
                        const u32 b8 = x < hi ? x : limit;
                        const float cut = T.cut2f;
                        const u32 axy = a_sxy + b8 * 8u, azz = a_szz + b8 * 4u;
                        float4 ca4[4];
                        float2 cz2[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {  // all loads first: the queue stores below alias smem
                            ca4[u] = lds_f4(axy + 16u * u);
                            if (D == 3) cz2[u] = lds_f2(azz + 8u * u);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float2 dx = __fadd2_rn(make_float2(ca4[u].x, ca4[u].y), xf2);
                            const float2 dy = __fadd2_rn(make_float2(ca4[u].z, ca4[u].w), yf2);
                            float2 d2 = __fmul2_rn(dx, dx);
                            d2 = __ffma2_rn(dy, dy, d2);
                            if (D == 3) {
                                const float2 dz = __fadd2_rn(cz2[u], zf2);
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
            }
            drain();

            if (valid) {
                if (t.calc_flag) latch_error(P.err, CRB_ERR_CALC, P.A.id[i]);
                if constexpr (FORCE_ONLY) {
                    add_domain_force<D>(P, t.x, t.acc);
#pragma unroll
                    for (int c = 0; c < D; ++c) P.force_out[(u64)c * P.cap + i] = t.acc[c];
                } else {
                    integrate_agent<D, true>(P, i, t.my_v, SS(src, zt), Xt, Yt, zt, t.x, t.acc, load_agent_dyn<D>(P, i));
                }
            }
        }
        z += kc;
    }
#undef SS
}

}  // namespace crb
