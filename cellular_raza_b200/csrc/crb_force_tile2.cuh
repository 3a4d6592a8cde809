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
// Staging: the exact data of a candidate is one 32-byte record {x, y, z, meta, voxel} (TileCand) filled
// by cp.async; the FP32 filter copy is loaded as pairs and stored transposed ({x0,x1,y0,y1}, {z0,z1})
// so that one LDS.128 + one LDS.64 feed the packed f32x2 distance test of two candidates.  Every row
// starts at an even staged index and is followed by >= 8 sentinels (NaN: never pass the test).
#pragma once
#include "crb_force_tile.cuh"

namespace crb {

constexpr int kT2Rows = 16;  // staged rows (4 x 4)
constexpr int kT2Pad = 8;    // sentinels after every staged row

__host__ __device__ inline size_t tile2_smem_bytes(u32 cap, int kmax, int threads) {
    const size_t c = cap + 34;  // cap includes the rows' sentinels; +34 filter over-read (even)
    size_t b = 0;
    b += c * sizeof(TileCand);                                         // exact positions + meta + voxel
    b += (c / 2) * sizeof(float4);                                     // {x0, x1, y0, y1} per pair
    b += (c / 2) * sizeof(float2);                                     // {z0, z1} per pair
    b += (size_t)kT2Rows * (kmax + 1) * sizeof(u32);                   // start[] of the 16 rows
    b += (size_t)(kTileQueue + 4) * threads * sizeof(unsigned short);  // per-lane FIFOs
    return b + 64;
}

template <int D, int INTER, bool FORCE_ONLY, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_force_tile2(const KParams P, const TileLaunch T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int nthreads = NT;  // launch contract: blockDim.x == NT (128: 4 CTAs/SM, 256: 2 CTAs/SM)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = nthreads >> 5;
    constexpr int NE = 4;  // candidates evaluated per drain iteration and lane
    const Grid& g = P.g;
    const int n1 = g.ng[1], n2 = g.ng[2];

    // ---- carve shared memory -----------------------------------------------------------------------
    const u32 capp = T.cap + 34;  // even
    TileCand* sc = reinterpret_cast<TileCand*>(smem_raw);
    float4* sxy = reinterpret_cast<float4*>(sc + capp);                  // [capp / 2]
    float2* szz = reinterpret_cast<float2*>(sxy + capp / 2);             // [capp / 2]
    u32* sst = reinterpret_cast<u32*>(szz + capp / 2);                   // [16][kmax + 1] copy of start[]
    unsigned short* q = reinterpret_cast<unsigned short*>(sst + kT2Rows * (T.kmax + 1));
    __shared__ u32 s_rowvox[kT2Rows];   // first voxel (local plain index) of every staged row
    __shared__ u32 s_g0[kT2Rows];       // first global agent index staged per row
    __shared__ u32 s_rs[kT2Rows + 1];   // staged index of the first candidate of every row (even)
    __shared__ u32 s_rn[kT2Rows];       // candidates staged per row (without sentinels)
    __shared__ int s_valid[kT2Rows];
    __shared__ u32 s_tpre[5];           // first slot of every target row (multiples of 4)
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
                        cp_async_8(&sc[k].x, P.A.pos[0] + gi);
                        cp_async_8(&sc[k].y, P.A.pos[1] + gi);
                        if (D == 3) cp_async_8(&sc[k].z, P.A.pos[2] + gi);
                        cp_async_4(&sc[k].meta, P.A.meta + gi);
                    } else {
                        sc[k].meta = 0xffffu;
                    }
                    if (o + 1 < n) {
                        const u32 gi = g0 + o + 1;
                        b = __ldg(P.A.posf + gi);
                        cp_async_8(&sc[k + 1].x, P.A.pos[0] + gi);
                        cp_async_8(&sc[k + 1].y, P.A.pos[1] + gi);
                        if (D == 3) cp_async_8(&sc[k + 1].z, P.A.pos[2] + gi);
                        cp_async_4(&sc[k + 1].meta, P.A.meta + gi);
                    } else {
                        sc[k + 1].meta = 0xffffu;
                    }
                    sxy[k >> 1] = make_float4(a.x, b.x, a.y, b.y);
                    szz[k >> 1] = make_float2(a.z, b.z);
                    sc[k].cv = __float_as_uint(a.w);
                    sc[k + 1].cv = __float_as_uint(b.w);
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
            // integrator inputs: issued now, consumed after the pair loop
            AgentDyn<D> dyn;
            if constexpr (!FORCE_ONLY) dyn = load_agent_dyn<D>(P, i);
            const u32 sj_self = s_rs[src] + (i - s_g0[src]);
            const TileCand self = sc[sj_self];
            TileTarget<D> t;
            t.x[0] = self.x;
            t.x[1] = self.y;
            if (D == 3) t.x[2] = self.z;
#pragma unroll
            for (int c = 0; c < D; ++c) t.acc[c] = 0.0;
            t.my_v = self.cv;
            t.sp = (u8)(self.meta & 0xffu);
            t.zw = self.meta >> 8;
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

#pragma unroll 1
            for (int r = 0; r < 9; ++r) {
                const int sr = sr0 + (r / 3) * 4 + (r % 3);
                const u32 off = s_rs[sr] - s_g0[sr];  // even for a row that is not staged
                const u32 lo = off + SS(sr, zlo), hi = off + SS(sr, zhi);
                const u32 lo2 = lo & ~1u;  // blocks start at an even candidate (pair layout); the extra
                                           // leading candidate belongs to voxel zlo-1 of the same row:
                                           // rejected by the distance test or by the adjacency test
                const u32 n = hi - lo2;
                const u32 nmax = __reduce_max_sync(0xffffffffu, n);
                if (nmax == 0) continue;  // warp-uniform
                for (u32 k0 = 0; k0 < nmax; k0 += 8) {
                    if (__any_sync(0xffffffffu, qa > qfull)) drain();
                    // a block past the lane's range is disabled through the threshold; a block that
                    // starts inside may over-read <= 7 entries (later voxels of the row or sentinels)
                    const float cut = k0 < n ? T.cut2f : -1.0f;
                    const u32 b8 = min(lo2 + k0, limit);
                    const u32 p4 = b8 >> 1;
                    float4 ca4[4];
                    float2 cz2[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {  // all loads first: the queue stores below alias smem
                        ca4[u] = sxy[p4 + u];
                        if (D == 3) cz2[u] = szz[p4 + u];
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
            drain();

            if (valid) {
                if (t.calc_flag) latch_error(P.err, CRB_ERR_CALC, P.A.id[i]);
                if constexpr (FORCE_ONLY) {
                    add_domain_force<D>(P, t.x, t.acc);
#pragma unroll
                    for (int c = 0; c < D; ++c) P.force_out[(u64)c * P.cap + i] = t.acc[c];
                } else {
                    integrate_agent<D, true>(P, i, t.my_v, SS(src, zt), Xt, Yt, zt, t.x, t.acc, dyn);
                }
            }
        }
        z += kc;
    }
#undef SS
}

}  // namespace crb
