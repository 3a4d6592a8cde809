// crb_rods_host.cuh — host side of the bacterial-rods path (included by crb_api.cu after the
// definition of crb_sim and its helpers).  Same C ABI entry points as the point-agent path; the
// functions here are what those entry points dispatch to when mechanics_kind == CRB_MECH_ROD.
#pragma once

namespace {

enum { RMP_DIFF = 0, RMP_TENSION, RMP_RIGIDITY, RMP_LENGTH, RMP_DAMPING, RMP_GROWTH, RMP_THRESHOLD, RMP_GROWTH_BASE };

int rods_alloc_state(crb_sim* s, RodState& st, bool dynamic, bool stat) {
    std::memset(&st, 0, sizeof(st));
    const u64 c = s->cap_total, W = 3ull * s->V;
    int rc;
    if (dynamic) {
        if ((rc = dalloc(s, &st.pos, c * W)) || (rc = dalloc(s, &st.vel, c * W)) ||
            (rc = dalloc(s, &st.h1x, c * W)) || (rc = dalloc(s, &st.h1v, c * W)) ||
            (rc = dalloc(s, &st.h0x, c * W)) || (rc = dalloc(s, &st.h0v, c * W)) ||
            (rc = dalloc(s, &st.len, c)) || (rc = dalloc(s, &st.growth, c)) ||
            (rc = dalloc(s, &st.cen, c * 4)) || (rc = dalloc(s, &st.hlen, c)))
            return rc;
    }
    if (stat) {
        if ((rc = dalloc(s, &st.ident, c * 3)) || (rc = dalloc(s, &st.parent, c * 3))) return rc;
    }
    return 0;
}

int rods_alloc(crb_sim* s) {
    int rc;
    if ((rc = rods_alloc_state(s, s->RX, true, true))) return rc;
    if ((rc = rods_alloc_state(s, s->RY, true, true))) return rc;
    if ((rc = rods_alloc_state(s, s->RT, true, false))) return rc;
    if ((rc = dalloc(s, &s->div_flag, s->cap_total)) || (rc = dalloc(s, &s->id_counter, s->nvox + 1)) ||
        (rc = dalloc(s, &s->n_div_dev, 4)) || (rc = dalloc(s, &s->div_list, s->cap_total)))
        return rc;
    CU(s, cudaMemsetAsync(s->div_flag, 0, s->cap_total, s->stream));
    CU(s, cudaMemsetAsync(s->id_counter, 0, (s->nvox + 1) * sizeof(u32), s->stream));
    return 0;
}

// smallest double x with sqrt(x) >= c (sqrt correctly rounded, monotone): (norm < c) <=> (norm^2 < x)
double sqrt_threshold(double c) {
    double x = c * c;
    while (std::sqrt(x) >= c) x = std::nextafter(x, 0.0);
    while (std::sqrt(x) < c) x = std::nextafter(x, INFINITY);
    return x;
}

RodParams rods_params(crb_sim* s) {
    RodParams P;
    std::memset(&P, 0, sizeof(P));
    P.g = s->g;
    P.A = s->RX;
    P.B = s->RT;
    P.n = s->n;
    P.first = s->G;
    P.src_limit = s->cap_total - s->G;  // [first, src_limit): owned rods, daughters, arrivals
    P.cap = s->cap_total;
    P.key = s->key;
    P.start = s->start;
    P.newkey = s->newkey;
    P.rank_arb = s->rank_arb;
    P.count_new = s->count_new;
    P.sortkey = s->sortkey;
    P.div_flag = s->div_flag;
    P.id_counter = s->id_counter;
    P.n_div = s->n_div_dev;
    P.div_list = s->div_list;
    P.force_out = s->force_out;
    P.nb_out = s->nb_out;
    P.err = s->err;
    P.dt = s->cfg.dt;
    const double* mp = s->uni.mp;
    const double* ip = s->uni.ip;
    P.tension = mp[RMP_TENSION];
    P.rigidity = mp[RMP_RIGIDITY];
    P.damping = mp[RMP_DAMPING];
    P.threshold = mp[RMP_THRESHOLD];
    P.growth_base = mp[RMP_GROWTH_BASE];
    P.radius = ip[0];
    P.stiffness = ip[1];
    P.cutoff = ip[2];
    P.strength = ip[3];
    const double two_r = 2.0 * P.radius;
    P.reject = std::max(P.cutoff, two_r) * (1.0 + 1e-9);
    P.nb_thresh2 = sqrt_threshold(two_r);
    P.order = (int)s->cfg.solver_order;
    P.weight_mode = (int)s->cfg.pair_weight_mode;
    P.save_prev = 0;
    P.noise_mode = mp[RMP_DIFF] != 0.0 ? (int)s->cfg.noise_mode : CRB_NOISE_NONE;  // D == 0: dpos is exactly 0
    P.noise = s->noise;
    P.noise_ids = s->noise_ids;
    P.noise_cursor = s->noise_cursor;
    P.iteration = s->iteration;
    P.seed = s->cfg.rng_seed;
    P.diffusion = mp[RMP_DIFF];
    P.dforce = (int)s->cfg.domain_force_kind;
    P.gel_pressure = s->cfg.domain_force_strength[0];
    P.surface_friction = s->cfg.domain_force_strength[1];
    P.surface_friction_distance = s->cfg.domain_force_strength[2];
    return P;
}

template <int V>
void rods_launch_force_v(crb_sim* s, const RodParams& P, bool force_only) {
    const u32 grid = blocks_for(P.n, kRodThreads);
    if (!grid) return;
    // (register caps of 168 / 128 for 3 / 4 resident CTAs were measured: 4.27 / 4.63 ms against 3.7 ms at
    // 2 M rods — the spills cost more than the occupancy gives)
    if (force_only)
        k_rod_force<V, true><<<grid, kRodThreads, 0, s->stream>>>(P);
    else
        k_rod_force<V, false><<<grid, kRodThreads, 0, s->stream>>>(P);
}
template <int V>
void rods_launch_update_v(crb_sim* s, const RodParams& P) {
    const u32 grid = blocks_for(P.n, 128);
    if (grid) k_rod_update<V><<<grid, 128, 0, s->stream>>>(P);
}
template <int V>
void rods_launch_divide_v(crb_sim* s, const RodParams& P) {
    // a fixed grid that strides over the division list (its length lives on the device)
    if (P.n) k_rod_divide<V><<<148 * 2, 128, 0, s->stream>>>(P);
}
template <int V>
void rods_launch_ingest_v(crb_sim* s, const RodParams& P) {
    const u32 grid = blocks_for(P.n, 256);
    if (grid) k_rod_ingest<V><<<grid, 256, 0, s->stream>>>(P);
}
#define RODS_DISPATCH_V(s, CALL)                      \
    switch ((s)->V) {                                 \
        case 2: CALL(2); break;                       \
        case 3: CALL(3); break;                       \
        case 4: CALL(4); break;                       \
        case 5: CALL(5); break;                       \
        case 6: CALL(6); break;                       \
        case 7: CALL(7); break;                       \
        case 8: CALL(8); break;                       \
    }

void rods_launch_update(crb_sim* s, const RodParams& P) {
#define CALL(VV) rods_launch_update_v<VV>(s, P)
    RODS_DISPATCH_V(s, CALL);
#undef CALL
}
// split step: pair forces (FORCE_ONLY kernel -> force_out / nb_out) and the per-rod epilogue as two kernels
inline bool rods_split() {
    static const bool v = !(std::getenv("CRB_ROD_SPLIT") && std::atoi(std::getenv("CRB_ROD_SPLIT")) == 0);
    return v;
}
void rods_launch_force(crb_sim* s, const RodParams& P, bool force_only) {
#define CALL(VV) rods_launch_force_v<VV>(s, P, force_only)
    RODS_DISPATCH_V(s, CALL)
#undef CALL
}
void rods_launch_divide(crb_sim* s, const RodParams& P) {
#define CALL(VV) rods_launch_divide_v<VV>(s, P)
    RODS_DISPATCH_V(s, CALL)
#undef CALL
}
void rods_launch_ingest(crb_sim* s, const RodParams& P) {
#define CALL(VV) rods_launch_ingest_v<VV>(s, P)
    RODS_DISPATCH_V(s, CALL)
#undef CALL
}

// scan count_new -> start_alt, fill, reorder RT (+ RX identifiers) -> RY, swap
// n_src sources at [G, G + n_src) (owned rods, daughters, arrivals; leavers are marked CRB_KEY_GONE), n_new of
// them remain = destination slots [G, G + n_new)
int rods_sort_into_place(crb_sim* s, u64 n_src, bool save_prev, u64 n_new) {
    const u64 nv = s->nvox;
    stage_begin(s, 1);
    if (s->n_tiles == 1) {
        k_scan_single<<<1, kScanThreads, 0, s->stream>>>(s->count_new, s->start_alt, nv, (u32)s->G);
        stage_end(s, 1, 1);
    } else {
        k_scan_tiles<<<s->n_tiles, kScanThreads, 0, s->stream>>>(s->count_new, s->start_alt, s->tile_sums, nv);
        k_scan_sums<<<1, kScanThreads, 0, s->stream>>>(s->tile_sums, s->n_tiles, s->start_alt + nv, (u32)s->G);
        k_scan_add<<<s->n_tiles, kScanThreads, 0, s->stream>>>(s->start_alt, s->tile_sums, nv);
        stage_end(s, 1, 3);
    }
    stage_begin(s, 2);
    int nl = 0;
    if (n_src) {
        k_fill<<<blocks_for(n_src, 256 * kFillPerThread), 256, 0, s->stream>>>(s->newkey, s->rank_arb, s->start_alt, s->tmp_idx,
                                                             s->G, n_src);
        RodReorderParams R;
        std::memset(&R, 0, sizeof(R));
        R.A = s->RX;
        R.B = s->RT;
        R.Y = s->RY;
        R.newkey = s->newkey;
        R.start_new = s->start_alt;
        R.tmp_idx = s->tmp_idx;
        R.sortkey = s->sortkey;
        R.key_out = s->key_alt;
        R.n = n_new;
        R.first = s->G;
        R.cap = s->cap_total;
        R.W = 3 * s->V;
        R.save_prev = save_prev ? 1 : 0;
        if (n_new) k_rod_reorder<<<blocks_for(n_new, 256), 256, 0, s->stream>>>(R);
        nl = 2;
    }
    stage_end(s, 2, nl);
    std::swap(s->RX, s->RY);
    std::swap(s->key, s->key_alt);
    std::swap(s->start, s->start_alt);
    CU(s, cudaGetLastError());
    return 0;
}

// agent-major host rows of width W -> planes of stride cap_total (through the staging buffer)
int rods_upload_planes(crb_sim* s, const double* host, int W, double* planes, u64 n) {
    if (!n) return 0;
    CU(s, cudaMemcpyAsync(s->stage, host, n * W * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    k_rod_unpack<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->stage, planes + s->G, W, s->cap_total, n);
    s->launches++;
    return 0;
}
int rods_download_planes(crb_sim* s, double* host, int W, const double* planes, u64 n) {
    if (!n) return 0;
    k_rod_pack<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->stage, planes + s->G, W, s->cap_total, n);
    s->launches++;
    CU(s, cudaMemcpyAsync(host, s->stage, n * W * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    return 0;
}

int rods_set_agents(crb_sim* s, u64 n, const double* pos, const double* vel, const double* mech_params,
                    const double* inter_params, const u64* ids) {
    const int W = 3 * s->V;
    // RodMechanics / MorsePotential fields are uniform except spring_length and growth_rate
    std::vector<double> len(n), growth(n);
    if (s->world > 1 && !s->uniform_declared)
        return fail(s, CRB_ERR_INVALID_ARG,
                    "rods on a slab handle: declare the uniform RodMechanics / MorsePotential rows with "
                    "crb_set_uniform_params on EVERY rank (a slab that starts empty has no row of its own)");
    if (s->uniform_declared && !mech_params) {
        for (u64 i = 0; i < n; ++i) {
            len[i] = s->uni.mp[RMP_LENGTH];
            growth[i] = s->uni.mp[RMP_GROWTH];
        }
    } else {
        if (n && (!mech_params || (!inter_params && !s->uniform_declared)))
            return fail(s, CRB_ERR_INVALID_ARG, "mech_params and inter_params are required");
        if (n) {
            for (int k = 0; k < 8; ++k) s->uni.mp[k] = mech_params[k];
            if (inter_params)
                for (int k = 0; k < 4; ++k) s->uni.ip[k] = inter_params[k];
        }
        // one pass over the rows on all host cores (2 M rods are 190 MB of parameters: a single-threaded scan with an
        // early exit was two thirds of crb_set_agents).  Branch-free inner loop: mismatches are OR-ed per chunk.
        const double um[8] = {s->uni.mp[0], s->uni.mp[1], s->uni.mp[2], s->uni.mp[3],
                              s->uni.mp[4], s->uni.mp[5], s->uni.mp[6], s->uni.mp[7]};
        const double ui[4] = {s->uni.ip[0], s->uni.ip[1], s->uni.ip[2], s->uni.ip[3]};
        const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        const unsigned nth = n < 65536 ? 1u : hw;
        std::vector<int> bad_m(nth, 0), bad_i(nth, 0);
        auto scan = [&](unsigned t) {
            const u64 a = n * t / nth, b = n * (t + 1) / nth;
            int bm = 0, bi = 0;
            for (u64 i = a; i < b; ++i) {
                const double* r = mech_params + i * 8;
                bm |= (r[0] != um[0]) | (r[1] != um[1]) | (r[2] != um[2]) | (r[4] != um[4]) | (r[6] != um[6]) | (r[7] != um[7]);
                len[i] = r[RMP_LENGTH];
                growth[i] = r[RMP_GROWTH];
            }
            if (inter_params)
                for (u64 i = a; i < b; ++i) {
                    const double* r = inter_params + i * 4;
                    bi |= (r[0] != ui[0]) | (r[1] != ui[1]) | (r[2] != ui[2]) | (r[3] != ui[3]);
                }
            bad_m[t] = bm;
            bad_i[t] = bi;
        };
        if (nth == 1) {
            scan(0);
        } else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nth; ++t) th.emplace_back(scan, t);
            for (auto& x : th) x.join();
        }
        for (unsigned t = 0; t < nth; ++t) {
            if (bad_m[t])
                return fail(s, CRB_ERR_INVALID_ARG,
                            "rods: all mechanics parameters except spring_length and growth_rate must be uniform");
            if (bad_i[t]) return fail(s, CRB_ERR_INVALID_ARG, "rods: interaction parameters must be uniform");
        }
    }
    // the device evaluates the Morse exponential without overflow handling (exp_bounded): its largest
    // argument, potential_stiffness * (R1 + R2) at distance 0, must stay below the f64 overflow of exp
    if (n && !(s->uni.ip[1] * (s->uni.ip[0] + s->uni.ip[0]) < 700.0))
        return fail(s, CRB_ERR_CALC, "CalcError: MorsePotential potential_stiffness * 2 radius >= 700 overflows exp in f64");
    if (s->uni.mp[RMP_DIFF] != 0.0 && s->cfg.noise_mode == CRB_NOISE_NONE)
        return fail(s, CRB_ERR_INVALID_ARG,
                    "rods: diffusion_constant != 0 needs noise_mode CRB_NOISE_TABLE or CRB_NOISE_PHILOX "
                    "(RodMechanics::get_random_contribution, bacterial_rods.rs:215-240)");
    s->uniform = true;
    int rc;
    if ((rc = rods_upload_planes(s, pos, W, s->RT.pos, n))) return rc;
    if (vel) {
        CU(s, cudaStreamSynchronize(s->stream));
        if ((rc = rods_upload_planes(s, vel, W, s->RT.vel, n))) return rc;
    } else {
        CU(s, cudaMemsetAsync(s->RT.vel, 0, s->cap_total * W * sizeof(double), s->stream));
    }
    for (double* p : {s->RT.h1x, s->RT.h1v, s->RT.h0x, s->RT.h0v})
        CU(s, cudaMemsetAsync(p, 0, s->cap_total * W * sizeof(double), s->stream));
    CU(s, cudaMemsetAsync(s->RT.hlen, 0, s->cap_total, s->stream));
    CU(s, cudaMemcpyAsync(s->RT.len + s->G, len.data(), n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(s, cudaMemcpyAsync(s->RT.growth + s->G, growth.data(), n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    u64* d_ids = nullptr;
    if (ids && n) {
        CU(s, cudaStreamSynchronize(s->stream));
        d_ids = (u64*)s->stage;
        CU(s, cudaMemcpyAsync(d_ids, ids, n * sizeof(u64), cudaMemcpyHostToDevice, s->stream));
    }
    if (n) {
        k_rod_init_ident<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->RX, d_ids, s->cap_total, n, s->G);
        s->launches++;
    }
    CU(s, cudaMemsetAsync(s->count_new, 0, (s->nvox + 1) * sizeof(u32), s->stream));
    CU(s, cudaMemsetAsync(s->div_flag, 0, s->cap_total, s->stream));
    s->n = n;
    RodParams P = rods_params(s);
    rods_launch_ingest(s, P);
    s->launches += n ? 1 : 0;
    if ((rc = check_latch(s))) {
        s->n = 0;
        s->have_agents = false;
        return rc;
    }
    if ((rc = rods_sort_into_place(s, n, true, n))) return rc;
    k_rod_init_counters<<<blocks_for(s->nvox, 256), 256, 0, s->stream>>>(s->start, s->id_counter, s->nvox);
    s->launches++;
    CU(s, cudaStreamSynchronize(s->stream));
    CU(s, cudaGetLastError());
    s->have_agents = true;
    s->aux_settable = false;
    s->last_step_n = 0;
    s->last_divisions = 0;
    s->noise_cursor = 0;
    return 0;
}

// crb_set_agents_async for rods: everything is queued on the handle's stream and nothing waits for the device.  All
// host transfers go first (into one arena: positions, velocities, parameter rows, ids side by side), then the kernels
// (a transfer queued behind a kernel would wait for that kernel to find free SMs next to the step of the handle this
// one alternates with); clears are kernels; the parameter rows are checked on the device (k_rod_param_rows); an
// ingestion error stays latched until the next synchronising call (crb_step of a rod handle always is one).
int rods_set_agents_async(crb_sim* s, u64 n, const double* pos, const double* vel, const double* mech_params,
                          const double* inter_params, const u64* ids) {
    const int W = 3 * s->V;
    if (s->world > 1) return fail(s, CRB_ERR_STATE, "crb_set_agents_async: not available for rods on a slab handle");
    if (!mech_params && !s->uniform_declared)
        return fail(s, CRB_ERR_INVALID_ARG, "mech_params and inter_params are required (or crb_set_uniform_params)");
    if (n && mech_params) {
        if (!inter_params && !s->uniform_declared)
            return fail(s, CRB_ERR_INVALID_ARG, "mech_params and inter_params are required");
        for (int k = 0; k < 8; ++k) s->uni.mp[k] = mech_params[k];  // (page-locked host memory: readable here)
        if (inter_params)
            for (int k = 0; k < 4; ++k) s->uni.ip[k] = inter_params[k];
    }
    if (n && !(s->uni.ip[1] * (s->uni.ip[0] + s->uni.ip[0]) < 700.0))
        return fail(s, CRB_ERR_CALC, "CalcError: MorsePotential potential_stiffness * 2 radius >= 700 overflows exp in f64");
    if (s->uni.mp[RMP_DIFF] != 0.0 && s->cfg.noise_mode == CRB_NOISE_NONE)
        return fail(s, CRB_ERR_INVALID_ARG,
                    "rods: diffusion_constant != 0 needs noise_mode CRB_NOISE_TABLE or CRB_NOISE_PHILOX "
                    "(RodMechanics::get_random_contribution, bacterial_rods.rs:215-240)");
    s->uniform = true;
    s->kernels_only = true;
    int rc;
    const u64 per = (u64)(2 * W + 8 + 4 + 1);
    if (!s->rod_up && (rc = dalloc(s, &s->rod_up, s->cap * per))) return rc;
    double* u_pos = s->rod_up;
    double* u_vel = u_pos + n * W;
    double* u_mp = u_vel + n * W;
    double* u_ip = u_mp + n * 8;
    u64* u_id = (u64*)(u_ip + n * 4);
    if (n) {
        CU(s, cudaMemcpyAsync(u_pos, pos, n * W * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (vel) CU(s, cudaMemcpyAsync(u_vel, vel, n * W * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (mech_params) CU(s, cudaMemcpyAsync(u_mp, mech_params, n * 8 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (mech_params && inter_params)
            CU(s, cudaMemcpyAsync(u_ip, inter_params, n * 4 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (ids) CU(s, cudaMemcpyAsync(u_id, ids, n * sizeof(u64), cudaMemcpyHostToDevice, s->stream));
    }
    RodZeroRanges Z;
    std::memset(&Z, 0, sizeof(Z));
    for (double* p : {s->RT.h1x, s->RT.h1v, s->RT.h0x, s->RT.h0v}) {
        Z.p[Z.n] = p;
        Z.count[Z.n++] = s->cap_total * W;
    }
    if (!vel) {
        Z.p[Z.n] = s->RT.vel;
        Z.count[Z.n++] = s->cap_total * W;
    }
    Z.b[0] = s->RT.hlen;
    Z.nb[0] = s->cap_total;
    Z.b[1] = s->div_flag;
    Z.nb[1] = s->cap_total;
    k_rod_zero<<<148 * 8, 256, 0, s->stream>>>(Z);
    k_zero_u32<<<blocks_for(s->nvox + 1, 256), 256, 0, s->stream>>>(s->count_new, s->nvox + 1, nullptr, 0, nullptr, 0);
    s->launches += 2;
    if (n) {
        k_rod_unpack<<<blocks_for(n, 256), 256, 0, s->stream>>>(u_pos, s->RT.pos + s->G, W, s->cap_total, n);
        if (vel) k_rod_unpack<<<blocks_for(n, 256), 256, 0, s->stream>>>(u_vel, s->RT.vel + s->G, W, s->cap_total, n);
        RodUniformRow U;
        for (int k = 0; k < 8; ++k) U.mp[k] = s->uni.mp[k];
        for (int k = 0; k < 4; ++k) U.ip[k] = s->uni.ip[k];
        k_rod_param_rows<<<blocks_for(n, 256), 256, 0, s->stream>>>(mech_params ? u_mp : nullptr,
                                                                   mech_params && inter_params ? u_ip : nullptr, U,
                                                                   s->RT.len + s->G, s->RT.growth + s->G, n, s->err);
        k_rod_init_ident<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->RX, ids ? u_id : nullptr, s->cap_total, n, s->G);
        s->launches += vel ? 4 : 3;
    }
    s->n = n;
    RodParams P = rods_params(s);
    rods_launch_ingest(s, P);
    s->launches += n ? 1 : 0;
    if ((rc = rods_sort_into_place(s, n, true, n))) return rc;
    k_rod_init_counters<<<blocks_for(s->nvox, 256), 256, 0, s->stream>>>(s->start, s->id_counter, s->nvox);
    s->launches++;
    CU(s, cudaGetLastError());
    s->have_agents = true;
    s->aux_settable = false;
    s->last_step_n = 0;
    s->last_divisions = 0;
    s->noise_cursor = 0;
    return 0;
}

// crb_snapshot_async for rods: positions and velocities as agent-major V x 3 rows (the layout of crb_get_agents), the
// number of the identifier, the voxel; packed on the handle's stream, copied on the side stream
int rods_snapshot_pack(crb_sim* s, double* b_pos, double* b_vel, u64* b_id, u64* b_vox, bool pos, bool vel, bool ids,
                       bool voxel) {
    const u64 n = s->n;
    const int W = 3 * s->V;
    if (!n) return 0;
    if (pos) {
        k_rod_pack<<<blocks_for(n, 256), 256, 0, s->stream>>>(b_pos, s->RX.pos + s->G, W, s->cap_total, n);
        s->launches++;
    }
    if (vel) {
        k_rod_pack<<<blocks_for(n, 256), 256, 0, s->stream>>>(b_vel, s->RX.vel + s->G, W, s->cap_total, n);
        s->launches++;
    }
    if (ids) {
        k_copy_u64<<<blocks_for(n, 256), 256, 0, s->stream>>>(b_id, s->RX.ident + s->cap_total + s->G, n);
        s->launches++;
    }
    if (voxel) {
        const u64 base = (u64)s->g.p0 * s->g.ng[1] * s->g.ng[2];
        k_global_keys<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->key + s->G, b_vox, base, n);
        s->launches++;
    }
    CU(s, cudaGetLastError());
    return 0;
}

// stage A of one step: pair forces + per-rod update + divisions for the owned rods (ghost planes must hold the
// neighbours' boundary rods on a slab); the number of daughters is read back (the sources of the rebuild are
// [G, G + n + n_div))
int rods_compute(crb_sim* s, bool last) {
    if (s->cfg.noise_mode == CRB_NOISE_TABLE && s->uni.mp[RMP_DIFF] != 0.0 && s->noise_cursor >= s->noise_steps)
        return fail(s, CRB_ERR_STATE, "noise table exhausted: call crb_set_noise");
    if (s->kernels_only) {
        k_zero_u32<<<blocks_for(s->nvox + 1, 256), 256, 0, s->stream>>>(s->count_new, s->nvox + 1, s->n_div_dev, 2, nullptr, 0);
        s->launches++;
    } else {
        CU(s, cudaMemsetAsync(s->count_new, 0, (s->nvox + 1) * sizeof(u32), s->stream));
        CU(s, cudaMemsetAsync(s->n_div_dev, 0, 2 * sizeof(u32), s->stream));
    }
    RodParams P = rods_params(s);
    P.save_prev = last ? 1 : 0;
    stage_begin(s, 0);
    if (rods_split()) {
        int rc0;
        if (!s->force_out && (rc0 = dalloc(s, &s->force_out, s->cap_total * 3 * s->V))) return rc0;
        if (!s->nb_out && (rc0 = dalloc(s, &s->nb_out, s->cap_total))) return rc0;
        P.force_out = s->force_out;
        P.nb_out = s->nb_out;
        rods_launch_force(s, P, true);
        rods_launch_update(s, P);
        stage_end(s, 0, s->n ? 2 : 0);
    } else {
        rods_launch_force(s, P, false);
        stage_end(s, 0, s->n ? 1 : 0);
    }
    stage_begin(s, 4);
    rods_launch_divide(s, P);
    if (s->n)
        k_rod_bump_counters<<<148, 256, 0, s->stream>>>(s->div_list, s->n_div_dev, s->key, s->id_counter);
    stage_end(s, 4, s->n ? 2 : 0);
    u32 n_div = 0;
    if (s->kernels_only) {  // (see check_latch)
        k_rod_report<<<1, 32, 0, s->stream>>>(s->n_div_dev, nullptr, s->cnt_host + 4, nullptr);
        s->launches++;
        CU(s, cudaStreamSynchronize(s->stream));
        n_div = (u32)s->cnt_host[4];
    } else {
        CU(s, cudaMemcpyAsync(&n_div, s->n_div_dev, sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
        CU(s, cudaStreamSynchronize(s->stream));
    }
    CU(s, cudaGetLastError());
    if (s->n + n_div > s->cap + (s->world > 1 ? s->mcap : 0)) {
        check_latch(s);
        return fail(s, CRB_ERR_CAPACITY, "more agents than capacity after division");
    }
    s->last_step_n = s->n;
    s->last_divisions = n_div;
    s->n_src = s->n + n_div;
    s->n_gone = 0;
    return 0;
}
// stage B: cell-list rebuild over the sources of stage A (+ arrivals, - leavers on a slab)
int rods_finish(crb_sim* s, bool last) {
    const u64 n_new = s->n_src - s->n_gone;
    if (n_new > s->cap) return fail(s, CRB_ERR_CAPACITY, "more agents than capacity after division / migration");
    int rc;
    if ((rc = rods_sort_into_place(s, s->n_src, last, n_new))) return rc;
    s->n = n_new;
    s->iteration++;
    s->noise_cursor++;
    return 0;
}

int rods_step(crb_sim* s, u64 n_steps) {
    for (u64 it = 0; it < n_steps; ++it) {
        const bool last = it + 1 == n_steps;
        int rc;
        if ((rc = rods_compute(s, last))) return rc;
        if ((rc = rods_finish(s, last))) return rc;
    }
    return check_latch(s);
}

// ---- slab protocol (crb_slab_* with mechanics_kind == CRB_MECH_ROD): same six phases as the point agents ------------
int rods_slab_halo_export(crb_sim* s, void* lo_buf, void* hi_buf) {
    const u64 plane = s->rhl.plane;
    const u32 grid = blocks_for(std::max<u64>(s->rhl.gcap, plane + 1), 256);
    stage_begin(s, 3);
    int nl = 0;
    if (s->has_lo && lo_buf) {
        k_rod_halo_export<<<grid, 256, 0, s->stream>>>(s->RX, s->cap_total, s->start, (u32)(s->g.own_lo * plane), s->rhl,
                                                      (unsigned char*)lo_buf, s->err);
        ++nl;
    }
    if (s->has_hi && hi_buf) {
        k_rod_halo_export<<<grid, 256, 0, s->stream>>>(s->RX, s->cap_total, s->start, (u32)((s->g.own_hi - 1) * plane),
                                                      s->rhl, (unsigned char*)hi_buf, s->err);
        ++nl;
    }
    stage_end(s, 3, nl);
    CU(s, cudaGetLastError());
    return 0;
}
int rods_slab_halo_import(crb_sim* s, const void* lo_buf, const void* hi_buf) {
    const u64 plane = s->rhl.plane;
    const u32 grid = blocks_for(std::max<u64>(s->rhl.gcap, plane + 1), 256);
    stage_begin(s, 3);
    int nl = 0;
    if (s->has_lo) {
        if (!lo_buf) return fail(s, CRB_ERR_INVALID_ARG, "lower halo buffer missing");
        k_rod_halo_import<<<grid, 256, 0, s->stream>>>(s->RX, s->cap_total, s->start, 0u, s->G, 1, s->rhl,
                                                      (const unsigned char*)lo_buf, s->err);
        ++nl;
    }
    if (s->has_hi) {
        if (!hi_buf) return fail(s, CRB_ERR_INVALID_ARG, "upper halo buffer missing");
        k_rod_halo_import<<<grid, 256, 0, s->stream>>>(s->RX, s->cap_total, s->start, (u32)((s->g.L0 - 1) * plane),
                                                      s->G + s->n, 0, s->rhl, (const unsigned char*)hi_buf, s->err);
        ++nl;
    }
    stage_end(s, 3, nl);
    CU(s, cudaGetLastError());
    return 0;
}
RodMigParams rods_mig_params(crb_sim* s) {
    RodMigParams M;
    std::memset(&M, 0, sizeof(M));
    const u64 plane = s->rhl.plane;
    M.A = s->RX;
    M.B = s->RT;
    M.cap = s->cap_total;
    M.newkey = s->newkey;
    M.sortkey = s->sortkey;
    M.rank_arb = s->rank_arb;
    M.count_new = s->count_new;
    M.lo_end = s->has_lo ? (u32)plane : 0u;
    M.hi_begin = s->has_hi ? (u32)((s->g.L0 - 1) * plane) : 0xffffffffu;
    M.vox_base = (u64)s->g.p0 * plane;
    M.nvox = s->nvox;
    M.L = s->rml;
    M.counter = s->scratch_u64;
    M.err = s->err;
    return M;
}
int rods_slab_migration_export(crb_sim* s, void* lo_buf, void* hi_buf, uint64_t n_out[2]) {
    const u64 plane = s->rhl.plane;
    if ((s->has_lo && !lo_buf) || (s->has_hi && !hi_buf)) return fail(s, CRB_ERR_INVALID_ARG, "migration buffer missing");
    CU(s, cudaMemsetAsync(s->scratch_u64, 0, 2 * sizeof(u64), s->stream));
    stage_begin(s, 4);
    int nl = 0;
    if (s->n_src) {
        RodMigParams M = rods_mig_params(s);
        M.first = s->G;
        M.n = s->n_src;  // owned rods of the step and the daughters born in it
        M.buf[0] = (unsigned char*)lo_buf;
        M.buf[1] = (unsigned char*)hi_buf;
        k_rod_mig_export<<<blocks_for(s->n_src, 256), 256, 0, s->stream>>>(M);
        ++nl;
    }
    // the leavers were counted in the histogram of the ghost planes: they take no slot here
    if (s->has_lo) CU(s, cudaMemsetAsync(s->count_new, 0, plane * sizeof(u32), s->stream));
    if (s->has_hi) CU(s, cudaMemsetAsync(s->count_new + (s->g.L0 - 1) * plane, 0, plane * sizeof(u32), s->stream));
    if (s->has_lo) CU(s, cudaMemcpyAsync(lo_buf, s->scratch_u64, sizeof(u64), cudaMemcpyDeviceToDevice, s->stream));
    if (s->has_hi) CU(s, cudaMemcpyAsync(hi_buf, s->scratch_u64 + 1, sizeof(u64), cudaMemcpyDeviceToDevice, s->stream));
    stage_end(s, 4, nl);
    CU(s, cudaGetLastError());
    if (!n_out) {
        s->mig_counts_pending = true;
        return 0;
    }
    u64 cnt[2] = {0, 0};
    CU(s, cudaMemcpyAsync(cnt, s->scratch_u64, 2 * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    if (cnt[0] > s->mcap || cnt[1] > s->mcap)
        return fail(s, CRB_ERR_CAPACITY, "more migrating rods than migration_capacity");
    n_out[0] = s->mig_out[0] = cnt[0];
    n_out[1] = s->mig_out[1] = cnt[1];
    s->n_gone = cnt[0] + cnt[1];
    return 0;
}
int rods_slab_migration_import(crb_sim* s, const void* lo_buf, const void* hi_buf, uint64_t n_in[2]) {
    u64 cnt[2] = {0, 0}, outc[2] = {0, 0};
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
            return fail(s, CRB_ERR_CAPACITY, "more migrating rods than migration_capacity");
        s->mig_out[0] = outc[0];
        s->mig_out[1] = outc[1];
        s->n_gone = outc[0] + outc[1];
    }
    if (cnt[0] > s->mcap || cnt[1] > s->mcap)
        return fail(s, CRB_ERR_CAPACITY, "more arriving rods than migration_capacity");
    if (s->G + s->n_src + cnt[0] + cnt[1] > s->cap_total - s->G)
        return fail(s, CRB_ERR_CAPACITY, "more rods than capacity");
    stage_begin(s, 4);
    int nl = 0;
    for (int side = 0; side < 2; ++side) {
        if (!cnt[side]) continue;
        RodMigParams M = rods_mig_params(s);
        M.first = s->G + s->n_src;
        M.n = cnt[side];
        M.buf[0] = (unsigned char*)lo_buf;
        M.buf[1] = (unsigned char*)hi_buf;
        k_rod_mig_import<<<blocks_for(cnt[side], 256), 256, 0, s->stream>>>(M, side);
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

int rods_get_agents(crb_sim* s, u64 cap, double* pos, double* vel, u64* ids, u64* voxel) {
    const u64 n = s->n;
    const int W = 3 * s->V;
    int rc;
    if (pos && (rc = rods_download_planes(s, pos, W, s->RX.pos, n))) return rc;
    if (vel && (rc = rods_download_planes(s, vel, W, s->RX.vel, n))) return rc;
    if (ids)
        CU(s, cudaMemcpyAsync(ids, s->RX.ident + s->cap_total + s->G, n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    if (voxel && n) {
        const u64 base = (u64)s->g.p0 * s->g.ng[1] * s->g.ng[2];
        k_global_keys<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->key + s->G, (u64*)s->stage, base, n);
        s->launches++;
        CU(s, cudaMemcpyAsync(voxel, s->stage, n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    }
    CU(s, cudaStreamSynchronize(s->stream));
    return 0;
}

int rods_get_identifiers(crb_sim* s, u64 cap, u64* ident, u64* parent) {
    const u64 n = s->n;
    if (!n) return 0;
    for (int which = 0; which < 2; ++which) {
        u64* out = which == 0 ? ident : parent;
        if (!out) continue;
        k_rod_pack_u64<<<blocks_for(n, 256), 256, 0, s->stream>>>((u64*)s->stage, (which == 0 ? s->RX.ident : s->RX.parent) + s->G,
                                                                 3, s->cap_total, n);
        s->launches++;
        CU(s, cudaMemcpyAsync(out, s->stage, n * 3 * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
        CU(s, cudaStreamSynchronize(s->stream));
    }
    return 0;
}

int rods_get_params(crb_sim* s, u64 cap, double* mech_params, double* inter_params, u8* species) {
    const u64 n = s->n;
    if (mech_params) {
        std::vector<double> len(n), growth(n);
        CU(s, cudaMemcpyAsync(len.data(), s->RX.len + s->G, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(s, cudaMemcpyAsync(growth.data(), s->RX.growth + s->G, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(s, cudaStreamSynchronize(s->stream));
        for (u64 i = 0; i < n; ++i) {
            for (int k = 0; k < 8; ++k) mech_params[i * 8 + k] = s->uni.mp[k];
            mech_params[i * 8 + RMP_LENGTH] = len[i];
            mech_params[i * 8 + RMP_GROWTH] = growth[i];
        }
    }
    if (inter_params)
        for (u64 i = 0; i < n; ++i)
            for (int k = 0; k < 4; ++k) inter_params[i * 4 + k] = s->uni.ip[k];
    if (species) std::memset(species, 0, n);
    return 0;
}

// ring buffers oldest-first: [n][2][V][3]
int rods_get_aux(crb_sim* s, u64 cap, double* hist_dx, double* hist_dv, uint32_t* hist_len) {
    const u64 n = s->n;
    const int W = 3 * s->V;
    std::vector<u8> hl(n);
    CU(s, cudaMemcpyAsync(hl.data(), s->RX.hlen + s->G, n, cudaMemcpyDeviceToHost, s->stream));
    CU(s, cudaStreamSynchronize(s->stream));
    std::vector<double> older(n * W), newest(n * W);
    int rc;
    for (int which = 0; which < 2; ++which) {
        double* outp = which == 0 ? hist_dx : hist_dv;
        if (!outp) continue;
        if ((rc = rods_download_planes(s, older.data(), W, which == 0 ? s->RX.h0x : s->RX.h0v, n))) return rc;
        if ((rc = rods_download_planes(s, newest.data(), W, which == 0 ? s->RX.h1x : s->RX.h1v, n))) return rc;
        for (u64 i = 0; i < n; ++i) {
            double* o = outp + i * 2 * W;
            for (int k = 0; k < 2 * W; ++k) o[k] = 0.0;
            if (hl[i] == 2) {
                for (int k = 0; k < W; ++k) {
                    o[k] = older[i * W + k];
                    o[W + k] = newest[i * W + k];
                }
            } else if (hl[i] == 1) {
                for (int k = 0; k < W; ++k) o[k] = newest[i * W + k];
            }
        }
    }
    if (hist_len)
        for (u64 i = 0; i < n; ++i) hist_len[i] = hl[i];
    return 0;
}

int rods_debug_forces(crb_sim* s, double* force, u64* counts) {
    int rc;
    if (!s->force_out)
        if ((rc = dalloc(s, &s->force_out, s->cap_total * 3 * s->V))) return rc;
    if (!s->nb_out)
        if ((rc = dalloc(s, &s->nb_out, s->cap_total))) return rc;
    RodParams P = rods_params(s);
    rods_launch_force(s, P, true);
    s->launches += s->n ? 1 : 0;
    CU(s, cudaGetLastError());
    if (force && (rc = rods_download_planes(s, force, 3 * s->V, s->force_out, s->n))) return rc;
    if (counts)
        CU(s, cudaMemcpyAsync(counts, s->nb_out + s->G, s->n * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    return check_latch(s);
}

}  // namespace
