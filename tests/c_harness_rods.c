/* c_harness_rods.c — the bacterial-rods agent and the ASYNCHRONOUS calls of the boundary from a plain C host.
 *
 * Compiled and run by tests/test_gpu_c_harness.py like c_harness.c.  Call sequence: crb_create (CRB_MECH_ROD) ->
 * crb_set_agents_async with page-locked rows -> crb_step -> crb_snapshot_async / crb_snapshot_wait -> crb_get_agents
 * (must hold the same bytes) -> crb_get_identifiers; then the device-side check of the parameter rows: a row that
 * differs in a field other than spring_length / growth_rate is reported by the next crb_step as CRB_ERR_INVALID_ARG.
 * The rods (agent of cellular_raza-examples/bacterial_rods/src/main.rs:111-212, parameters of :120-138) are laid out
 * with + and * only from a fixed LCG, so that the Python twin builds the same bits without a libm in between.
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crb.h"

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc__ = (call);                                                                  \
        if (rc__ != 0) {                                                                    \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc__, crb_last_error(sim));      \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

#define V 5

static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static double lcg_uniform(void) {
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    return ((double)(lcg_state >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 48;
    const uint64_t steps = argc > 2 ? strtoull(argv[2], NULL, 10) : 12;
    const double um = 1e-6;
    crb_sim* sim = NULL;

    crb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = CRB_ABI_VERSION;
    cfg.dim = 3;
    cfg.mechanics_kind = CRB_MECH_ROD;
    cfg.interaction_kind = CRB_INT_MORSE;
    cfg.rod_n_vertices = V;
    cfg.solver_order = 2;
    cfg.pair_weight_mode = CRB_WEIGHT_CHILI;
    cfg.ref_n_subdomains = 8; /* Settings { n_threads: 8 } of the example */
    cfg.noise_mode = CRB_NOISE_NONE;
    cfg.world_size = 1;
    /* two tiles of the shipped domain (200 x 15 x 45 um, [8, 1, 2] voxels) side by side */
    const double mx[3] = {400.0 * um, 15.0 * um, 45.0 * um};
    const uint64_t nv[3] = {16, 1, 2};
    for (int i = 0; i < 3; ++i) {
        cfg.domain_min[i] = 0.0;
        cfg.domain_max[i] = mx[i];
        cfg.n_voxels[i] = nv[i];
        cfg.dx[i] = mx[i] / (double)nv[i];
    }
    cfg.dt = 0.1;
    cfg.capacity = 4 * n;
    const uint64_t cap = 4 * n;
    if (crb_create(&cfg, &sim) != 0) {
        fprintf(stderr, "crb_create: %s\n", crb_last_error(NULL));
        return 1;
    }

    double* pos = crb_host_alloc(n * V * 3 * sizeof(double));
    double* vel = crb_host_alloc(n * V * 3 * sizeof(double));
    double* mp = crb_host_alloc(n * 8 * sizeof(double));
    double* ip = crb_host_alloc(n * 4 * sizeof(double));
    if (!pos || !vel || !mp || !ip) return 1;
    memset(vel, 0, n * V * 3 * sizeof(double));
    for (uint64_t i = 0; i < n; ++i) {
        const double x0 = (20.0 + 340.0 * lcg_uniform()) * um;
        const double y0 = (5.0 + 5.0 * lcg_uniform()) * um;
        const double z0 = (10.0 + 25.0 * lcg_uniform()) * um;
        for (int v = 0; v < V; ++v) {
            pos[(i * V + v) * 3 + 0] = x0 + 3.0 * um * (double)v;
            pos[(i * V + v) * 3 + 1] = y0 + 0.2 * um * (lcg_uniform() - 0.5);
            pos[(i * V + v) * 3 + 2] = z0;
        }
        double* m = mp + i * 8;
        m[0] = 0.0;                                    /* diffusion_constant */
        m[1] = 10.0;                                   /* spring_tension */
        m[2] = 1.0 * um;                               /* rigidity */
        m[3] = (5.9 + 0.09 * lcg_uniform()) * um;      /* spring_length: divisions within a few steps */
        m[4] = 1.5;                                    /* damping */
        m[5] = 0.1 * um;                               /* growth_rate */
        m[6] = 6.0 * um;                               /* spring_length_threshold */
        m[7] = 0.1 * um;                               /* growth_base_rate */
        double* q = ip + i * 4;
        q[0] = 3.0 * um;       /* radius */
        q[1] = 0.5 / um;       /* potential_stiffness */
        q[2] = 5.0 * um;       /* cutoff */
        q[3] = 0.1 * um * um;  /* strength */
    }
    CHECK(crb_set_agents_async(sim, n, pos, vel, mp, ip, NULL, NULL));
    CHECK(crb_step(sim, steps)); /* a rod step synchronises: latched ingestion errors would surface here */
    const uint64_t n1 = crb_n_agents(sim);
    if (n1 <= n || n1 > cap) {
        fprintf(stderr, "expected divisions: %" PRIu64 " -> %" PRIu64 " rods\n", n, n1);
        return 1;
    }

    double* sp = crb_host_alloc(cap * V * 3 * sizeof(double));
    double* sv = crb_host_alloc(cap * V * 3 * sizeof(double));
    uint64_t* sid = crb_host_alloc(cap * sizeof(uint64_t));
    uint64_t* svox = crb_host_alloc(cap * sizeof(uint64_t));
    if (!sp || !sv || !sid || !svox) return 1;
    /* the ring buffers are not part of a rod snapshot */
    if (crb_snapshot_async(sim, cap, sp, sv, sid, svox, sp, NULL, NULL) != CRB_ERR_INVALID_ARG) return 1;
    CHECK(crb_snapshot_async(sim, cap, sp, sv, sid, svox, NULL, NULL, NULL));
    uint64_t n_snap = 0;
    CHECK(crb_snapshot_wait(sim, &n_snap));
    double* p1 = malloc(cap * V * 3 * sizeof(double));
    double* v1 = malloc(cap * V * 3 * sizeof(double));
    uint64_t* vox = malloc(cap * sizeof(uint64_t));
    uint64_t* ident = malloc(cap * 3 * sizeof(uint64_t));
    uint64_t* parent = malloc(cap * 3 * sizeof(uint64_t));
    CHECK(crb_get_agents(sim, cap, p1, v1, NULL, vox));
    CHECK(crb_get_identifiers(sim, cap, ident, parent));
    if (n_snap != n1 || memcmp(sp, p1, n1 * V * 3 * sizeof(double)) || memcmp(sv, v1, n1 * V * 3 * sizeof(double)) ||
        memcmp(svox, vox, n1 * sizeof(uint64_t))) {
        fprintf(stderr, "snapshot differs from crb_get_agents\n");
        return 1;
    }
    uint64_t n_division_ids = 0;
    for (uint64_t a = 0; a < n1; ++a) {
        if (sid[a] != ident[a * 3 + 1]) {
            fprintf(stderr, "snapshot id differs from crb_get_identifiers\n");
            return 1;
        }
        if (ident[a * 3] == 0) ++n_division_ids; /* CellIdentifier::Division(voxel, counter) */
    }
    double sx = 0, sy = 0, sz = 0; /* device order is part of the contract: the twin adds in the same order */
    for (uint64_t a = 0; a < n1 * V; ++a) {
        sx += p1[a * 3];
        sy += p1[a * 3 + 1];
        sz += p1[a * 3 + 2];
    }
    printf("C_HARNESS_RODS n0=%" PRIu64 " n1=%" PRIu64 " steps=%" PRIu64 " division_ids=%" PRIu64 " checksum=%a,%a,%a\n", n,
           n1, steps, n_division_ids, sx, sy, sz);

    /* a parameter row that differs in a uniform field is found on the device and reported by the next crb_step */
    mp[(n / 2) * 8 + 4] = 3.0;
    CHECK(crb_set_agents_async(sim, n, pos, vel, mp, ip, NULL, NULL));
    const int rc = crb_step(sim, 1);
    if (rc != CRB_ERR_INVALID_ARG || !strstr(crb_last_error(sim), "uniform")) {
        fprintf(stderr, "expected CRB_ERR_INVALID_ARG from the device-side row check, got %d (%s)\n", rc,
                crb_last_error(sim));
        return 1;
    }

    crb_host_free(pos);
    crb_host_free(vel);
    crb_host_free(mp);
    crb_host_free(ip);
    crb_host_free(sp);
    crb_host_free(sv);
    crb_host_free(sid);
    crb_host_free(svox);
    crb_destroy(sim);
    free(p1);
    free(v1);
    free(vox);
    free(ident);
    free(parent);
    printf("C_HARNESS_RODS OK\n");
    return 0;
}
