/* c_harness.c — the drop-in boundary driven from a plain C host (no Python, no torch).
 *
 * Compiled by tests/test_gpu_c_harness.py with `gcc -std=c11 -I include tests/c_harness.c -L cellular_raza_b200/lib
 * -lcrb200` and run on the GPU box: crb_create -> crb_set_agents -> crb_step -> crb_get_agents ->
 * crb_snapshot_async / crb_snapshot_wait -> crb_destroy, plus the error convention (status code + crb_last_error).
 * This is the call sequence the generated main loop of the reference's run_simulation! (run_sim.rs:380-403) maps to
 * (INTEGRATION.md).  The positions come from a fixed LCG so that the Python test can upload the same population
 * through ctypes and compare the printed hexadecimal checksums bit for bit.
 */
#include <inttypes.h>
#include <math.h>
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

static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static double lcg_uniform(void) { /* 53 random bits in (0, 1) */
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    return ((double)(lcg_state >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 2000;
    const uint64_t steps = argc > 2 ? strtoull(argv[2], NULL, 10) : 10;
    const double L = 72.0;
    crb_sim* sim = NULL;

    if (crb_abi_version() != CRB_ABI_VERSION) {
        fprintf(stderr, "ABI mismatch: header %d, library %u\n", CRB_ABI_VERSION, crb_abi_version());
        return 1;
    }
    crb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = CRB_ABI_VERSION;
    cfg.dim = 3;
    cfg.mechanics_kind = CRB_MECH_NEWTON_DAMPED;
    cfg.interaction_kind = CRB_INT_CELL_SORTING;
    cfg.solver_order = 2;
    cfg.pair_weight_mode = CRB_WEIGHT_CHILI;
    cfg.ref_n_subdomains = 4; /* Settings { n_threads: 4 } of the shipped example */
    cfg.noise_mode = CRB_NOISE_NONE;
    cfg.world_size = 1;
    double mn[3] = {0, 0, 0}, mx[3] = {L, L, L};
    /* CartesianCuboid::from_boundaries_and_interaction_range(min, max, 18.0) */
    if (crb_domain_from_interaction_range(3, mn, mx, 18.0, cfg.n_voxels, cfg.dx) != 0) return 1;
    for (int i = 0; i < 3; ++i) {
        cfg.domain_min[i] = mn[i];
        cfg.domain_max[i] = mx[i];
    }
    cfg.dt = 0.2;
    cfg.capacity = n;

    /* error convention before anything else: a step without agents is a call-sequence violation */
    if (crb_create(&cfg, &sim) != 0) {
        fprintf(stderr, "crb_create: %s\n", crb_last_error(NULL));
        return 1;
    }
    if (crb_step(sim, 1) != CRB_ERR_STATE || strlen(crb_last_error(sim)) == 0) {
        fprintf(stderr, "expected CRB_ERR_STATE with a message\n");
        return 1;
    }

    double* pos = malloc(n * 3 * sizeof(double));
    double* vel = calloc(n * 3, sizeof(double));
    double* mp = malloc(n * 2 * sizeof(double));
    double* ip = malloc(n * 3 * sizeof(double));
    uint8_t* species = malloc(n);
    for (uint64_t i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) pos[i * 3 + d] = lcg_uniform() * L;
        mp[i * 2 + 0] = 2.0; /* damping_constant */
        mp[i * 2 + 1] = 1.0; /* mass */
        ip[i * 3 + 0] = 6.0; /* cell_radius */
        ip[i * 3 + 1] = 2.0; /* potential_strength */
        ip[i * 3 + 2] = 1.5; /* relative_interaction_range */
        species[i] = i <= n / 2 ? 1 : 0;
    }
    CHECK(crb_set_agents(sim, n, pos, vel, mp, ip, species, NULL));
    if (crb_n_agents(sim) != n) return 1;
    CHECK(crb_step(sim, steps));
    if (crb_iteration(sim) != steps) return 1;

    double* p1 = malloc(n * 3 * sizeof(double));
    double* v1 = malloc(n * 3 * sizeof(double));
    uint64_t* ids = malloc(n * sizeof(uint64_t));
    uint64_t* vox = malloc(n * sizeof(uint64_t));
    CHECK(crb_get_agents(sim, n, p1, v1, ids, vox));
    uint64_t n_pairs = 0;
    CHECK(crb_count_candidate_pairs(sim, &n_pairs));

    /* storage snapshot on the side stream into page-locked buffers, while two more steps are queued */
    double* sp = crb_host_alloc(n * 3 * sizeof(double));
    double* sv = crb_host_alloc(n * 3 * sizeof(double));
    uint64_t* sid = crb_host_alloc(n * sizeof(uint64_t));
    double* shx = crb_host_alloc(n * 6 * sizeof(double));
    double* shv = crb_host_alloc(n * 6 * sizeof(double));
    uint32_t* shl = crb_host_alloc(n * sizeof(uint32_t));
    if (!sp || !sv || !sid || !shx || !shv || !shl) return 1;
    CHECK(crb_snapshot_async(sim, n, sp, sv, sid, NULL, shx, shv, shl));
    CHECK(crb_step(sim, 2));
    uint64_t n_snap = 0;
    CHECK(crb_snapshot_wait(sim, &n_snap));
    if (n_snap != n || memcmp(sp, p1, n * 3 * sizeof(double)) || memcmp(sv, v1, n * 3 * sizeof(double)) ||
        memcmp(sid, ids, n * sizeof(uint64_t))) {
        fprintf(stderr, "snapshot differs from crb_get_agents\n");
        return 1;
    }

    /* order-independent checksum of the state after `steps` steps: sums over agents in identifier order */
    double* by_id = malloc(n * 3 * sizeof(double));
    uint64_t seen = 0, max_hl = 0;
    for (uint64_t a = 0; a < n; ++a) {
        if (ids[a] >= n) return 1;
        memcpy(by_id + ids[a] * 3, p1 + a * 3, 3 * sizeof(double));
        seen += 1;
        if (shl[a] > max_hl) max_hl = shl[a];
        for (int d = 0; d < 3; ++d)
            if (!(p1[a * 3 + d] >= 0.0 && p1[a * 3 + d] <= L)) {
                fprintf(stderr, "agent %" PRIu64 " left the domain\n", ids[a]);
                return 1;
            }
        const uint64_t v = vox[a];
        const uint64_t i0 = v / 16, i1 = (v / 4) % 4, i2 = v % 4;
        if ((uint64_t)(p1[a * 3] / cfg.dx[0]) != i0 || (uint64_t)(p1[a * 3 + 1] / cfg.dx[1]) != i1 ||
            (uint64_t)(p1[a * 3 + 2] / cfg.dx[2]) != i2) {
            fprintf(stderr, "voxel index of agent %" PRIu64 " does not match its position\n", ids[a]);
            return 1;
        }
    }
    double sx = 0, sy = 0, sz = 0;
    for (uint64_t i = 0; i < n; ++i) {
        sx += by_id[i * 3];
        sy += by_id[i * 3 + 1];
        sz += by_id[i * 3 + 2];
    }
    printf("C_HARNESS n=%" PRIu64 " steps=%" PRIu64 " ring=%" PRIu64 " pairs=%" PRIu64 " launches=%" PRIu64
           " kernel=\"%s\" checksum=%a,%a,%a\n",
           seen, steps, max_hl, n_pairs, crb_kernel_launches(sim), crb_force_kernel_name(sim), sx, sy, sz);

    crb_host_free(sp);
    crb_host_free(sv);
    crb_host_free(sid);
    crb_host_free(shx);
    crb_host_free(shv);
    crb_host_free(shl);
    crb_destroy(sim);
    free(pos);
    free(vel);
    free(mp);
    free(ip);
    free(species);
    free(p1);
    free(v1);
    free(ids);
    free(vox);
    free(by_id);
    printf("C_HARNESS OK\n");
    return 0;
}
