/* c_abi_replay.c — a plain C caller of libising_b200.so: compiled with gcc against include/ising_b200.h and linked to the
 * shared library (no Python, no ctypes, no torch).  Replays one golden single-spin trajectory and one golden block-Gibbs
 * trajectory (the .npz fixtures of tests/golden, flattened to a binary blob by tests/test_c_abi.py) through the entry points the
 * reference-side binding uses and compares spins, flip counts and energies with the stored results.
 *
 *   c_abi_replay <blob>      exit 0: identical; 1: mismatch; 2: usage / file error; 3: no usable CUDA device
 *   c_abi_replay --symbols   exit 0 after taking the address of every entry point it links against (CPU link check)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ising_b200.h"

static void *rd(FILE *f, size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) {
        fprintf(stderr, "c_abi_replay: short read (%zu bytes)\n", bytes);
        exit(2);
    }
    return p;
}

static int fail_rc(isb_ctx *ctx, const char *what, int rc) {
    fprintf(stderr, "c_abi_replay: %s failed (%d): %s\n", what, rc, isb_last_error(ctx));
    return 1;
}
#define CHECK(call)                                    \
    do {                                               \
        int rc_ = (call);                              \
        if (rc_ != ISB_OK) return fail_rc(ctx, #call, rc_); \
    } while (0)

/* blob section 1: int64 header {N, R, nsteps, rule, nT, steps_per_T, trace_every, order_list}, J[N*N] (column-major),
 * h[N], s0[R*N] int8, nodes[nsteps] int32 (order_list only), fluct[R*nsteps], T[nT], then the expected s_final[R*N] int8,
 * flips[R] int64, E[ntr*R] */
static int replay_ssf(isb_ctx *ctx, FILE *f) {
    int64_t *hd = rd(f, 8 * sizeof(int64_t));
    const int N = (int)hd[0], R = (int)hd[1], rule = (int)hd[3];
    const int64_t nsteps = hd[2], nT = hd[4], spT = hd[5], tre = hd[6], list = hd[7];
    const int64_t ntr = tre > 0 ? nsteps / tre : 0;
    double *J = rd(f, sizeof(double) * N * N), *h = rd(f, sizeof(double) * N);
    int8_t *s0 = rd(f, (size_t)R * N);
    int32_t *nodes = list ? rd(f, sizeof(int32_t) * nsteps) : NULL;
    double *fl = rd(f, sizeof(double) * R * nsteps), *T = rd(f, sizeof(double) * nT);
    int8_t *want_s = rd(f, (size_t)R * N);
    int64_t *want_fl = rd(f, sizeof(int64_t) * R);
    double *want_E = rd(f, sizeof(double) * ntr * R);
    isb_model *m = NULL;
    isb_ens *e = NULL, *e2 = NULL;
    int warn = 0;
    CHECK(isb_model_dense(ctx, N, J, N, h, ISB_PREC_F64, &warn, &m));
    CHECK(isb_ens_create(m, R, &e));
    CHECK(isb_ens_set_spins(e, s0, N));
    CHECK(isb_ens_clone(e, &e2)); /* deepcopy(ss): the copy must not move when the original runs */
    int8_t *got = malloc((size_t)R * N), *got2 = malloc((size_t)R * N);
    int64_t *flips = malloc(sizeof(int64_t) * R);
    double *E = malloc(sizeof(double) * (ntr * R + 1));
    CHECK(isb_ssf_run(e, rule, nsteps, list ? ISB_ORDER_LIST : ISB_ORDER_SEQUENTIAL, nodes, 0,
                      R > 1 ? ISB_FLUCT_PER_REPLICA : ISB_FLUCT_SHARED, fl, 0, 0, T, nT, spT, tre, ntr ? E : NULL, NULL, flips));
    CHECK(isb_ens_get_spins(e, got, N));
    CHECK(isb_ens_get_spins(e2, got2, N));
    int bad = memcmp(got, want_s, (size_t)R * N) != 0 || memcmp(got2, s0, (size_t)R * N) != 0;
    for (int r = 0; r < R; ++r) bad |= flips[r] != want_fl[r];
    for (int64_t i = 0; i < ntr * R; ++i) bad |= !(fabs(E[i] - want_E[i]) <= 1e-9 * fmax(1.0, fabs(want_E[i])));
    double kms = 0;
    int64_t launches = 0;
    CHECK(isb_ens_last_stats(e, &kms, &launches, NULL, NULL));
    printf("ssf: N=%d R=%d steps=%lld rule=%d -> %s (%lld launches, %.3f ms)\n", N, R, (long long)nsteps, rule,
           bad ? "MISMATCH" : "identical", (long long)launches, kms);
    isb_ens_destroy(e2);
    isb_ens_destroy(e);
    isb_model_destroy(m);
    return bad;
}

/* blob section 2: int64 header {nv, nh, R, nsteps, rule, nT}, W[nv*nh] (column-major), h[nv], b[nh], s0[R*nv], t0[R*nh],
 * Fv[nsteps*nv], Fh[nsteps*nh], T[nT], then the expected s[R*nv], t[R*nh] */
static int replay_bip(isb_ctx *ctx, FILE *f) {
    int64_t *hd = rd(f, 6 * sizeof(int64_t));
    const int nv = (int)hd[0], nh = (int)hd[1], R = (int)hd[2], rule = (int)hd[4];
    const int64_t nsteps = hd[3], nT = hd[5];
    double *W = rd(f, sizeof(double) * nv * nh), *h = rd(f, sizeof(double) * nv), *b = rd(f, sizeof(double) * nh);
    int8_t *s0 = rd(f, (size_t)R * nv), *t0 = rd(f, (size_t)R * nh);
    double *Fv = rd(f, sizeof(double) * nsteps * nv), *Fh = rd(f, sizeof(double) * nsteps * nh), *T = rd(f, sizeof(double) * nT);
    int8_t *want_s = rd(f, (size_t)R * nv), *want_t = rd(f, (size_t)R * nh);
    isb_model *m = NULL;
    isb_ens *e = NULL;
    CHECK(isb_model_bipartite(ctx, nv, nh, W, nv, h, b, ISB_PREC_F64, &m));
    CHECK(isb_ens_create(m, R, &e));
    CHECK(isb_ens_set_spins(e, s0, nv));
    CHECK(isb_ens_set_hidden(e, t0, nh));
    CHECK(isb_bip_run(e, rule, nsteps, ISB_FLUCT_SHARED, Fv, Fh, 0, 0, T, nT, 1, 0, NULL));
    int8_t *gs = malloc((size_t)R * nv), *gt = malloc((size_t)R * nh);
    CHECK(isb_ens_get_spins(e, gs, nv));
    CHECK(isb_ens_get_hidden(e, gt, nh));
    const int bad = memcmp(gs, want_s, (size_t)R * nv) != 0 || memcmp(gt, want_t, (size_t)R * nh) != 0;
    printf("bip: %d x %d R=%d steps=%lld rule=%d -> %s\n", nv, nh, R, (long long)nsteps, rule, bad ? "MISMATCH" : "identical");
    isb_ens_destroy(e);
    isb_model_destroy(m);
    return bad;
}

int main(int argc, char **argv) {
    if (argc == 2 && strcmp(argv[1], "--symbols") == 0) {
        const void *syms[] = {(void *)isb_version, (void *)isb_create, (void *)isb_destroy, (void *)isb_last_error,
                              (void *)isb_model_dense, (void *)isb_model_sparse, (void *)isb_model_bipartite,
                              (void *)isb_model_retain, (void *)isb_model_destroy, (void *)isb_ens_create, (void *)isb_ens_clone,
                              (void *)isb_ens_destroy, (void *)isb_ens_set_spins, (void *)isb_ens_get_spins,
                              (void *)isb_ens_set_hidden, (void *)isb_ens_get_hidden, (void *)isb_ens_energy,
                              (void *)isb_ssf_run, (void *)isb_ssf_run_snap, (void *)isb_bip_run, (void *)isb_bip_run_snap,
                              (void *)isb_ens_last_stats};
        size_t n = 0;
        for (size_t i = 0; i < sizeof syms / sizeof syms[0]; ++i) n += syms[i] != NULL;
        printf("libising_b200 version %d: %zu entry points linked\n", isb_version(), n);
        return 0;
    }
    if (argc != 2) {
        fprintf(stderr, "usage: c_abi_replay <blob> | --symbols\n");
        return 2;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) {
        perror(argv[1]);
        return 2;
    }
    isb_ctx *ctx = NULL;
    if (isb_create(0, &ctx) != ISB_OK) {
        fprintf(stderr, "c_abi_replay: %s\n", isb_last_error(NULL));
        return 3;
    }
    int bad = replay_ssf(ctx, f);
    bad |= replay_bip(ctx, f);
    isb_destroy(ctx);
    fclose(f);
    printf("%s\n", bad ? "FAILED" : "OK");
    return bad ? 1 : 0;
}
