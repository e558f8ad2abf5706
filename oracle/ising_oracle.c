/*
 * ising_oracle.c — CPU restatement of the IsingModel.jl spin-update hot path (see ising_oracle.h for
 * the contract; TEST INFRASTRUCTURE ONLY, never loaded by the product).
 *
 * Build:  gcc -O3 -ffp-contract=off -pthread -shared -fPIC ising_oracle.c -o liboracle.so
 * Every function cites the reference lines (path under /root/reference) it follows.
 */
#include "ising_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define JAT(J, ld, i, j) ((J)[(int64_t)(i) + (int64_t)(j) * (ld)])

/* src/SpinSystems.jl:163-171 — H(x)=1 for x>0, 0 for x<0, c=1 at x==0. */
double orc_heaviside(double x) {
    if (x > 0.0) return 1.0;
    if (x < 0.0) return 0.0;
    return 1.0;
}

/* src/SpinSystems.jl:80-83 — couplingCoefficients[i, :]' * s + h[i].
 * The row is summed sequentially over ascending j (the reference's order depends on the storage
 * type of J: generic dot, BLAS ddot or sparse dot; sequential is the generic one). */
double orc_local_field_site(int n, const double *J, int64_t ld, const double *h, const int8_t *s, int i) {
    double acc = 0.0;
    for (int j = 0; j < n; ++j) acc += JAT(J, ld, i, j) * (double)s[j];
    return acc + h[i];
}

/* src/SpinSystems.jl:75-78 — J*s + h, each component a sequential row sum. */
void orc_local_field(int n, const double *J, int64_t ld, const double *h, const int8_t *s, double *out) {
    for (int i = 0; i < n; ++i) out[i] = orc_local_field_site(n, J, ld, h, s, i);
}

/* src/SpinSystems.jl:68-71 — -0.5 * s' * J * s - h' * s.  Evaluated as
 * -0.5 * sum_i s_i (sum_j J_ij s_j)  -  sum_i h_i s_i  (ascending i, j). */
double orc_energy(int n, const double *J, int64_t ld, const double *h, const int8_t *s) {
    double quad = 0.0, lin = 0.0;
    for (int i = 0; i < n; ++i) {
        double row = 0.0;
        for (int j = 0; j < n; ++j) row += JAT(J, ld, i, j) * (double)s[j];
        quad += (double)s[i] * row;
        lin += h[i] * (double)s[i];
    }
    return -0.5 * quad - lin;
}

/* src/SingleSpinFlip.jl:31-36 (Hopfield: J_i.s - h_i, note the MINUS),
 *                      :46-55 (Glauber:  2*h_loc - f*T),
 *                      :65-74 (Metropolis: 2*h_loc - f*T*s_i, evaluated (f*T)*s_i). */
int orc_ssf_update(int rule, int n, const double *J, int64_t ld, const double *h, int8_t *s,
                   int node, double fluct, double T) {
    double x;
    if (rule == ORC_HOPFIELD) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += JAT(J, ld, node, j) * (double)s[j];
        x = acc - h[node];
    } else {
        double hloc = orc_local_field_site(n, J, ld, h, s, node);
        double two_h = 2.0 * hloc;
        double ft = fluct * T;
        if (rule == ORC_METROPOLIS) ft = ft * (double)s[node];
        x = two_h - ft;
    }
    int v = (int)(2.0 * orc_heaviside(x) - 1.0);
    s[node] = (int8_t)v;
    return v;
}

static double magnetization(int n, const int8_t *s) {
    double m = 0.0;
    for (int i = 0; i < n; ++i) m += (double)s[i];
    return m;
}

/* src/SamplingHelper.jl:45-49 — temperature for step k is set BEFORE the update of step k. */
int64_t orc_ssf_run(int rule, int n, const double *J, int64_t ld, const double *h, int8_t *s,
                    int64_t nsteps, const int32_t *nodes, int start, const double *fluct,
                    const double *Tsched, int64_t steps_per_T,
                    int64_t trace_every, double *out_E, double *out_M) {
    int64_t flips = 0, ntr = 0;
    if (steps_per_T < 1) steps_per_T = 1;
    for (int64_t k = 0; k < nsteps; ++k) {
        int node = nodes ? nodes[k] : (int)(((int64_t)start + k) % n);
        double T = Tsched ? Tsched[k / steps_per_T] : 0.0;
        double f = fluct ? fluct[k] : 0.0;
        int8_t old = s[node];
        int v = orc_ssf_update(rule, n, J, ld, h, s, node, f, T);
        if (v != old) ++flips;
        if (trace_every > 0 && (k + 1) % trace_every == 0) {
            if (out_E) out_E[ntr] = orc_energy(n, J, ld, h, s);
            if (out_M) out_M[ntr] = magnetization(n, s);
            ++ntr;
        }
    }
    return flips;
}

int orc_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* Minimal pthread parallel-for (one chain per work item, dynamic distribution). */
typedef void (*orc_item_fn)(int item, void *arg);
typedef struct { orc_item_fn fn; void *arg; int nitems; int next; pthread_mutex_t mu; } orc_pool;
static void *orc_worker(void *p) {
    orc_pool *pool = (orc_pool *)p;
    for (;;) {
        pthread_mutex_lock(&pool->mu);
        int it = pool->next++;
        pthread_mutex_unlock(&pool->mu);
        if (it >= pool->nitems) break;
        pool->fn(it, pool->arg);
    }
    return NULL;
}
static void orc_parallel_for(int nitems, int nthreads, orc_item_fn fn, void *arg) {
    if (nthreads > nitems) nthreads = nitems;
    if (nthreads <= 1) { for (int i = 0; i < nitems; ++i) fn(i, arg); return; }
    orc_pool pool; pool.fn = fn; pool.arg = arg; pool.nitems = nitems; pool.next = 0;
    pthread_mutex_init(&pool.mu, NULL);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, orc_worker, &pool);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&pool.mu);
}

typedef struct {
    int rule, n; const double *J; int64_t ld; const double *h; int8_t *s; int64_t lds;
    int64_t nsteps; const int32_t *nodes; int start; const double *fluct; int per_replica;
    const double *Tsched; int64_t steps_per_T; int64_t *flips;
} orc_ssf_job;
static void orc_ssf_item(int r, void *arg) {
    orc_ssf_job *j = (orc_ssf_job *)arg;
    const double *f = j->fluct ? (j->per_replica ? j->fluct + (int64_t)r * j->nsteps : j->fluct) : NULL;
    j->flips[r] = orc_ssf_run(j->rule, j->n, j->J, j->ld, j->h, j->s + (int64_t)r * j->lds, j->nsteps,
                              j->nodes, j->start, f, j->Tsched, j->steps_per_T, 0, NULL, NULL);
}

int64_t orc_ssf_run_batch(int rule, int n, const double *J, int64_t ld, const double *h,
                          int R, int8_t *s, int64_t lds,
                          int64_t nsteps, const int32_t *nodes, int start,
                          const double *fluct, int fluct_per_replica,
                          const double *Tsched, int64_t steps_per_T, int nthreads) {
    int64_t total = 0;
    if (nthreads < 1) nthreads = 1;
    int64_t *flips = (int64_t *)calloc((size_t)(R > 0 ? R : 1), sizeof(int64_t));
    orc_ssf_job job = {rule, n, J, ld, h, s, lds, nsteps, nodes, start, fluct, fluct_per_replica,
                       Tsched, steps_per_T, flips};
    orc_parallel_for(R, nthreads, orc_ssf_item, &job);
    for (int r = 0; r < R; ++r) total += flips[r];
    free(flips);
    return total;
}

/* ---------------------------------------------------------------- bipartite ----------------- */

/* src/SpinSystems.jl:147-150 — W*tau + h (row sums over ascending hidden index). */
void orc_bip_local_field(int nv, int nh, const double *W, int64_t ld, const double *h,
                         const int8_t *tau, double *out) {
    /* Column sweep (the loop nest of a column-major gemv, as Julia's W * tau runs it): every out[i] still
     * receives its terms in ascending j, so the result is bit-identical to the row-by-row sum. */
    for (int i = 0; i < nv; ++i) out[i] = 0.0;
    for (int j = 0; j < nh; ++j) {
        const double t = (double)tau[j];
        const double *col = W + (int64_t)j * ld;
        for (int i = 0; i < nv; ++i) out[i] += col[i] * t;
    }
    for (int i = 0; i < nv; ++i) out[i] = out[i] + h[i];
}

/* src/SpinSystems.jl:154-157 — W'*sigma + b (column sums over ascending visible index). */
void orc_bip_aux_bias(int nv, int nh, const double *W, int64_t ld, const double *b,
                      const int8_t *sigma, double *out) {
    for (int j = 0; j < nh; ++j) {
        double acc = 0.0;
        for (int i = 0; i < nv; ++i) acc += JAT(W, ld, i, j) * (double)sigma[i];
        out[j] = acc + b[j];
    }
}

/* src/SpinSystems.jl:139-143 — -sigma'*W*tau - h'*sigma - b'*tau. */
double orc_bip_energy(int nv, int nh, const double *W, int64_t ld, const double *h, const double *b,
                      const int8_t *sigma, const int8_t *tau) {
    double quad = 0.0, lv = 0.0, lh = 0.0;
    for (int i = 0; i < nv; ++i) {
        double row = 0.0;
        for (int j = 0; j < nh; ++j) row += JAT(W, ld, i, j) * (double)tau[j];
        quad += (double)sigma[i] * row;
        lv += h[i] * (double)sigma[i];
    }
    for (int j = 0; j < nh; ++j) lh += b[j] * (double)tau[j];
    return -quad - lv - lh;
}

/* src/OnBipartiteGraph.jl:30-43 (SCA) / :53-66 (MomentumAnnealing).
 * Hidden layer first from the OLD visible layer, then visible from the NEW hidden layer.
 * MA multiplies the noise term by the unit's own previous value: (F*T) .* old. */
void orc_bip_update(int rule, int nv, int nh, const double *W, int64_t ld, const double *h,
                    const double *b, int8_t *sigma, int8_t *tau,
                    const double *Fv, const double *Fh, double T) {
    double *aux = (double *)malloc(sizeof(double) * (size_t)(nv > nh ? nv : nh));
    orc_bip_aux_bias(nv, nh, W, ld, b, sigma, aux);
    for (int j = 0; j < nh; ++j) {
        double ft = Fh[j] * T;
        if (rule == ORC_MA) ft = ft * (double)tau[j];
        double x = 2.0 * aux[j] - ft;
        tau[j] = (int8_t)(int)(2.0 * orc_heaviside(x) - 1.0);
    }
    orc_bip_local_field(nv, nh, W, ld, h, tau, aux);
    for (int i = 0; i < nv; ++i) {
        double ft = Fv[i] * T;
        if (rule == ORC_MA) ft = ft * (double)sigma[i];
        double x = 2.0 * aux[i] - ft;
        sigma[i] = (int8_t)(int)(2.0 * orc_heaviside(x) - 1.0);
    }
    free(aux);
}

/* src/SamplingHelper.jl:127-131. */
void orc_bip_run(int rule, int nv, int nh, const double *W, int64_t ld, const double *h,
                 const double *b, int8_t *sigma, int8_t *tau, int64_t nsteps,
                 const double *Fv, const double *Fh, const double *Tsched, int64_t steps_per_T,
                 double *out_E) {
    if (steps_per_T < 1) steps_per_T = 1;
    for (int64_t k = 0; k < nsteps; ++k) {
        double T = Tsched[k / steps_per_T];
        orc_bip_update(rule, nv, nh, W, ld, h, b, sigma, tau, Fv + k * nv, Fh + k * nh, T);
        if (out_E) out_E[k] = orc_bip_energy(nv, nh, W, ld, h, b, sigma, tau);
    }
}

typedef struct {
    int rule, nv, nh; const double *W; int64_t ld; const double *h, *b; int8_t *sigma; int64_t ldsig;
    int8_t *tau; int64_t ldtau; int64_t nsteps; const double *Fv, *Fh; int per_replica;
    const double *Tsched; int64_t steps_per_T;
} orc_bip_job;
static void orc_bip_item(int r, void *arg) {
    orc_bip_job *j = (orc_bip_job *)arg;
    const double *fv = j->per_replica ? j->Fv + (int64_t)r * j->nsteps * j->nv : j->Fv;
    const double *fh = j->per_replica ? j->Fh + (int64_t)r * j->nsteps * j->nh : j->Fh;
    orc_bip_run(j->rule, j->nv, j->nh, j->W, j->ld, j->h, j->b, j->sigma + (int64_t)r * j->ldsig,
                j->tau + (int64_t)r * j->ldtau, j->nsteps, fv, fh, j->Tsched, j->steps_per_T, NULL);
}

void orc_bip_run_batch(int rule, int nv, int nh, const double *W, int64_t ld, const double *h,
                       const double *b, int R, int8_t *sigma, int64_t ldsig, int8_t *tau,
                       int64_t ldtau, int64_t nsteps, const double *Fv, const double *Fh,
                       int fluct_per_replica, const double *Tsched, int64_t steps_per_T,
                       int nthreads) {
    if (nthreads < 1) nthreads = 1;
    orc_bip_job job = {rule, nv, nh, W, ld, h, b, sigma, ldsig, tau, ldtau, nsteps, Fv, Fh,
                       fluct_per_replica, Tsched, steps_per_T};
    orc_parallel_for(R, nthreads, orc_bip_item, &job);
}

/* ---------------------------------------------------------------- Philox4x32-10 ------------- */
/* Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11).
 * Multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85, 10 rounds. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
