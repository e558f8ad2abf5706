/*
 * ising_oracle.h — CPU restatement (Float64, scalar) of the spin-update hot path of
 * Wandao123/IsingModel.jl.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product (libising_b200.so) never links, loads or calls it.
 *
 * Parity status: the reference holds exactly seven fixtures for this path (test/runtests.jl:20-24,28-31:
 * two energy known-answer tests and five end-state membership checks); the oracle is checked against
 * all seven in tests/test_oracle.py.  The reference holds NO trajectory, fluctuation-stream or
 * field-value fixture, and Julia is not available in this image, so trajectory-level parity is
 * UNPINNED beyond the in-tree arithmetic at the 3-argument update! boundary that is restated here
 * line by line (see the file:line citations on each function).
 *
 * Conventions: matrices are column-major with a leading dimension (Julia layout); site indices are
 * 0-based; spins are int8 (+1/-1); all arithmetic is IEEE double, no FMA contraction
 * (-ffp-contract=off), summations run sequentially in ascending index order.
 */
#ifndef ISING_ORACLE_H
#define ISING_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_HOPFIELD = 0, ORC_GLAUBER = 1, ORC_METROPOLIS = 2 };
enum { ORC_SCA = 0, ORC_MA = 1 };

/* SpinSystems.jl:163-171 */
double orc_heaviside(double x);

/* SpinSystems.jl:68-71   E = -1/2 s'Js - h's */
double orc_energy(int n, const double *J, int64_t ld, const double *h, const int8_t *s);
/* SpinSystems.jl:80-83   J[i,:]'s + h[i] */
double orc_local_field_site(int n, const double *J, int64_t ld, const double *h, const int8_t *s, int i);
/* SpinSystems.jl:75-78   J s + h */
void orc_local_field(int n, const double *J, int64_t ld, const double *h, const int8_t *s, double *out);

/* SingleSpinFlip.jl:31-36,46-55,65-74 — one 3-argument update!; returns the new spin. */
int orc_ssf_update(int rule, int n, const double *J, int64_t ld, const double *h, int8_t *s,
                   int node, double fluct, double T);

/*
 * SamplingHelper.jl:45-49 step loop for one chain: for k in 0..nsteps-1:
 *   T <- Tsched[k / steps_per_T]; update!(ua, nodes[k], fluct[k]).
 * nodes == NULL means the sequential order node = (start + k) % n.
 * trace_every > 0: after every trace_every-th step write E and M (sum of spins) into out_E/out_M.
 * Returns the number of steps that changed the spin.
 */
int64_t orc_ssf_run(int rule, int n, const double *J, int64_t ld, const double *h, int8_t *s,
                    int64_t nsteps, const int32_t *nodes, int start, const double *fluct,
                    const double *Tsched, int64_t steps_per_T,
                    int64_t trace_every, double *out_E, double *out_M);

/* R independent chains (replica-major spins [R][lds], fluct either shared [nsteps] or per replica
 * [R][nsteps]); nthreads pthreads, one chain per thread at a time. Returns total flips. */
int64_t orc_ssf_run_batch(int rule, int n, const double *J, int64_t ld, const double *h,
                          int R, int8_t *s, int64_t lds,
                          int64_t nsteps, const int32_t *nodes, int start,
                          const double *fluct, int fluct_per_replica,
                          const double *Tsched, int64_t steps_per_T, int nthreads);

/* SpinSystems.jl:139-143   E = -sigma' W tau - h' sigma - b' tau */
double orc_bip_energy(int nv, int nh, const double *W, int64_t ld, const double *h, const double *b,
                      const int8_t *sigma, const int8_t *tau);
/* SpinSystems.jl:147-150   W tau + h */
void orc_bip_local_field(int nv, int nh, const double *W, int64_t ld, const double *h,
                         const int8_t *tau, double *out);
/* SpinSystems.jl:154-157   W' sigma + b */
void orc_bip_aux_bias(int nv, int nh, const double *W, int64_t ld, const double *b,
                      const int8_t *sigma, double *out);
/* OnBipartiteGraph.jl:30-43 (SCA) and :53-66 (MomentumAnnealing): one update!. */
void orc_bip_update(int rule, int nv, int nh, const double *W, int64_t ld, const double *h,
                    const double *b, int8_t *sigma, int8_t *tau,
                    const double *Fv, const double *Fh, double T);
/* SamplingHelper.jl:127-131 loop: step k uses Fv[:,k] (column k of an nv x nsteps column-major
 * array with leading dimension nv), Fh[:,k], T = Tsched[k / steps_per_T]. out_E (may be NULL)
 * receives the energy after every step. */
void orc_bip_run(int rule, int nv, int nh, const double *W, int64_t ld, const double *h,
                 const double *b, int8_t *sigma, int8_t *tau, int64_t nsteps,
                 const double *Fv, const double *Fh, const double *Tsched, int64_t steps_per_T,
                 double *out_E);
/* R chains; Fv/Fh either shared or per replica ([R][nsteps][nv] i.e. replica-major blocks). */
void orc_bip_run_batch(int rule, int nv, int nh, const double *W, int64_t ld, const double *h,
                       const double *b, int R, int8_t *sigma, int64_t ldsig, int8_t *tau,
                       int64_t ldtau, int64_t nsteps, const double *Fv, const double *Fh,
                       int fluct_per_replica, const double *Tsched, int64_t steps_per_T,
                       int nthreads);

/* Philox4x32-10 (Salmon et al., SC'11), used by the product's counter RNG; restated here so the
 * tests can check the device generator word for word. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
