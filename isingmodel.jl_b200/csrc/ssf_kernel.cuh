// ssf_kernel.cuh — K1: replica-batched single-spin-flip sweeps (Hopfield / Glauber / Metropolis).
//
// Replaces, for R chains at once, the reference's per-step path
//   SamplingHelper.makeSampler! loop            src/SamplingHelper.jl:45-49
//     -> SingleSpinFlip.update!(ua, node, f)    src/SingleSpinFlip.jl:31-36,46-55,65-74
//        -> calcLocalMagneticField(ua, i)       src/SpinSystems.jl:80-83   (an O(N) row dot per step)
//
// Design (B200-first, not a translation):
//   * one warp per chain; the chain's N local fields live in registers (site i <-> lane i%32,
//     register i/32), its spins in one 32-bit mask per lane;
//   * the local field is maintained incrementally: an accepted flip of site i adds +-2*J[i,:] to the
//     fields (J symmetric, src/SpinSystems.jl:31-38), so a row of J is touched only on accepted flips;
//   * every chain of a CTA visits the sites in the same order, so one copy of row i serves all of
//     them: a producer warp streams the rows of J through a shared-memory ring with 1-D bulk async
//     copies (TMA, cp.async.bulk -> UBLKCP) completing on mbarriers, G rows per ring slot;
//   * sequential sweeps are processed 32 sites at a time *speculatively*: all 32 lanes evaluate the
//     decision of their own site at once; a ballot finds the first lane that flips, its row is
//     applied, and only the lanes after it are re-evaluated.  This is exactly the sequential
//     single-site dynamics (a site's decision is taken with the fields left by all earlier sites).
//   * decisions are taken in double with the reference's operation order: 2*h, f*T, (f*T)*s, subtract,
//     compare with heaviside(0)=1 (src/SpinSystems.jl:163-171).
#pragma once
#include <stdio.h>

#include <type_traits>

#include "common.cuh"

namespace isb {

constexpr int SSF_G = 8;  // rows of J per ring slot

struct SsfParams {
    const void *J;  // permuted couplings [npad][ldj] (double or float)
    int64_t ldj;    // row stride in elements
    const double *hext;  // external field [npad]
    int8_t *spins;       // [R][lds]
    int64_t lds;
    void *fields;  // [R][npad] (HT)
    int n, npad, R;
    int rule;
    double ecoef;  // energy = -1/2 sum s*fld - ecoef * sum h*s
    int64_t nsteps;
    int start;
    const int32_t *nodes;
    int fluct_mode;
    const double *fluct;
    uint64_t seed, step_offset;
    const double *Tsched;
    int64_t steps_per_T;
    int64_t trace_every;
    double *out_E, *out_M;
    unsigned long long *flips;
    unsigned long long *near_ties;
    double tie_eps;
    int nw;  // consumer warps (chains) per CTA
    int NG;  // ring slots
};

// consumer warps (chains) per CTA for a chain that keeps `field_regs` 32-bit registers of local
// fields per lane: 15 warps -> 128 registers/thread, 21 -> 96, 29 -> 64 (one CTA per SM, +1 producer warp)
__host__ __device__ constexpr int ssf_max_chains(int field_regs) {
    return field_regs > 32 ? 14 : (field_regs > 16 ? 20 : 28);
}

template <typename HT, int NPL>
struct SsfCfg {
    static constexpr int kFieldRegs = NPL * (int)sizeof(HT) / 4;
    static constexpr int kMaxChains = ssf_max_chains(kFieldRegs);
    static constexpr int kMaxThreads = 32 * (kMaxChains + 1);
};

// compile-time loop: f(integral_constant<int, I>) for I in [0, N) — keeps register-array indices static
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

template <typename JT, int VEC>
struct alignas(sizeof(JT) * VEC) JPack {
    JT v[VEC];
};

// binary select tree: hf[k] with k known only at run time, without dynamic register indexing
template <typename HT, int LO, int CNT>
struct FieldSel {
    static __device__ __forceinline__ HT get(const HT *hf, int k) {
        if constexpr (CNT == 1) {
            return hf[LO];
        } else {
            const HT a = FieldSel<HT, LO, CNT / 2>::get(hf, k);
            const HT b = FieldSel<HT, LO + CNT / 2, CNT / 2>::get(hf, k);
            return (k & (CNT / 2)) ? b : a;
        }
    }
};

template <typename HT, typename JT, int NPL, bool LIST, bool TMA>
__global__ void __launch_bounds__(SsfCfg<HT, NPL>::kMaxThreads, 1) ssf_kernel(const SsfParams p) {
    constexpr int G = SSF_G;
    constexpr int NPAD = NPL * 32;
    constexpr int ROWB = NPAD * (int)sizeof(JT);
    constexpr int VEC = (16 / (int)sizeof(JT)) < NPL ? (16 / (int)sizeof(JT)) : NPL;
    constexpr int NCH = NPL / VEC;
    constexpr uint32_t FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    JT *ring = reinterpret_cast<JT *>(smem_raw);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)p.NG * G * ROWB);
    uint64_t *empty_bar = full_bar + p.NG;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = p.nw;
    const int first_chain = blockIdx.x * nw;
    const int nactive = min(nw, p.R - first_chain);
    const int NG = p.NG;
    const JT *Jg = reinterpret_cast<const JT *>(p.J);

    if constexpr (TMA) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < NG; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], (uint32_t)nactive);
            }
            mbar_fence_init();
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ producer warp
    if (warp == nw) {
        if constexpr (TMA) {
            if (lane == 0) {
                const int64_t nq = (p.nsteps + G - 1) / G;
                int site = p.start;
                int slot = 0;
                uint32_t ph = 0;  // slot = q % NG, ph = (q / NG) & 1, kept incrementally (no 64-bit divisions)
                for (int64_t q = 0; q < nq; ++q) {
                    mbar_wait(&empty_bar[slot], ph ^ 1u);
                    const int64_t left = p.nsteps - q * G;
                    const int rows = left < G ? (int)left : G;
                    mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(rows * ROWB));
                    for (int g = 0; g < rows; ++g) {
                        int i;
                        if constexpr (LIST) {
                            i = __ldg(&p.nodes[q * G + g]);
                        } else {
                            i = site;
                            if (++site == p.n) site = 0;
                        }
                        bulk_g2s(ring + ((size_t)slot * G + g) * NPAD, Jg + (int64_t)i * p.ldj, ROWB,
                                 &full_bar[slot]);
                    }
                    if (++slot == NG) {
                        slot = 0;
                        ph ^= 1u;
                    }
                }
                // no bulk copy may still be in flight when the CTA retires: wait for the last min(nq, NG) groups
                const int last = (int)(nq < NG ? nq : NG);
                for (int b = 0; b < last; ++b) {
                    if (--slot < 0) {
                        slot = NG - 1;
                        ph ^= 1u;
                    }
                    mbar_wait(&full_bar[slot], ph);
                }
            }
        }
        return;
    }
    if (warp >= nactive) return;

    // ------------------------------------------------------------------ consumer warp = one chain
    const int r = first_chain + warp;
    HT hf[NPL];
    uint32_t sw = 0;
    {
        const HT *fr = reinterpret_cast<const HT *>(p.fields) + (int64_t)r * NPAD;
        const int8_t *sr = p.spins + (int64_t)r * p.lds;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            hf[k] = fr[k * 32 + lane];
            sw |= (sr[k * 32 + lane] > 0 ? 1u : 0u) << k;
        }
    }
    const bool metro = p.rule == 2;
    const int rule = p.rule;
    const bool audit = p.tie_eps > 0.0;
    unsigned long long nflips = 0, nties = 0;
    int64_t qcur = -1;  // ring group currently held; cslot = qcur % NG, cph = (qcur / NG) & 1 (incremental)
    int cslot = -1;
    uint32_t cph = 0;

    // make group q the held one: release the previous groups in order, wait for each new one
    auto advance_to = [&](int64_t q) {
        if constexpr (TMA) {
            while (qcur < q) {
                if (qcur >= 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[cslot]);
                }
                ++qcur;
                if (++cslot == NG) {
                    cslot = 0;
                    cph ^= 1u;
                }
                mbar_wait(&full_bar[cslot], cph);
            }
        }
    };
    // row of J for step t (visiting `site`; its group must be the held one): ring slot when streamed by TMA,
    // global memory otherwise
    auto row_ptr = [&](int64_t t, int site) -> const JT * {
        if constexpr (TMA) {
            return ring + ((size_t)cslot * G + (size_t)((int)t & (G - 1))) * NPAD;
        } else {
            return Jg + (int64_t)site * p.ldj;
        }
    };
    // fields += d * J[row]
    auto apply_row = [&](const JT *row, HT d) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const JPack<JT, VEC> pk = *reinterpret_cast<const JPack<JT, VEC> *>(row + c * 32 * VEC + lane * VEC);
#pragma unroll
            for (int b = 0; b < VEC; ++b) hf[c * VEC + b] += d * (HT)pk.v[b];
        }
    };
    auto write_trace = [&](int64_t idx) {
        double sf = 0.0, sh = 0.0;
        int m = 0;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int site = k * 32 + lane;
            const bool up = (sw >> k) & 1u;
            const double f = (double)hf[k];
            sf += up ? f : -f;
            if (site < p.n) {
                const double hv = __ldg(&p.hext[site]);
                sh += up ? hv : -hv;
                m += up ? 1 : -1;
            }
        }
        sf = warp_sum(sf);
        sh = warp_sum(sh);
        m = warp_sum_int(m);
        if (lane == 0) {
            if (p.out_E) p.out_E[idx * p.R + r] = -0.5 * sf - p.ecoef * sh;
            if (p.out_M) p.out_M[idx * p.R + r] = (double)m;
        }
    };

    const uint64_t spT = (uint64_t)p.steps_per_T;
    int64_t next_trace = p.trace_every > 0 ? p.trace_every : INT64_MAX;
    int64_t trace_idx = 0;
    int64_t cached_ti = -1;
    double cachedT = 0.0;

    if constexpr (!LIST) {
        // ---------------------------------------------------------- sequential order, 32-site speculation
        int64_t t = 0;
        int site = p.start;
        uint64_t ti = 0, tr = 0;  // t = ti*spT + tr
        while (t < p.nsteps) {
            const int k = site >> 5;
            const int l_first = site & 31;
            int len = 32 - l_first;
            if (p.n - site < len) len = p.n - site;
            if (p.nsteps - t < len) len = (int)(p.nsteps - t);
            if (next_trace - t < len) len = (int)(next_trace - t);
            const int off = lane - l_first;
            const bool mine = off >= 0 && off < len;
            // temperature of my step
            double Tl;
            {
                const uint64_t a_hi = tr + (uint64_t)(len - 1);
                if (a_hi < spT) {
                    if ((int64_t)ti != cached_ti) {
                        cachedT = __ldg(&p.Tsched[ti]);
                        cached_ti = (int64_t)ti;
                    }
                    Tl = cachedT;
                } else {
                    const uint64_t a = tr + (uint64_t)(mine ? off : 0);
                    Tl = __ldg(&p.Tsched[ti + a / spT]);
                }
            }
            // fluctuation of my step
            double f = 0.0;
            if (rule != 0 && mine) {
                const int64_t tl = t + off;
                if (p.fluct_mode == 0) {
                    f = ssf_fluct_from_word(
                        rule, philox_step_word(p.seed, DOM_SSF_FLUCT, (uint32_t)r, p.step_offset + (uint64_t)tl));
                } else if (p.fluct_mode == 1) {
                    f = __ldg(&p.fluct[tl]);
                } else {
                    f = __ldg(&p.fluct[(int64_t)r * p.nsteps + tl]);
                }
            }
            const double ftl = __dmul_rn(f, Tl);
            bool mybit = (sw >> k) & 1u;
            // hk mirrors hf[k] (my own site's field) for this block; both receive identical updates
            HT hk = FieldSel<HT, 0, NPL>::get(hf, k);
            const int kpos = (k / VEC) * (32 * VEC) + lane * VEC + (k % VEC);
            uint32_t rem = __ballot_sync(FULL, mine);
            while (true) {
                const double h2 = 2.0 * (double)hk;
                const double fts = metro ? (mybit ? ftl : -ftl) : ftl;
                const double x = __dsub_rn(h2, fts);
                const bool nb = !(x < 0.0);  // heaviside(0) = 1, src/SpinSystems.jl:163-171
                const uint32_t fm = __ballot_sync(FULL, nb != mybit) & rem;
                if (audit) {
                    const uint32_t tm = __ballot_sync(FULL, fabs(x) < p.tie_eps) & rem;
                    nties += __popc(fm ? (tm & ((2u << (__ffs(fm) - 1)) - 1u)) : tm);
                }
                if (fm == 0) break;
                const int l0 = __ffs(fm) - 1;
                const uint32_t upto = (2u << l0) - 1u;  // lanes <= l0 (l0 == 31 -> all ones)
                const bool up = (__ballot_sync(FULL, nb) >> l0) & 1u;
                const int64_t tt = t + (l0 - l_first);
                advance_to(tt / G);
                const JT *row = row_ptr(tt, k * 32 + l0);
                const HT d = up ? (HT)2 : (HT)-2;
                hk += d * (HT)row[kpos];
                apply_row(row, d);
                if (lane == l0) {
                    sw ^= 1u << k;
                    mybit = !mybit;
                }
                ++nflips;
                rem &= ~upto;
                if (rem == 0) break;
            }
            t += len;
            site += len;
            if (site >= p.n) site = 0;
            tr += (uint64_t)len;
            if (tr >= spT) {
                ti += tr / spT;
                tr %= spT;
            }
            advance_to((t - 1) / G);
            if (t == next_trace) {
                write_trace(trace_idx++);
                next_trace += p.trace_every;
            }
        }
    } else {
        // ---------------------------------------------------------- explicit site list, one step at a time
        uint64_t ti = 0, tr = 0;
        double Tcur = 0.0;
        int node_batch = 0;
        double f_batch = 0.0;
        for (int64_t t = 0; t < p.nsteps; ++t) {
            const int j = (int)(t & 31);
            if (j == 0) {  // every 32 steps: each lane fetches the node and fluctuation of step t + lane
                const int64_t tl = t + lane;
                node_batch = 0;
                f_batch = 0.0;
                if (tl < p.nsteps) {
                    node_batch = __ldg(&p.nodes[tl]);
                    if (rule != 0) {
                        if (p.fluct_mode == 0) {
                            f_batch = ssf_fluct_from_word(rule, philox_step_word(p.seed, DOM_SSF_FLUCT, (uint32_t)r,
                                                                                p.step_offset + (uint64_t)tl));
                        } else if (p.fluct_mode == 1) {
                            f_batch = __ldg(&p.fluct[tl]);
                        } else {
                            f_batch = __ldg(&p.fluct[(int64_t)r * p.nsteps + tl]);
                        }
                    }
                }
            }
            if ((int64_t)ti != cached_ti) {
                Tcur = __ldg(&p.Tsched[ti]);
                cached_ti = (int64_t)ti;
            }
            const int site = __shfl_sync(FULL, node_batch, j);
            const double f = __shfl_sync(FULL, f_batch, j);
            const int k = site >> 5, l = site & 31;
            const double ft = __dmul_rn(f, Tcur);
            const bool mybit = (sw >> k) & 1u;
            const double h2 = 2.0 * (double)FieldSel<HT, 0, NPL>::get(hf, k);
            const double fts = metro ? (mybit ? ft : -ft) : ft;
            const double x = __dsub_rn(h2, fts);
            const bool nb = !(x < 0.0);
            uint32_t code = (nb != mybit ? 1u : 0u) | (nb ? 2u : 0u) | (fabs(x) < p.tie_eps ? 4u : 0u);
            code = __shfl_sync(FULL, code, l);
            if (audit) nties += (code >> 2) & 1u;
            advance_to(t / G);
            if (code & 1u) {
                apply_row(row_ptr(t, site), (code & 2u) ? (HT)2 : (HT)-2);
                if (lane == l) sw ^= 1u << k;
                ++nflips;
            }
            if (++tr == spT) {
                tr = 0;
                ++ti;
            }
            if (t + 1 == next_trace) {
                write_trace(trace_idx++);
                next_trace += p.trace_every;
            }
        }
    }

    // ------------------------------------------------------------------ write the chain back
    {
        HT *fr = reinterpret_cast<HT *>(p.fields) + (int64_t)r * NPAD;
        int8_t *sr = p.spins + (int64_t)r * p.lds;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            fr[k * 32 + lane] = hf[k];
            if (k * 32 + lane < p.n) sr[k * 32 + lane] = ((sw >> k) & 1u) ? (int8_t)1 : (int8_t)-1;
        }
        if (lane == 0) {
            p.flips[r] = nflips;
            if (nties) atomicAdd(p.near_ties, nties);
        }
    }
}

}  // namespace isb
