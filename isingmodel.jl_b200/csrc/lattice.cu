// lattice.cu — single-spin-flip sweeps on a periodic square lattice (the stencil specialisation of
// the sparse-J path; BASELINE config 1 is the reference's own test / demo lattice, test/runtests.jl:20, demo.jl:60-62).
//
// isb_model_sparse recognises a CSC matrix whose every site i = x + L y (L a multiple of 32) is coupled to exactly its
// four periodic neighbours (any bond values).  Sequential sweeps of such a model run here instead of in the
// neighbour-list kernel (sparse.cu):
//   * one warp = one chain; the chain's spins are L x L BITS in shared memory (32 words for the 32 x 32 lattice), no
//     cached local fields: a decision recomputes its field from the four neighbour spins, summed in ascending
//     neighbour index exactly like the reference's row dot (src/SpinSystems.jl:80-83) — the field is bit-identical to
//     the Float64 reference for ANY bond values, not only dyadic ones;
//   * a window of 32 consecutive sites is one lattice row segment: lane l decides site (y, x0 + l).  The only earlier
//     site of the window that lane l depends on is its left neighbour, so every lane evaluates its decision for BOTH
//     values of that neighbour and the 32 sequential decisions are resolved by a warp scan over function composition
//     (5 shuffle steps) — exactly the sequential single-site dynamics (src/SamplingHelper.jl:45-49), without the
//     flip-by-flip replay of the generic kernels.  (L = 32: lane 31's right neighbour is lane 0 of the same window,
//     already updated; lane 0 depends on nothing in the window, so it is decided first.)
//   * the Philox words of 128 consecutive steps are drawn at once (one block per lane, the same words as every other
//     kernel of the library) and handed out with shuffles.
// Rules, noise transforms, tie rule, schedules, per-replica temperatures, traces: as in ssf_kernel.cuh / sparse.cu.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "handles.hpp"

namespace isb {

struct LatParams {
    const double *Jn;      // [4][n]: couplings of site i to its neighbours in ascending neighbour index
    const uint8_t *kind;   // [n]: 2 bits per slot, which neighbour that is (0 up y-1, 1 left x-1, 2 right x+1, 3 down y+1)
    const double *hext;
    int8_t *spins;
    int64_t lds;
    int L, n, R, rule;
    int64_t nsteps;
    int start;             // first site (sequential order) / first position of the sweep order (checkerboard order)
    int fluct_mode;
    const double *fluct;
    uint64_t step_offset;
    const double *Tsched;
    const double *tscale;
    int64_t steps_per_T;
    int64_t trace_every;
    double *out_E, *out_M;
    int8_t *out_S;
    int64_t ldS;
    unsigned long long *flips, *near_ties;
    double tie_eps;
    int chains_per_cta;
    PhiloxKeys keys;
};

__global__ void __launch_bounds__(256) ssf_lattice_kernel(const LatParams p) {
    extern __shared__ uint32_t lat_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * p.chains_per_cta + warp;
    if (r >= p.R) return;
    constexpr uint32_t FULL = 0xffffffffu;
    const int L = p.L, n = p.n, WPR = L >> 5, nwords = n >> 5;
    uint32_t *W = lat_smem + (size_t)warp * nwords;
    for (int wi = 0; wi < nwords; ++wi) {
        const uint32_t m = __ballot_sync(FULL, p.spins[(int64_t)r * p.lds + wi * 32 + lane] > 0);
        if (lane == 0) W[wi] = m;
    }
    __syncwarp();
    const int rule = p.rule;
    const bool metro = rule == 2, audit = p.tie_eps > 0.0;
    const double hsign = rule == 0 ? -1.0 : 1.0;   // Hopfield thresholds J s - h (src/SingleSpinFlip.jl:31-36)
    const double tsc = p.tscale ? __ldg(&p.tscale[r]) : 1.0;
    unsigned long long nflips = 0, nties = 0;

    // sum_j J_ij s_j in ascending j (+ / - h_i) for the site of this lane.  nbits: the neighbour spins as bits (bit 0 up,
    // 1 left, 2 right, 3 down; 1 = +1); sh0..sh3 = which of them slot k of the ascending order is.  Branch-free: the
    // sign of J_k s is applied by flipping the sign bit of J_k's high word.
    auto field = [&](double j0, double j1, double j2, double j3, uint32_t sh0, uint32_t sh1, uint32_t sh2, uint32_t sh3,
                     uint32_t nbits, double hx) {
        auto term = [&](double j, uint32_t sh) {
            const uint32_t flip = (((nbits >> sh) & 1u) ^ 1u) << 31;     // spin -1: negate
            return __hiloint2double(__double2hiint(j) ^ (int)flip, __double2loint(j));
        };
        double acc = __dadd_rn(term(j0, sh0), term(j1, sh1));
        acc = __dadd_rn(acc, term(j2, sh2));
        acc = __dadd_rn(acc, term(j3, sh3));
        return __dadd_rn(acc, hx);
    };
    auto write_trace = [&](int64_t idx) {
        double q = 0.0, l = 0.0;
        int m = 0;
        for (int wi = 0; wi < nwords; ++wi) {
            const int i = wi * 32 + lane, y = wi / WPR, wc = wi % WPR;
            const uint32_t cur = W[wi], up = W[((y + L - 1) % L) * WPR + wc], dn = W[((y + 1) % L) * WPR + wc];
            const uint32_t lw = W[y * WPR + (wc + WPR - 1) % WPR], rw = W[y * WPR + (wc + 1) % WPR];
            const bool sb = (cur >> lane) & 1u;
            const bool lb = lane > 0 ? (cur >> (lane - 1)) & 1u : (lw >> 31) & 1u;
            const bool rb = lane < 31 ? (cur >> (lane + 1)) & 1u : rw & 1u;
            const uint32_t code = __ldg(&p.kind[i]);
            const uint32_t nbits = ((up >> lane) & 1u) | ((lb ? 1u : 0u) << 1) | ((rb ? 1u : 0u) << 2) | (((dn >> lane) & 1u) << 3);
            const double g = field(__ldg(&p.Jn[i]), __ldg(&p.Jn[n + i]), __ldg(&p.Jn[2 * n + i]), __ldg(&p.Jn[3 * n + i]),
                                   code & 3u, (code >> 2) & 3u, (code >> 4) & 3u, (code >> 6) & 3u, nbits, 0.0);
            const double hv = __ldg(&p.hext[i]);
            q += sb ? g : -g;
            l += sb ? hv : -hv;
            m += sb ? 1 : -1;
        }
        q = warp_sum(q);
        l = warp_sum(l);
        m = warp_sum_int(m);
        if (lane == 0) {
            if (p.out_E) p.out_E[idx * p.R + r] = -0.5 * q - l;
            if (p.out_M) p.out_M[idx * p.R + r] = (double)m;
        }
        if (p.out_S)
            for (int wi = 0; wi < nwords; ++wi)
                p.out_S[(idx * p.R + r) * p.ldS + wi * 32 + lane] = ((W[wi] >> lane) & 1u) ? (int8_t)1 : (int8_t)-1;
    };

    int64_t next_trace = p.trace_every > 0 ? p.trace_every : INT64_MAX, trace_idx = 0;
    const uint64_t spT = (uint64_t)p.steps_per_T;
    uint64_t ti = 0, tr = 0;  // t = ti * spT + tr
    int64_t cached_ti = -1;
    double cachedT = 0.0;
    // Philox blocks of 128 consecutive steps: lane j holds the block of steps 4 (cb + j) .. 4 (cb + j) + 3
    Philox4 blk{0, 0, 0, 0};
    uint64_t cb = 0;
    bool have_blk = false;

    int64_t t = 0;
    int site = p.start;
    int y = site / L, x = site - y * L;   // kept incrementally: a window never crosses the end of a lattice row
    while (t < p.nsteps) {
        const int l_first = site & 31;
        int len = 32 - l_first;
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
                    cachedT = __dmul_rn(__ldg(&p.Tsched[ti]), tsc);
                    cached_ti = (int64_t)ti;
                }
                Tl = cachedT;
            } else {
                const uint64_t a = tr + (uint64_t)(mine ? off : 0);
                Tl = __dmul_rn(__ldg(&p.Tsched[ti + a / spT]), tsc);
            }
        }
        // fluctuation of my step
        double f = 0.0;
        if (rule != 0) {
            const int64_t tl = t + (mine ? off : 0);
            if (p.fluct_mode == ISB_FLUCT_PHILOX) {
                const uint64_t g0 = p.step_offset + (uint64_t)t;            // first step of the window
                if (!have_blk || ((g0 + 31) >> 2) >= cb + 32 || (g0 >> 2) < cb) {
                    cb = g0 >> 2;
                    const uint64_t q = cb + (uint64_t)lane;
                    blk = philox4x32_10k((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)r, DOM_SSF_FLUCT << 28, p.keys);
                    have_blk = true;
                }
                const uint64_t gs = p.step_offset + (uint64_t)tl;
                const int src = (int)((gs >> 2) - cb);
                const uint32_t wx = __shfl_sync(FULL, blk.x, src), wy = __shfl_sync(FULL, blk.y, src);
                const uint32_t wz = __shfl_sync(FULL, blk.z, src), ww = __shfl_sync(FULL, blk.w, src);
                const uint32_t pick = (uint32_t)(gs & 3u);
                f = ssf_fluct_from_word(rule, pick == 0 ? wx : (pick == 1 ? wy : (pick == 2 ? wz : ww)));
            } else if (p.fluct_mode == ISB_FLUCT_SHARED) {
                f = __ldg(&p.fluct[tl]);
            } else {
                f = __ldg(&p.fluct[(int64_t)r * p.nsteps + tl]);
            }
        }
        const double ftl = __dmul_rn(f, Tl);
        // the window's row segment and its neighbours
        const int base = site - l_first, wc = x >> 5;
        const uint32_t cur = W[y * WPR + wc];
        const uint32_t up = W[(y == 0 ? L - 1 : y - 1) * WPR + wc], dn = W[(y == L - 1 ? 0 : y + 1) * WPR + wc];
        const uint32_t lw = W[y * WPR + (wc == 0 ? WPR - 1 : wc - 1)], rw = W[y * WPR + (wc == WPR - 1 ? 0 : wc + 1)];
        const bool mybit = (cur >> lane) & 1u;
        const uint32_t left_old = ((cur << 1) | (lw >> 31)) >> lane & 1u;           // the spin to my left, as stored
        const uint32_t right_old = (uint32_t)((((uint64_t)(rw & 1u) << 32) | cur) >> (lane + 1)) & 1u;
        // neighbour bits without the left one: bit 0 up, 2 right, 3 down
        uint32_t nb0 = ((up >> lane) & 1u) | (right_old << 2) | (((dn >> lane) & 1u) << 3);
        const int i = base + lane;
        const double j0 = __ldg(&p.Jn[i]), j1 = __ldg(&p.Jn[n + i]), j2 = __ldg(&p.Jn[2 * n + i]), j3 = __ldg(&p.Jn[3 * n + i]);
        const uint32_t code = __ldg(&p.kind[i]);
        const uint32_t sh0 = code & 3u, sh1 = (code >> 2) & 3u, sh2 = (code >> 4) & 3u, sh3 = (code >> 6) & 3u;
        const double hx = hsign * __ldg(&p.hext[i]);
        const double fts = metro ? (mybit ? ftl : -ftl) : ftl;
        // x = 2 h_loc - f T [s_i]; new spin = +1 unless x < 0 (heaviside(0) = 1, src/SpinSystems.jl:163-171)
        double x0 = __dsub_rn(2.0 * field(j0, j1, j2, j3, sh0, sh1, sh2, sh3, nb0, hx), fts);
        double x1 = __dsub_rn(2.0 * field(j0, j1, j2, j3, sh0, sh1, sh2, sh3, nb0 | 2u, hx), fts);
        bool t0 = !(x0 < 0.0), t1 = !(x1 < 0.0);
        if (WPR == 1 && l_first == 0) {
            // L = 32: lane 31's right neighbour is lane 0 of this window, which is updated first (lane 0 itself depends on
            // nothing in the window).  Only when lane 0 flips does lane 31 see something else than the stored spin.
            const bool b0 = __shfl_sync(FULL, left_old ? t1 : t0, 0);
            if (b0 != (bool)(cur & 1u)) {
                const uint32_t nb31 = lane == 31 ? (nb0 ^ 4u) : nb0;
                x0 = __dsub_rn(2.0 * field(j0, j1, j2, j3, sh0, sh1, sh2, sh3, nb31, hx), fts);
                x1 = __dsub_rn(2.0 * field(j0, j1, j2, j3, sh0, sh1, sh2, sh3, nb31 | 2u, hx), fts);
                t0 = !(x0 < 0.0);
                t1 = !(x1 < 0.0);
            }
        }
        // my site as a function of the NEW value of my left neighbour: (f0, f1) = value for left = -1 / +1.  Lanes outside
        // the window keep their spin; the first lane of the window sees its left neighbour's current value.
        bool f0 = mine ? t0 : mybit, f1 = mine ? t1 : mybit;
        if (off == 0 || lane == 0) f0 = f1 = mine ? (left_old ? t1 : t0) : mybit;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const bool g0 = __shfl_up_sync(FULL, f0, d), g1 = __shfl_up_sync(FULL, f1, d);
            if (lane >= d) {
                const bool n0 = g0 ? f1 : f0, n1 = g1 ? f1 : f0;
                f0 = n0;
                f1 = n1;
            }
        }
        const uint32_t newmask = __ballot_sync(FULL, f0);   // after the scan f0 == f1 = the new spin of every lane
        nflips += __popc(newmask ^ cur);
        if (audit) {
            const uint32_t lmask = (newmask << 1) | (left_old ? 1u : 0u);   // the left neighbour each lane actually saw
            const bool tie = ((lmask >> lane) & 1u) ? fabs(x1) < p.tie_eps : fabs(x0) < p.tie_eps;
            nties += __popc(__ballot_sync(FULL, tie && mine));
        }
        __syncwarp();
        if (lane == 0) W[y * WPR + wc] = newmask;
        __syncwarp();
        t += len;
        site += len;
        x += len;
        if (x >= L) {
            x = 0;
            if (++y == L) y = 0;
        }
        if (site >= n) site = 0;
        tr += (uint64_t)len;
        if (tr >= spT) {
            ti += tr / spT;
            tr %= spT;
        }
        if (t == next_trace) {
            write_trace(trace_idx++);
            next_trace += p.trace_every;
        }
    }
    __syncwarp();
    for (int wi = 0; wi < nwords; ++wi)
        p.spins[(int64_t)r * p.lds + wi * 32 + lane] = ((W[wi] >> lane) & 1u) ? (int8_t)1 : (int8_t)-1;
    if (lane == 0) {
        p.flips[r] = nflips;
        if (nties) atomicAdd(p.near_ties, nties);
    }
}

// ISB_ORDER_CHECKERBOARD ("checkerboard colouring"): a sweep visits the sites of colour (x + y) even in
// ascending site index, then the sites of the other colour — one particular site list of the reference's 3-argument
// update! (src/SingleSpinFlip.jl:46-55; tests replay exactly that list through the oracle).  The four neighbours of a
// site all have the other colour, so the 32 sites of a window do not depend on each other: every lane decides its site
// from the stored spins (one field evaluation, no scan), and the decisions of 16 lanes are spread over the colour's bit
// positions of one word.  Everything else — rules, noise words per step, schedules, traces — as in ssf_lattice_kernel.
__global__ void __launch_bounds__(256) ssf_lattice_cb_kernel(const LatParams p) {
    extern __shared__ uint32_t lat_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * p.chains_per_cta + warp;
    if (r >= p.R) return;
    constexpr uint32_t FULL = 0xffffffffu;
    const int L = p.L, n = p.n, WPR = L >> 5, nwords = n >> 5, Lh = L >> 1, n2 = n >> 1;
    uint32_t *W = lat_smem + (size_t)warp * nwords;
    for (int wi = 0; wi < nwords; ++wi) {
        const uint32_t m = __ballot_sync(FULL, p.spins[(int64_t)r * p.lds + wi * 32 + lane] > 0);
        if (lane == 0) W[wi] = m;
    }
    __syncwarp();
    const int rule = p.rule;
    const bool metro = rule == 2, audit = p.tie_eps > 0.0;
    const double hsign = rule == 0 ? -1.0 : 1.0;
    const double tsc = p.tscale ? __ldg(&p.tscale[r]) : 1.0;
    unsigned long long nflips = 0, nties = 0;
    auto bit = [&](int yy, int xx) { return (W[yy * WPR + (xx >> 5)] >> (xx & 31)) & 1u; };
    // sum_j J_ij s_j in ascending j (+ / - h_i): the reference's row-dot order (src/SpinSystems.jl:80-83)
    auto field_at = [&](int i, uint32_t nbits, double hx) {
        const uint32_t code = __ldg(&p.kind[i]);
        auto term = [&](int k) {
            const double j = __ldg(&p.Jn[(size_t)k * n + i]);
            const uint32_t flip = (((nbits >> ((code >> (2 * k)) & 3u)) & 1u) ^ 1u) << 31;     // spin -1: negate
            return __hiloint2double(__double2hiint(j) ^ (int)flip, __double2loint(j));
        };
        double acc = __dadd_rn(term(0), term(1));
        acc = __dadd_rn(acc, term(2));
        acc = __dadd_rn(acc, term(3));
        return __dadd_rn(acc, hx);
    };
    auto nbits_at = [&](int y, int x) {
        const int yu = y == 0 ? L - 1 : y - 1, yd = y == L - 1 ? 0 : y + 1, xl = x == 0 ? L - 1 : x - 1, xr = x == L - 1 ? 0 : x + 1;
        return bit(yu, x) | (bit(y, xl) << 1) | (bit(y, xr) << 2) | (bit(yd, x) << 3);
    };
    auto write_trace = [&](int64_t idx) {
        double q = 0.0, l = 0.0;
        int m = 0;
        for (int wi = 0; wi < nwords; ++wi) {
            const int i = wi * 32 + lane, y = i / L, x = i - y * L;
            const bool sb = bit(y, x);
            const double g = field_at(i, nbits_at(y, x), 0.0);
            const double hv = __ldg(&p.hext[i]);
            q += sb ? g : -g;
            l += sb ? hv : -hv;
            m += sb ? 1 : -1;
        }
        q = warp_sum(q);
        l = warp_sum(l);
        m = warp_sum_int(m);
        if (lane == 0) {
            if (p.out_E) p.out_E[idx * p.R + r] = -0.5 * q - l;
            if (p.out_M) p.out_M[idx * p.R + r] = (double)m;
        }
        if (p.out_S)
            for (int wi = 0; wi < nwords; ++wi)
                p.out_S[(idx * p.R + r) * p.ldS + wi * 32 + lane] = ((W[wi] >> lane) & 1u) ? (int8_t)1 : (int8_t)-1;
    };
    int64_t next_trace = p.trace_every > 0 ? p.trace_every : INT64_MAX, trace_idx = 0;
    const uint64_t spT = (uint64_t)p.steps_per_T;
    uint64_t ti = 0, tr = 0;  // t = ti * spT + tr
    int64_t cached_ti = -1;
    double cachedT = 0.0;
    Philox4 blk{0, 0, 0, 0};   // Philox blocks of 128 consecutive steps, as in ssf_lattice_kernel
    uint64_t cb = 0;
    bool have_blk = false;

    int64_t t = 0;
    int pos = p.start;         // position in the sweep order: [0, n/2) first colour, [n/2, n) second colour
    while (t < p.nsteps) {
        const int c = pos >= n2 ? 1 : 0;
        const int pc = pos - c * n2;
        const int l_first = pc & 31;          // windows are aligned to 32 positions of a colour (n / 2 is a multiple of 32)
        int len = 32 - l_first;
        if (p.nsteps - t < len) len = (int)(p.nsteps - t);
        if (next_trace - t < len) len = (int)(next_trace - t);
        const int off = lane - l_first;
        const bool mine = off >= 0 && off < len;
        double Tl;
        {
            const uint64_t a_hi = tr + (uint64_t)(len - 1);
            if (a_hi < spT) {
                if ((int64_t)ti != cached_ti) {
                    cachedT = __dmul_rn(__ldg(&p.Tsched[ti]), tsc);
                    cached_ti = (int64_t)ti;
                }
                Tl = cachedT;
            } else {
                const uint64_t a = tr + (uint64_t)(mine ? off : 0);
                Tl = __dmul_rn(__ldg(&p.Tsched[ti + a / spT]), tsc);
            }
        }
        double f = 0.0;
        if (rule != 0) {
            const int64_t tl = t + (mine ? off : 0);
            if (p.fluct_mode == ISB_FLUCT_PHILOX) {
                const uint64_t g0 = p.step_offset + (uint64_t)t;
                if (!have_blk || ((g0 + 31) >> 2) >= cb + 32 || (g0 >> 2) < cb) {
                    cb = g0 >> 2;
                    const uint64_t q = cb + (uint64_t)lane;
                    blk = philox4x32_10k((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)r, DOM_SSF_FLUCT << 28, p.keys);
                    have_blk = true;
                }
                const uint64_t gs = p.step_offset + (uint64_t)tl;
                const int src = (int)((gs >> 2) - cb);
                const uint32_t wx = __shfl_sync(FULL, blk.x, src), wy = __shfl_sync(FULL, blk.y, src);
                const uint32_t wz = __shfl_sync(FULL, blk.z, src), ww = __shfl_sync(FULL, blk.w, src);
                const uint32_t pick = (uint32_t)(gs & 3u);
                f = ssf_fluct_from_word(rule, pick == 0 ? wx : (pick == 1 ? wy : (pick == 2 ? wz : ww)));
            } else if (p.fluct_mode == ISB_FLUCT_SHARED) {
                f = __ldg(&p.fluct[tl]);
            } else {
                f = __ldg(&p.fluct[(int64_t)r * p.nsteps + tl]);
            }
        }
        const double ftl = __dmul_rn(f, Tl);
        // my site: the (pc - l_first + lane)-th site of colour c in ascending site index
        const int pl = pc - l_first + lane;
        const int y = pl / Lh, j = pl - y * Lh;
        const int par = (y + c) & 1;
        const int x = 2 * j + par, i = y * L + x;
        const uint32_t mybit = bit(y, x);
        const double hx = hsign * __ldg(&p.hext[i]);
        const double fts = metro ? (mybit ? ftl : -ftl) : ftl;
        // x = 2 h_loc - f T [s_i]; new spin = +1 unless x < 0 (heaviside(0) = 1, src/SpinSystems.jl:163-171)
        const double xv = __dsub_rn(2.0 * field_at(i, nbits_at(y, x), hx), fts);
        const bool nb = mine ? !(xv < 0.0) : (bool)mybit;
        nflips += __popc(__ballot_sync(FULL, mine && nb != (bool)mybit));
        if (audit) nties += __popc(__ballot_sync(FULL, mine && fabs(xv) < p.tie_eps));
        // lanes 0-15 and 16-31 each hold the 16 sites of colour c of one 32-bit word (same row, x = 2 j + par)
        const uint32_t b = __ballot_sync(FULL, nb);
        uint32_t v = (b >> (lane & 16)) & 0xFFFFu;
        v = (v | (v << 8)) & 0x00FF00FFu;      // spread the 16 decisions over every other bit
        v = (v | (v << 4)) & 0x0F0F0F0Fu;
        v = (v | (v << 2)) & 0x33333333u;
        v = (v | (v << 1)) & 0x55555555u;
        __syncwarp();
        if ((lane & 15) == 0) {
            uint32_t *wp = &W[y * WPR + (x >> 5)];
            *wp = (*wp & ~(0x55555555u << par)) | (v << par);
        }
        __syncwarp();
        t += len;
        pos += len;
        if (pos >= n) pos = 0;
        tr += (uint64_t)len;
        if (tr >= spT) {
            ti += tr / spT;
            tr %= spT;
        }
        if (t == next_trace) {
            write_trace(trace_idx++);
            next_trace += p.trace_every;
        }
    }
    __syncwarp();
    for (int wi = 0; wi < nwords; ++wi)
        p.spins[(int64_t)r * p.lds + wi * 32 + lane] = ((W[wi] >> lane) & 1u) ? (int8_t)1 : (int8_t)-1;
    if (lane == 0) {
        p.flips[r] = nflips;
        if (nties) atomicAdd(p.near_ties, nties);
    }
}

// ------------------------------------------------------------------ host
struct LatticeModel {
    int L = 0;
    double *Jn = nullptr;
    uint8_t *kind = nullptr;
};

// rows[i]: the (column, value) pairs of row i of the symmetric J, ascending.  Returns a LatticeModel when the graph is
// the periodic L x L square lattice with sites numbered i = x + L y (L a multiple of 32), else NULL.
void *lattice_detect(isb_ctx *ctx, int n, const std::vector<std::vector<std::pair<int, double>>> &rows) {
    if (const char *env = getenv("ISB_LATTICE"))
        if (atoi(env) == 0) return nullptr;
    const int L = (int)lrint(sqrt((double)n));
    if (L * L != n || L < 32 || L % 32 != 0) return nullptr;
    std::vector<double> Jn((size_t)4 * n);
    std::vector<uint8_t> kind((size_t)n);
    for (int i = 0; i < n; ++i) {
        if (rows[i].size() != 4) return nullptr;
        const int x = i % L, y = i / L;
        const int nb[4] = {x + L * ((y + L - 1) % L), (x + L - 1) % L + L * y, (x + 1) % L + L * y, x + L * ((y + 1) % L)};
        uint32_t code = 0;
        for (int k = 0; k < 4; ++k) {
            int which = -1;
            for (int c = 0; c < 4; ++c)
                if (nb[c] == rows[i][k].first) which = c;
            if (which < 0) return nullptr;
            code |= (uint32_t)which << (2 * k);
            Jn[(size_t)k * n + i] = rows[i][k].second;
        }
        kind[i] = (uint8_t)code;
    }
    LatticeModel *lm = new LatticeModel();
    lm->L = L;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(&lm->Jn, Jn.size() * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&lm->kind, kind.size()) != cudaSuccess ||
        cudaMemcpy(lm->Jn, Jn.data(), Jn.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(lm->kind, kind.data(), kind.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(lm->Jn);
        cudaFree(lm->kind);
        delete lm;
        return nullptr;   // the neighbour-list kernel serves the model
    }
    return lm;
}

void lattice_free(void *lat) {
    LatticeModel *lm = (LatticeModel *)lat;
    if (!lm) return;
    cudaFree(lm->Jn);
    cudaFree(lm->kind);
    delete lm;
}

int lattice_side(const void *lat) { return lat ? ((const LatticeModel *)lat)->L : 0; }

int ssf_lattice_run_device(isb_ens *e, void *lat, int rule, int64_t nsteps, int order, int start, int fluct_mode, const double *d_fluct,
                           uint64_t seed, uint64_t step_offset, const double *d_T, int64_t steps_per_T, int64_t trace_every,
                           double *d_E, double *d_M, int8_t *d_S) {
    isb_model *m = e->model;
    isb_ctx *ctx = m->ctx;
    LatticeModel *lm = (LatticeModel *)lat;
    if (nsteps <= 0) return ISB_OK;
    LatParams p{};
    p.Jn = lm->Jn; p.kind = lm->kind; p.hext = m->h64;
    p.spins = e->spins; p.lds = e->lds;
    p.L = lm->L; p.n = m->n; p.R = e->R; p.rule = rule;
    p.nsteps = nsteps; p.start = start;
    p.fluct_mode = fluct_mode; p.fluct = d_fluct; p.step_offset = step_offset;
    p.Tsched = d_T; p.tscale = e->d_tscale; p.steps_per_T = steps_per_T; p.trace_every = trace_every;
    p.out_E = d_E; p.out_M = d_M; p.out_S = d_S; p.ldS = m->n;
    p.flips = e->d_flips; p.near_ties = e->d_counters; p.tie_eps = e->tie_eps;
    p.keys = philox_keys(seed);
    // one warp per chain, the chains spread over the SMs first; a chain's bits take n / 8 bytes of shared memory
    const size_t per_chain = (size_t)m->n / 8;
    int chains = (int)std::min<size_t>(8, (ctx->smem_optin - 1024) / per_chain);
    if (chains < 1) return fail(ctx, ISB_ERR_UNSUPPORTED, "lattice sweeps: L = %d is too large for shared memory", lm->L);
    const int per_sm = (e->R + ctx->num_sms - 1) / ctx->num_sms;
    chains = std::max(1, std::min(chains, std::max(1, per_sm / 4)));   // >= 4 CTAs per SM when the replicas allow it
    p.chains_per_cta = chains;
    const int grid = (e->R + chains - 1) / chains;
    const size_t smem = per_chain * chains;
    auto kern = order == ISB_ORDER_CHECKERBOARD ? ssf_lattice_cb_kernel : ssf_lattice_kernel;
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce == cudaSuccess) {
        kern<<<grid, 32 * chains, smem, ctx->stream>>>(p);
        ce = cudaGetLastError();
    }
    if (ce != cudaSuccess) return fail(ctx, ISB_ERR_CUDA, "ssf_lattice_kernel launch failed: %s", cudaGetErrorString(ce));
    e->fields_rule_sign = 0;   // the cached local fields of the neighbour-list kernel no longer match the spins
    e->last_launches += 1;
    return ISB_OK;
}

}  // namespace isb
