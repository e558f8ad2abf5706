// bip_exact.cu — Float64 block-Gibbs path for SpinSystemOnBipartiteGraph ensembles (ISB_PREC_F64).
//
// Replaces, for R chains at once,
//   OnBipartiteGraph.update!(::StochasticCellularAutomata, Fv, Fh)   src/OnBipartiteGraph.jl:30-43
//   OnBipartiteGraph.update!(::MomentumAnnealing, Fv, Fh)            src/OnBipartiteGraph.jl:53-66
//   calcLocalAuxiliaryBias  W' sigma + b                              src/SpinSystems.jl:154-157
//   calcLocalMagneticField  W tau + h                                 src/SpinSystems.jl:147-150
//   calcEnergy              -sigma' W tau - h' sigma - b' tau         src/SpinSystems.jl:139-143
//
// This is the bit-exact path: every output unit sums its W column sequentially over ascending input
// index in double (the order of the reference's generic mat-vec and of the CPU test oracle), so the
// decision quantity 2*(W's + b) - F*T is bit-identical to the Float64 reference for ANY W, not only
// for integer / dyadic couplings.  The tensor-core path (bip_tc.cu) is the throughput path.
//
// Layout: one thread per output unit, CH chains per CTA.  The CH input spin vectors are packed into a
// shared-memory bit mask per input unit, so one coalesced read of a W row serves CH chains.
#include "common.cuh"
#include "handles.hpp"

namespace isb {

constexpr int BIPX_CH = 8;        // chains per CTA
constexpr int BIPX_THREADS = 128; // output units per CTA

struct BipHalfParams {
    const double *Wm;    // [nin][ldw]: element (input i, output j) at i*ldw + j
    int64_t ldw;
    const double *bias;  // [nout]
    const int8_t *in;    // [R][ldin]
    int64_t ldin;
    int nin;
    int8_t *out;         // [R][ldout]  (also the unit's own previous value for MomentumAnnealing)
    int64_t ldout;
    int nout;
    int R;
    int rule;            // ISB_BIP_SCA / ISB_BIP_MA
    int fluct_mode;
    const double *F;     // SHARED: [nsteps][nout]; PER_REPLICA: [R][nsteps][nout]
    int64_t nsteps, k;   // step index within this run
    const double *Tsched;
    const double *tscale;     // per-replica temperature factors [R] or NULL
    int64_t steps_per_T;
    uint64_t seed, step_abs;  // Philox: absolute step = step_offset + k
    uint32_t domain;
    double *field_out;   // FIELD mode: [R][ldf]
    int64_t ldf;
};

template <bool FIELD>
__global__ void __launch_bounds__(BIPX_THREADS) bip_half_kernel(const BipHalfParams p) {
    extern __shared__ unsigned char smask[];  // [nin] bit c = spin of chain c is +1
    const int r0 = blockIdx.y * BIPX_CH;
    const int nch = min(BIPX_CH, p.R - r0);
    for (int i = threadIdx.x; i < p.nin; i += blockDim.x) {
        unsigned m = 0;
        for (int c = 0; c < nch; ++c) m |= (p.in[(int64_t)(r0 + c) * p.ldin + i] > 0 ? 1u : 0u) << c;
        smask[i] = (unsigned char)m;
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.nout) return;
    double acc[BIPX_CH];
#pragma unroll
    for (int c = 0; c < BIPX_CH; ++c) acc[c] = 0.0;
    const double *wp = p.Wm + j;
    for (int i = 0; i < p.nin; ++i) {
        const double w = __ldg(wp + (int64_t)i * p.ldw);
        const unsigned m = smask[i];
#pragma unroll
        for (int c = 0; c < BIPX_CH; ++c) acc[c] = __dadd_rn(acc[c], ((m >> c) & 1u) ? w : -w);
    }
    const double bj = p.bias[j];
    if constexpr (FIELD) {
        for (int c = 0; c < nch; ++c) p.field_out[(int64_t)(r0 + c) * p.ldf + j] = __dadd_rn(acc[c], bj);
        return;
    }
    const double T = p.Tsched[p.k / p.steps_per_T];
    for (int c = 0; c < nch; ++c) {
        const int r = r0 + c;
        double f;
        if (p.fluct_mode == ISB_FLUCT_PHILOX) {
            const Philox4 blk = philox_unit_block(p.seed, p.domain, (uint32_t)r, p.step_abs, (uint32_t)(j >> 2));
            const uint32_t w = philox_pick(blk, (uint32_t)(j & 3));
            f = p.rule == ISB_BIP_SCA ? logistic_from_word(w) : exponential_from_word(w);
        } else if (p.fluct_mode == ISB_FLUCT_SHARED) {
            f = p.F[p.k * p.nout + j];
        } else {
            f = p.F[((int64_t)r * p.nsteps + p.k) * p.nout + j];
        }
        int8_t *o = p.out + (int64_t)r * p.ldout + j;
        double ft = __dmul_rn(f, p.tscale ? __dmul_rn(T, p.tscale[r]) : T);
        if (p.rule == ISB_BIP_MA) ft = __dmul_rn(ft, (double)*o);
        const double x = __dsub_rn(__dmul_rn(2.0, __dadd_rn(acc[c], bj)), ft);
        *o = (x < 0.0) ? (int8_t)-1 : (int8_t)1;  // heaviside(0) = 1, src/SpinSystems.jl:163-171
    }
}

// E_r = -sum_i sigma_i (sum_j W_ij tau_j) - sum_i h_i sigma_i - sum_j b_j tau_j.
// Pass 1: CTA (x, y) covers 128 visible units of BIPE_CH chains (one coalesced read of a W row serves all of them: W is
// pulled from L2 R / 32 times per call — 32 chains instead of bip_half_kernel's 8 took config 3's energies from 37 to
// ~12 ms) and writes one partial sum per chain; pass 2 adds the partials of a chain in a fixed order (no atomics: the
// energies are reproducible bit for bit).  Every chain's sums run over ascending j like before: same values.
constexpr int BIPE_CH = 32;
__global__ void __launch_bounds__(BIPX_THREADS)
bip_energy_partial_kernel(const double *__restrict__ Wt /*[nh][nv]*/, const double *__restrict__ h,
                          const double *__restrict__ b, const int8_t *sig, int64_t lds, const int8_t *tau, int64_t ldh,
                          int nv, int nh, int R, double *partial /*[gridDim.x][R]*/) {
    extern __shared__ uint32_t tmask[];  // [nh] bit c = hidden spin of chain c is +1
    __shared__ double red[BIPX_THREADS / 32][BIPE_CH];
    const int r0 = blockIdx.y * BIPE_CH;
    const int nch = min(BIPE_CH, R - r0);
    for (int j = threadIdx.x; j < nh; j += blockDim.x) {
        uint32_t m = 0;
        for (int c = 0; c < nch; ++c) m |= (tau[(int64_t)(r0 + c) * ldh + j] > 0 ? 1u : 0u) << c;
        tmask[j] = m;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double part[BIPE_CH];
#pragma unroll
    for (int c = 0; c < BIPE_CH; ++c) part[c] = 0.0;
    if (i < nv) {
        double row[BIPE_CH];
#pragma unroll
        for (int c = 0; c < BIPE_CH; ++c) row[c] = 0.0;
        for (int j = 0; j < nh; ++j) {
            const double w = __ldg(Wt + (int64_t)j * nv + i);
            const uint32_t m = tmask[j];
#pragma unroll
            for (int c = 0; c < BIPE_CH; ++c) row[c] += ((m >> c) & 1u) ? w : -w;
        }
        const double hi = h[i];
#pragma unroll
        for (int c = 0; c < BIPE_CH; ++c)
            if (c < nch) {
                const double si = (double)sig[(int64_t)(r0 + c) * lds + i];
                part[c] = -(si * row[c] + hi * si);
            }
    }
    if (blockIdx.x == 0)  // the hidden-bias term, once per chain
        for (int j = threadIdx.x; j < nh; j += blockDim.x) {
            const double bj = b[j];
            const uint32_t m = tmask[j];
#pragma unroll
            for (int c = 0; c < BIPE_CH; ++c) part[c] -= ((m >> c) & 1u) ? bj : -bj;
        }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < BIPE_CH; ++c) {
        const double v = warp_sum(part[c]);
        if (lane == 0) red[w][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < nch) {
        double acc = 0.0;
        for (int ww = 0; ww < BIPX_THREADS / 32; ++ww) acc += red[ww][threadIdx.x];
        partial[(int64_t)blockIdx.x * R + r0 + threadIdx.x] = acc;
    }
}
__global__ void bip_energy_reduce_kernel(const double *partial, int nparts, int R, double *E) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double acc = 0.0;
    for (int k = 0; k < nparts; ++k) acc += partial[(int64_t)k * R + r];
    E[r] = acc;
}

__global__ void philox_bip_fluct_kernel(int rule, uint64_t seed, uint64_t step_offset, uint32_t domain, int nunits,
                                        int r0, int nr, int64_t nsteps, double *out) {
    const int64_t total = (int64_t)nr * nsteps * nunits;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx % nunits);
        const int64_t k = (idx / nunits) % nsteps;
        const int rr = (int)(idx / ((int64_t)nunits * nsteps));
        const Philox4 blk =
            philox_unit_block(seed, domain, (uint32_t)(r0 + rr), step_offset + (uint64_t)k, (uint32_t)(j >> 2));
        const uint32_t w = philox_pick(blk, (uint32_t)(j & 3));
        out[idx] = rule == ISB_BIP_SCA ? logistic_from_word(w) : exponential_from_word(w);
    }
}

int philox_bip_fluct_device(isb_ctx *ctx, int rule, uint64_t seed, uint64_t step_offset, int layer, int nunits,
                            int r0, int nr, int64_t nsteps, double *d_out) {
    const int64_t total = (int64_t)nr * nsteps * nunits;
    if (total == 0) return ISB_OK;
    const int64_t g = (total + 255) / 256;
    philox_bip_fluct_kernel<<<(int)(g < 148 * 16 ? g : 148 * 16), 256, 0, ctx->stream>>>(
        rule, seed, step_offset, layer == 0 ? DOM_BIP_VISIBLE : DOM_BIP_HIDDEN, nunits, r0, nr, nsteps, d_out);
    ISB_CUDA(ctx, cudaGetLastError());
    return ISB_OK;
}

int bip_energy_device(isb_ens *e, double *d_E) {
    isb_model *m = e->model;
    isb_ctx *ctx = m->ctx;
    const int nparts = (m->nv + BIPX_THREADS - 1) / BIPX_THREADS;
    double *partial;
    int rc = dev_reserve(ctx, SCR_TC0, (size_t)nparts * e->R * sizeof(double), (void **)&partial);
    if (rc) return rc;
    dim3 grid(nparts, (e->R + BIPE_CH - 1) / BIPE_CH);
    const size_t mask_bytes = (size_t)m->nh * sizeof(uint32_t);
    if (mask_bytes > 48 * 1024) {
        if (mask_bytes + 8192 > ctx->smem_optin)
            return fail(ctx, ISB_ERR_UNSUPPORTED, "bipartite energy: %d hidden units exceed the shared memory", m->nh);
        ISB_CUDA(ctx, cudaFuncSetAttribute(bip_energy_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mask_bytes));
    }
    bip_energy_partial_kernel<<<grid, BIPX_THREADS, mask_bytes, ctx->stream>>>(m->Wt64, m->hb64, m->bb64, e->spins, e->lds,
                                                                          e->hidden, e->ldh, m->nv, m->nh, e->R, partial);
    bip_energy_reduce_kernel<<<(e->R + 255) / 256, 256, 0, ctx->stream>>>(partial, nparts, e->R, d_E);
    ISB_CUDA(ctx, cudaGetLastError());
    return ISB_OK;
}

static void fill_half(BipHalfParams &p, isb_ens *e, int layer) {
    isb_model *m = e->model;
    if (layer == 1) {  // hidden from visible: sum_i W[i][j] sigma_i + b_j
        p.Wm = m->W64; p.ldw = m->nh; p.bias = m->bb64;
        p.in = e->spins; p.ldin = e->lds; p.nin = m->nv;
        p.out = e->hidden; p.ldout = e->ldh; p.nout = m->nh;
        p.domain = DOM_BIP_HIDDEN;
    } else {           // visible from hidden: sum_j W[i][j] tau_j + h_i
        p.Wm = m->Wt64; p.ldw = m->nv; p.bias = m->hb64;
        p.in = e->hidden; p.ldin = e->ldh; p.nin = m->nh;
        p.out = e->spins; p.ldout = e->lds; p.nout = m->nv;
        p.domain = DOM_BIP_VISIBLE;
    }
    p.R = e->R;
}

int bip_field_device(isb_ens *e, int layer, double *d_out, int64_t ld) {
    isb_model *m = e->model;
    BipHalfParams p{};
    fill_half(p, e, layer);
    p.field_out = d_out;
    p.ldf = ld;
    dim3 grid((p.nout + BIPX_THREADS - 1) / BIPX_THREADS, (e->R + BIPX_CH - 1) / BIPX_CH);
    bip_half_kernel<true><<<grid, BIPX_THREADS, p.nin, m->ctx->stream>>>(p);
    ISB_CUDA(m->ctx, cudaGetLastError());
    return ISB_OK;
}


// copies the canonical int8 layers into the snapshot buffers of trace point `ntr` ([ntr][R][n], dense pitch)
int bip_snapshot(isb_ens *e, int64_t ntr, int8_t *d_Sv, int8_t *d_Sh) {
    isb_model *m = e->model;
    isb_ctx *ctx = m->ctx;
    if (d_Sv)
        ISB_CUDA(ctx, cudaMemcpy2DAsync(d_Sv + ntr * e->R * m->nv, (size_t)m->nv, e->spins, (size_t)e->lds, (size_t)m->nv,
                                        (size_t)e->R, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_Sh)
        ISB_CUDA(ctx, cudaMemcpy2DAsync(d_Sh + ntr * e->R * m->nh, (size_t)m->nh, e->hidden, (size_t)e->ldh, (size_t)m->nh,
                                        (size_t)e->R, cudaMemcpyDeviceToDevice, ctx->stream));
    return ISB_OK;
}

int bip_run_exact_device(isb_ens *e, int rule, int64_t nsteps, int fluct_mode, const double *d_Fv,
                         const double *d_Fh, uint64_t seed, uint64_t step_offset, const double *d_T,
                         int64_t steps_per_T, int64_t trace_every, double *d_E, int8_t *d_Sv, int8_t *d_Sh) {
    isb_model *m = e->model;
    isb_ctx *ctx = m->ctx;
    int64_t ntr = 0;
    for (int64_t k = 0; k < nsteps; ++k) {
        for (int layer = 1; layer >= 0; --layer) {  // hidden first (from the OLD visible), then visible
            BipHalfParams p{};
            fill_half(p, e, layer);
            p.rule = rule;
            p.fluct_mode = fluct_mode;
            p.F = layer == 1 ? d_Fh : d_Fv;
            p.nsteps = nsteps;
            p.k = k;
            p.Tsched = d_T;
            p.tscale = e->d_tscale;
            p.steps_per_T = steps_per_T;
            p.seed = seed;
            p.step_abs = step_offset + (uint64_t)k;
            dim3 grid((p.nout + BIPX_THREADS - 1) / BIPX_THREADS, (e->R + BIPX_CH - 1) / BIPX_CH);
            bip_half_kernel<false><<<grid, BIPX_THREADS, p.nin, ctx->stream>>>(p);
            ISB_CUDA(ctx, cudaGetLastError());
            e->last_launches += 1;
        }
        if ((d_E || d_Sv || d_Sh) && trace_every > 0 && (k + 1) % trace_every == 0) {
            if (d_E) {
                int rc = bip_energy_device(e, d_E + ntr * e->R);
                if (rc) return rc;
                e->last_launches += 2;
            }
            int rc = bip_snapshot(e, ntr, d_Sv, d_Sh);
            if (rc) return rc;
            ++ntr;
        }
    }
    return ISB_OK;
}

}  // namespace isb
