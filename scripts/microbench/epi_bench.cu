// Micro-benchmark of the block-Gibbs sampling epilogue (bip_tc.cu) without the contraction: every epilogue warp
// runs the per-chunk work (16 units per thread: Philox4x32-10 words, logistic rule with one ex2, sign packing,
// two 16-byte stores) on register inputs.  Prints cycles per chunk per scheduler for several variants, to see
// which part of the epilogue sets its cost.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_bench epi_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../isingmodel.jl_b200/csrc/common.cuh"
using namespace isb;

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE 0: full chunk; 1: no Philox (cheap words); 2: Philox only; 3: full, u from mantissa bits (no I2F);
// 5: full, decision as a compare: (float)w (1 + e) > 2^32  (no u scaling, no 1 - s subtraction)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) epi(const float *bias, uint4 *out, int ldo16, int iters, uint64_t seed, float cE, long long *cyc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 4) return;
    const PhiloxKeys keys = philox_keys(seed);
    const int r = blockIdx.x * 128 + (warp & 3) * 32 + lane;
    float acc = (float)(r & 15) * 0.01f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int u0 = (it & 15) * 16;
        float v[16], bf[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = acc + (float)j * 0.125f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + u0) + q);
            bf[4 * q] = b4.x; bf[4 * q + 1] = b4.y; bf[4 * q + 2] = b4.z; bf[4 * q + 3] = b4.w;
        }
        Philox4 blk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (MODE == 1) blk[q] = Philox4{(uint32_t)it * 2654435761u + (uint32_t)r, (uint32_t)(u0 + q) * 40503u, (uint32_t)r << 7, keys.k0[3]};
            else blk[q] = philox4x32_10k((uint32_t)it, 0u, (uint32_t)r, (4u << 28) | (uint32_t)((u0 >> 2) + q), keys);
        }
        uint32_t wb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wb[j] = 0x3F803F80u;
        if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { wb[2 * q] ^= blk[q].x ^ blk[q].y; wb[2 * q + 1] ^= blk[q].z ^ blk[q].w; }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t w = philox_pick(blk[j >> 2], (uint32_t)(j & 3));
                const float x = v[j] + bf[j];
                if (MODE == 5) {
                    const float wf = (float)w;
                    const float sv = fmaf(wf, ex2a(cE * x), wf);
                    if (sv > 4294967296.0f) wb[j >> 1] |= 0x8000u << (16 * (j & 1));
                    continue;
                }
                float u;
                if (MODE == 3) u = __uint_as_float((w & 0x007FFFFFu) | 0x3F800000u) - 0.99999994f;
                else u = fmaf((float)w, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                const float t = 1.0f - fmaf(u, ex2a(cE * x), u);
                wb[j >> 1] |= (j & 1) ? (__float_as_uint(t) & 0x80000000u) : ((__float_as_uint(t) & 0x80000000u) >> 16);
            }
        }
        uint4 *o = out + (size_t)r * ldo16 + (u0 >> 3);
        o[0] = make_uint4(wb[0], wb[1], wb[2], wb[3]);
        o[1] = make_uint4(wb[4], wb[5], wb[6], wb[7]);
        acc += __uint_as_float((wb[0] & 0x80000000u) | 0x3A000000u);   // loop-carried, keeps the inputs live
    }
    const long long t1 = clock64();
    if (lane == 0 && warp == 4 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (acc == 12345.f) out[0].x = 1;
}


// MODE 6/7: the int8 digit-plane chunk of bip_tc.cu's fast path (three plane sums per unit recombined in int32 (6) or one
// conversion + FFMA per plane (7), compare-form decision, 16 int8 spins per store); plane sums come from registers.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) epi8(const float *bias, uint4 *out, int ldo16, int iters, uint64_t seed, float cE, long long *cyc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 4) return;
    const PhiloxKeys keys = philox_keys(seed);
    const int r = blockIdx.x * 128 + (warp & 3) * 32 + lane;
    uint32_t acc = (uint32_t)(r & 15);
    const float s0 = 65536.f * 1e-7f, s1 = 256.f * 1e-7f, s2 = 1e-7f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int u0 = (it & 15) * 16;
        float xf[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + u0) + q);
            xf[4 * q] = b4.x; xf[4 * q + 1] = b4.y; xf[4 * q + 2] = b4.z; xf[4 * q + 3] = b4.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t v0 = acc + (uint32_t)j, v1 = acc * 3u + (uint32_t)j, v2 = acc ^ (uint32_t)(j * 77);
            if (MODE == 6) xf[j] = fmaf((float)(int)((((v0 << 8) + v1) << 8) + v2), s2, xf[j]);
            else xf[j] = fmaf((float)(int)v0, s0, fmaf((float)(int)v1, s1, fmaf((float)(int)v2, s2, xf[j])));
        }
        Philox4 blk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) blk[q] = philox4x32_10k((uint32_t)it, 0u, (uint32_t)r, (4u << 28) | (uint32_t)((u0 >> 2) + q), keys);
        uint32_t wb[4] = {0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float wf = (float)philox_pick(blk[j >> 2], (uint32_t)(j & 3));
            if (fmaf(wf, ex2a(cE * xf[j]), wf) > 4294967296.0f) wb[j >> 2] |= 0xFEu << (8 * (j & 3));
        }
        out[(size_t)r * ldo16 + (u0 >> 4)] = make_uint4(wb[0], wb[1], wb[2], wb[3]);
        acc += wb[0] >> 24;   // loop-carried, keeps the inputs live
    }
    const long long t1 = clock64();
    if (lane == 0 && warp == 4 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (acc == 12345u) out[0].x = 1;
}
template <int MODE>
static void run8(const char *name, int ew, const float *bias, uint4 *out, long long *cyc) {
    const int iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    epi8<MODE><<<148, 32 * (4 + ew)>>>(bias, out, 64, 64, 1, -1.3f, cyc);
    cudaEventRecord(a);
    epi8<MODE><<<148, 32 * (4 + ew)>>>(bias, out, 64, iters, 1, -1.3f, cyc);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s epi_warps=%2d  %8.3f ms  %7.1f cycles/chunk/scheduler  (%5.1f per warp-chunk)  err=%s\n", name, ew, ms,
           (double)c / (iters * (ew / 4.0)), (double)c / iters, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE>
static void run(const char *name, int ew, const float *bias, uint4 *out, long long *cyc) {
    const int iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    epi<MODE><<<148, 32 * (4 + ew)>>>(bias, out, 64, 64, 1, -1.3f, cyc);
    cudaEventRecord(a);
    epi<MODE><<<148, 32 * (4 + ew)>>>(bias, out, 64, iters, 1, -1.3f, cyc);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // chunk-slots per scheduler = iters * (ew / 4)
    printf("%-28s epi_warps=%2d  %8.3f ms  %7.1f cycles/chunk/scheduler  (%5.1f per warp-chunk)  err=%s\n", name, ew, ms,
           (double)c / (iters * (ew / 4.0)), (double)c / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *bias; uint4 *out; long long *cyc;
    cudaMalloc(&bias, 4096); cudaMemset(bias, 0, 4096);
    cudaMalloc(&out, (size_t)148 * 128 * 64 * 16);
    cudaMalloc(&cyc, 8);
    for (int ew : {4, 8, 12, 16, 24}) {
        run8<6>("int8 chunk, int32 recombine", ew, bias, out, cyc);
        run8<7>("int8 chunk, per-plane FFMA", ew, bias, out, cyc);
        run<0>("full", ew, bias, out, cyc);
        run<1>("no philox", ew, bias, out, cyc);
        run<2>("philox only", ew, bias, out, cyc);
        run<3>("full, mantissa u", ew, bias, out, cyc);
        run<5>("full, compare form", ew, bias, out, cyc);
    }
    return 0;
}
