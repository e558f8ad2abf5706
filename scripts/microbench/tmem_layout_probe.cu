// Probe of the TMEM accumulator layout of tcgen05.mma for the shapes the block-Gibbs kernel could use next
// (DESIGN.md §7, item 1): where do the rows and columns of D land for
//   mode 0: cta_group::1, M = 128   (known: row r -> lane r, column n -> column n; the self-check of this tool)
//   mode 1: cta_group::1, M = 64
//   mode 2: cta_group::2, M = 256   (what bip_tc.cu uses: each CTA its 128 rows on lanes 0..127)
//   mode 3: cta_group::2, M = 128   (64 rows per CTA: on 64 lanes x N columns, or on 128 lanes x N/2 columns?)
// Method: two MMAs with K = 16 whose operands are exact in bf16 —
//   D1[r][n] = r + 1   (A[r][0] = r + 1, B[n][0] = 1)      D2[r][n] = n + 1   (A[r][0] = 1, B[n][0] = n + 1)
// into TMEM that a zero M = 128 / 256 MMA has cleared before, then every lane x column of both CTAs is read back with
// tcgen05.ld and printed as (row, column) or '.' where nothing was written.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_layout_probe tmem_layout_probe.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../isingmodel.jl_b200/csrc/common.cuh"
using namespace isb;

constexpr int N = 64;        // UMMA N of the probe
constexpr int TCOLS = 128;   // TMEM columns allocated (D1 at column 0, D2 at column 64)

__device__ __forceinline__ uint64_t desc_sw128(const void *smem_tile) {  // as bip_tc.cu: K-major, 128B swizzle, SBO 1024
    const uint32_t lo = ((smem_u32(smem_tile) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// element (row r, k) of a K-major 128B-swizzled tile whose rows hold 64 bf16 (128 bytes)
__device__ __forceinline__ void put(unsigned char *tile, int r, int k, float v) {
    const uint32_t u = __float_as_uint(v);
    const unsigned short h = (unsigned short)(u >> 16);  // the probe's values are exact in bf16
    const int off = (r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2;
    *reinterpret_cast<unsigned short *>(tile + off) = h;
}
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                     "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t *bar) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
}

// One CTA (CG = 1) or one cluster of two (CG = 2), 128 threads each.  M = rows of the whole MMA.
template <int CG, int M>
__global__ void __launch_bounds__(128, 1) probe(float *out /*[CG][128 lanes][TCOLS]*/) {
    __shared__ __align__(1024) unsigned char A[128 * 128];   // up to 128 rows x 64 bf16
    __shared__ __align__(1024) unsigned char B[64 * 128];    // N (or N / 2) rows x 64 bf16
    __shared__ __align__(1024) unsigned char Z[128 * 128];   // zeros: the clearing MMA's A operand
    __shared__ uint64_t bar[3];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
    constexpr int MROWS = M / CG;  // rows of A this CTA holds
    constexpr int NROWS = N / CG;  // rows of B this CTA holds
    for (int i = threadIdx.x; i < (int)sizeof(A); i += 128) A[i] = 0, Z[i] = 0;
    for (int i = threadIdx.x; i < (int)sizeof(B); i += 128) B[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TCOLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TCOLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    // pass 0: clear the TMEM columns with a full-height zero MMA; pass 1: D1 (row code); pass 2: D2 (column code)
    for (int pass = 0; pass < 3; ++pass) {
        __syncthreads();
        if (pass > 0) {
            for (int r = threadIdx.x; r < MROWS; r += 128) put(A, r, 0, pass == 1 ? (float)(rank * MROWS + r + 1) : 1.0f);
            for (int n = threadIdx.x; n < NROWS; n += 128) put(B, n, 0, pass == 1 ? 1.0f : (float)(rank * NROWS + n + 1));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> the MMA's async reads
        if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
        if (threadIdx.x == 0 && rank == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (pass == 0) {
                for (int c = 0; c < TCOLS; c += N)
                    mma<CG>(tmem + (uint32_t)c, desc_sw128(Z), desc_sw128(B), idesc_bf16(128 * CG, N), 0u);
            } else {
                mma<CG>(tmem + (uint32_t)((pass - 1) * N), desc_sw128(A), desc_sw128(B), idesc_bf16(M, N), 0u);
            }
            commit<CG>(&bar[pass]);
        }
        mbar_wait(&bar[pass], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // read everything back: warp w may read the lanes 32 w .. 32 w + 31
    for (int c = 0; c < TCOLS; c += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[((size_t)rank * 128 + warp * 32 + lane) * TCOLS + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if constexpr (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
    }
}

template <int CG, int M>
static void run(const char *title) {
    float *d = nullptr;
    const size_t n = (size_t)CG * 128 * TCOLS;
    cudaMalloc(&d, n * sizeof(float));
    cudaMemset(d, 0, n * sizeof(float));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CG);
    cfg.blockDim = dim3(128);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CG > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CG, M>, d);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    printf("== %s: %s\n", title, cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
    float *h = (float *)malloc(n * sizeof(float));
    cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost);
    // per CTA and lane: the row code found in the D1 columns and the runs of column codes in the D2 columns
    for (int c = 0; c < CG; ++c)
        for (int lane = 0; lane < 128; ++lane) {
            const float *row = h + ((size_t)c * 128 + lane) * TCOLS;
            int first = -1, last = -1, rcode = 0;
            bool uniform = true;
            for (int j = 0; j < N; ++j)
                if (row[j] != 0.f) {
                    if (first < 0) first = j, rcode = (int)row[j];
                    last = j;
                    uniform = uniform && (int)row[j] == rcode;
                }
            if (first < 0) {
                if (lane % 32 == 0) printf("cta %d lane %3d..: nothing written on this lane (printed once per quadrant if all empty)\n", c, lane);
                continue;
            }
            printf("cta %d lane %3d: row %3d%s in TMEM columns [%d, %d]; D columns there:", c, lane, rcode - 1, uniform ? "" : " (MIXED rows!)",
                   first, last);
            int run0 = -1, prev = -2;
            for (int j = 0; j <= N; ++j) {
                const int v = j < N ? (int)row[N + j] - 1 : -5;
                if (v != prev + 1 || j == N) {
                    if (run0 >= 0) printf(" col%d..%d=n%d..%d", run0, j - 1, (int)row[N + run0] - 1, prev);
                    run0 = (j < N && v >= 0) ? j : -1;
                }
                prev = v;
            }
            printf("\n");
        }
    free(h);
    cudaFree(d);
}

int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : -1;
    if (mode < 0 || mode == 0) run<1, 128>("cta_group::1, M = 128 (self-check: lane = row, column = n)");
    if (mode < 0 || mode == 1) run<1, 64>("cta_group::1, M = 64");
    if (mode < 0 || mode == 2) run<2, 256>("cta_group::2, M = 256");
    if (mode < 0 || mode == 3) run<2, 128>("cta_group::2, M = 128");
    return 0;
}
