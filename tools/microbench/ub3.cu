// Micro-benchmark 3 (sm_100a): can TMEM serve as the operand store of the stage-1 gather?
//   tmem_mlp : tcgen05.ld bandwidth vs loads in flight per warp, load width, warps per SM
//   cp_test  : TMA 2-D tensor-map load (SWIZZLE_64B) -> smem -> tcgen05.cp.128x256b with a
//              descriptor whose start address advances by 32 B per copy, so that TMEM lane R
//              ends up holding the CONTIGUOUS window E[8R .. 8R+28) ("Toeplitz rows": a
//              dynamic column address then is a time shift).  Prints mismatch counts per
//              descriptor variant, then the copy throughput alone and under concurrent loads.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, \
                   __LINE__);                                                          \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

template <int X>
struct Regs {
    uint32_t v[X];
};

template <int X>
__device__ __forceinline__ void tmem_ld(uint32_t (&v)[X], uint32_t addr);
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t (&r)[8], uint32_t addr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                   "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t (&r)[16], uint32_t addr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,"
        "%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
          "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t (&r)[32], uint32_t addr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,"
        "%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
          "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(addr));
}
#define TMEM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t *slot) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            smem_u32(slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t base) {
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

// ---------------------------------------------------------------- tmem_mlp
template <int X, int INFL>
__global__ void __launch_bounds__(512, 1)
k_tmem_mlp(unsigned long long *cyc, double *sink, int iters, int col_step) {
    __shared__ uint32_t slot;
    const uint32_t base = tmem_alloc_all(&slot);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tbase = base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t col = (uint32_t)(warp * 3);
    double acc[4] = {0, 0, 0, 0};
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t v[INFL][X];
#pragma unroll
        for (int u = 0; u < INFL; ++u) {
            tmem_ld<X>(v[u], tbase + col);
            col = (col + col_step) & 255u;
        }
        TMEM_WAIT_LD();
#pragma unroll
        for (int u = 0; u < INFL; ++u)
#pragma unroll
            for (int k = 0; k < X / 2; ++k)
                acc[k & 3] = fma(__hiloint2double((int)v[u][2 * k + 1], (int)v[u][2 * k]),
                                 1.0000001, acc[k & 3]);
    }
    const long long t1 = clock64();
    if (acc[0] + acc[1] + acc[2] + acc[3] == 123.456) sink[0] = acc[0];
    if (lane == 0) atomicMax(cyc + blockIdx.x, (unsigned long long)(t1 - t0));
    tmem_free_all(base);
}

// ---------------------------------------------------------------- cp_test (SW128, 16 bins per lane)
constexpr int kRowB = 128;                    // bytes per tensor row = swizzle span
constexpr int kBoxRows = 132;                 // 128 lanes + halo rows
constexpr int kStageBytes = kBoxRows * kRowB; // 16896
constexpr int kCopies = 7;                    // 7 x 32 B = 28 doubles per lane
constexpr int kDepth = 4;                     // records in flight in the rate test

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int layout, uint32_t lbo,
                                              uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                                       // descriptor version (sm_100)
    d |= (uint64_t)layout << 61;
    return d;
}

// mode 0: correctness (one CTA, dumps TMEM), mode 1: copy throughput, mode 2: copies + loads
// shape 0: 128x256b (32 B per lane per copy), shape 1: 128x128b (16 B per lane per copy)
__global__ void __launch_bounds__(288, 1)
k_cp_test(const __grid_constant__ CUtensorMap tmap, double *out, unsigned long long *cyc,
          double *sink, int joff, int row0, int iters, int mode, int shape) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar_tma, bar_cp[kDepth];
    const uint32_t base = tmem_alloc_all(&slot);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *stage = smem;
    if (threadIdx.x == 0) {
        mbar_init(&bar_tma, 1);
        for (int i = 0; i < kDepth; ++i) mbar_init(&bar_cp[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_tma, kStageBytes);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(stage)),
            "l"(&tmap), "r"(0), "r"(row0), "r"(smem_u32(&bar_tma))
            : "memory");
    }
    mbar_wait(&bar_tma, 0);
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % kDepth;
                if (it >= kDepth) mbar_wait(&bar_cp[s], ((it / kDepth) - 1) & 1);
                const uint32_t cbase = base + (uint32_t)(s * 64);
                const uint32_t src = smem_u32(stage) + 16u * joff;
                if (shape == 0) {
#pragma unroll
                    for (int m = 0; m < kCopies; ++m) {
                        const uint64_t desc = make_desc(src + 32u * m, 2, 16, 1024);
                        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(cbase + 8u * m),
                                     "l"(desc) : "memory");
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 2 * kCopies; ++m) {
                        const uint64_t desc = make_desc(src + 16u * m, 2, 16, 1024);
                        asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(cbase + 4u * m),
                                     "l"(desc) : "memory");
                    }
                }
                asm volatile(
                    "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                        smem_u32(&bar_cp[s])) : "memory");
            }
            for (int it = (iters > kDepth ? iters - kDepth : 0); it < iters; ++it)
                mbar_wait(&bar_cp[it % kDepth], (it / kDepth) & 1);
        }
    } else if (mode == 2) {
        // 8 consumer warps stream x32 loads (2 in flight) from the upper 256 columns
        const uint32_t tbase = base + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        uint32_t col = (uint32_t)(warp * 3);
        double acc[4] = {0, 0, 0, 0};
        for (int i = 0; i < iters * 2; ++i) {
            uint32_t v[2][32];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                tmem_ld<32>(v[u], tbase + col);
                col = (col + 2) & 127u;
            }
            TMEM_WAIT_LD();
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    acc[k & 3] = fma(__hiloint2double((int)v[u][2 * k + 1], (int)v[u][2 * k]),
                                     1.0000001, acc[k & 3]);
        }
        if (acc[0] + acc[1] + acc[2] + acc[3] == 123.456) sink[0] = acc[0];
    }
    const long long t1 = clock64();
    if (lane == 0) atomicMax(cyc + 2 * blockIdx.x + (warp == 8 ? 0 : 1), (unsigned long long)(t1 - t0));
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (mode == 0 && warp < 4 && blockIdx.x == 0) {
        const uint32_t tbase = base + ((uint32_t)(warp * 32) << 16);
        uint32_t v[16];
        for (int c0 = 0; c0 < 64; c0 += 16) {
            tmem_ld<16>(v, tbase + c0);
            TMEM_WAIT_LD();
            for (int k = 0; k < 8; ++k)
                out[(warp * 32 + lane) * 32 + c0 / 2 + k] =
                    __hiloint2double((int)v[2 * k + 1], (int)v[2 * k]);
        }
    }
    tmem_free_all(base);
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    unsigned long long *cyc;
    double *sink;
    CK(cudaMalloc(&cyc, sizeof(unsigned long long) * 2 * sms));
    CK(cudaMalloc(&sink, 1024));
    const int n_rows = 4096;                       // rows of 16 doubles
    double *e, *out;
    CK(cudaMalloc(&e, sizeof(double) * 16 * n_rows));
    CK(cudaMalloc(&out, sizeof(double) * 128 * 32));
    {
        double *h = (double *)malloc(sizeof(double) * 16 * n_rows);
        for (int i = 0; i < 16 * n_rows; ++i) h[i] = 1000.0 + i;
        CK(cudaMemcpy(e, h, sizeof(double) * 16 * n_rows, cudaMemcpyHostToDevice));
        free(h);
    }
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("{\"error\": \"no cuTensorMapEncodeTiled\"}\n"); return 1; }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    cuuint64_t dims[2] = {16, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {16, (cuuint32_t)kBoxRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, e, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"error\": \"encode failed %d\"}\n", (int)r); return 1; }
    const int smem_bytes = kStageBytes + 1024;
    CK(cudaFuncSetAttribute(k_cp_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    double *hout = (double *)malloc(sizeof(double) * 128 * 32);
    for (int shape = 0; shape < 2; ++shape)
        for (int joff = 0; joff < 8; joff += (shape ? 7 : 1)) {
            const int row0 = 3;
            CK(cudaMemset(out, 0, sizeof(double) * 128 * 32));
            k_cp_test<<<1, 288, smem_bytes>>>(tmap, out, cyc, sink, joff, row0, 1, 0, shape);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) {
                printf("{\"bench\": \"cp_toeplitz128\", \"joff\": %d, \"error\": \"%s\"}\n", joff,
                       cudaGetErrorString(err));
                return 1;
            }
            CK(cudaMemcpy(hout, out, sizeof(double) * 128 * 32, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int lane = 0; lane < 128; ++lane)
                for (int k = 0; k < 28; ++k)
                    if (hout[lane * 32 + k] != 1000.0 + 16.0 * (row0 + lane) + 2.0 * joff + k) ++bad;
            printf("{\"bench\": \"cp_toeplitz128\", \"shape\": %d, \"joff\": %d, \"mismatch\": %d, \"of\": %d}\n",
                   shape, joff, bad, 128 * 28);
            if (bad)
                for (int lane = 0; lane < 10; ++lane) {
                    printf("  lane %3d:", lane);
                    for (int k = 0; k < 28; ++k) printf(" %.0f", hout[lane * 32 + k] - 1000.0 - 16.0 * row0);
                    printf("\n");
                }
            fflush(stdout);
        }
    for (int shape = 0; shape < 2; ++shape)
        for (int mode = 1; mode <= 2; ++mode) {
            const int iters = 2000;
            k_cp_test<<<sms, 288, smem_bytes>>>(tmap, out, cyc, sink, 1, 0, 10, mode, shape);
            CK(cudaMemset(cyc, 0, sizeof(unsigned long long) * 2 * sms));
            k_cp_test<<<sms, 288, smem_bytes>>>(tmap, out, cyc, sink, 1, 0, iters, mode, shape);
            CK(cudaDeviceSynchronize());
            unsigned long long h[512];
            CK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * 2 * sms, cudaMemcpyDeviceToHost));
            unsigned long long mc = 0, ml = 0;
            for (int i = 0; i < sms; ++i) { mc = h[2 * i] > mc ? h[2 * i] : mc; ml = h[2 * i + 1] > ml ? h[2 * i + 1] : ml; }
            printf("{\"bench\": \"cp_rate\", \"shape\": \"%s\", \"depth\": %d, \"with_loads\": %d, "
                   "\"cp_cycles_per_record\": %.1f, \"cp_B_per_clk_sm\": %.1f, \"ld_B_per_clk_sm\": %.1f}\n",
                   shape ? "128x128b" : "128x256b", kDepth, mode - 1, (double)mc / iters,
                   (double)iters * kCopies * 128 * 32 / mc,
                   mode == 2 ? (double)iters * 2 * 8 * 2 * 32 * 4 * 32 / ml : 0.0);
        }
    return 0;
}
