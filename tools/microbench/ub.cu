// Micro-benchmarks behind the roofline denominators of DESIGN.md (sm_100a):
//   dfma   : FP64 FMA pipe peak (independent chains)
//   dmma   : mma.sync m8n8k4 f64 rate
//   lds    : conflict-free LDS.64 operand bandwidth (the roof k_gather_tma sits on)
//   tmem   : tcgen05.ld TMEM->RF bandwidth at dynamic (also odd) column addresses,
//            alone and together with LDS (are the two operand paths independent?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ub ub.cu ; run: ./ub
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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

// ------------------------------------------------------------------ dfma
template <int CH>
__global__ void k_dfma(double *out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CH; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += acc[k];
    if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void k_ffma(float *out, int iters, float a, float b) {
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CH; ++k) acc[k] = fmaf(acc[k], a, b);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += acc[k];
    if (s == 123.456f) out[0] = s;
}

// ------------------------------------------------------------------ dmma
__global__ void k_dmma(double *out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile(
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c[k][0]), "+d"(c[k][1])
                : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    if (s == 123.456) out[0] = s;
}

// ------------------------------------------------------------------ lds / tmem
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

#define TMEM_LD16(r, addr)                                                                 \
    asm volatile(                                                                          \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,"   \
        "%12,%13,%14,%15}, [%16];"                                                         \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),          \
          "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),        \
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                               \
        : "r"(addr))
#define TMEM_ST16(r, addr)                                                                 \
    asm volatile(                                                                          \
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,"    \
        "%11,%12,%13,%14,%15,%16};" ::"r"(addr),                                           \
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),       \
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),   \
        "r"(r[14]), "r"(r[15])                                                             \
        : "memory")
#define TMEM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define TMEM_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")

// mode bit 0: tcgen05.ld stream, bit 1: LDS.64 stream, bit 2: DFMA on the loaded values
// Each iteration issues 4 x (x16 TMEM loads = 8 doubles per lane) and/or 4 x (8 LDS.64).
__global__ void __launch_bounds__(512, 1)
k_operand(unsigned long long *cyc, uint32_t *bad, double *sink, int iters, int mode,
          int col_step) {
    extern __shared__ __align__(16) double sh[];          // 4096 doubles
    __shared__ uint32_t tmem_base_sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = i;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            smem_u32(&tmem_base_sh)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base_sh + ((uint32_t)((warp & 3) * 32) << 16);
    // fill: column c of lane l holds (l << 16) | c  (warps 0..3 only: one per lane quarter)
    if (warp < 4) {
        for (int c0 = 0; c0 < 512; c0 += 16) {
            uint32_t v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = ((uint32_t)((warp & 3) * 32 + lane) << 16) | (c0 + k);
            TMEM_ST16(v, tbase + c0);
        }
        TMEM_WAIT_ST();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // correctness at an odd, dynamic column
    {
        uint32_t v[16];
        const uint32_t c0 = (uint32_t)(col_step * 3 + 1 + 2 * warp);   // odd when col_step even
        TMEM_LD16(v, tbase + c0);
        TMEM_WAIT_LD();
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (v[k] != (((uint32_t)((warp & 3) * 32 + lane) << 16) | (c0 + k))) atomicAdd(bad, 1);
    }
    __syncthreads();
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t x = 0;
    uint32_t col = (uint32_t)(warp * 3);
    const double *row = sh + lane;
    const long long t_start = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t v[16];
            double d[8];
            if (mode & 1) {
                TMEM_LD16(v, tbase + col);
                col = (col + col_step) & 255u;          // stay inside 512 columns (col + 16 <= 272)
            }
            if (mode & 2) {
                const int off = (int)((col * 5 + u * 7 + i) & 1023);
#pragma unroll
                for (int k = 0; k < 8; ++k) d[k] = row[off + 32 * k];
            }
            if (mode & 1) {
                TMEM_WAIT_LD();
                if (mode & 4) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        acc[k] = fma(__hiloint2double((int)v[2 * k + 1], (int)v[2 * k]), 1.0000001,
                                     acc[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) x ^= v[k];
                }
            }
            if (mode & 2) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = (mode & 4) ? fma(d[k], 1.0000001, acc[k]) : acc[k] + d[k];
            }
        }
    }
    const long long t_end = clock64();
    double s = x;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == 123.456) sink[0] = s;
    if (lane == 0) atomicMax(cyc + blockIdx.x, (unsigned long long)(t_end - t_start));
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_sh));
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d}\n", p.name, sms, clk_khz / 1000);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double *dout;
    CK(cudaMalloc(&dout, 1024));

    // ---- FP64 / FP32 FMA peak
    for (int rep = 0; rep < 2; ++rep) {
        const int iters = 8192, threads = 256, blocks = sms * 8;
        k_dfma<8><<<blocks, threads>>>(dout, 64, 1.0000001, 1e-9);
        CK(cudaEventRecord(e0));
        k_dfma<8><<<blocks, threads>>>(dout, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        const double fma = (double)blocks * threads * 8.0 * iters;
        const float ms = time_ms(e0, e1);
        if (rep) printf("{\"bench\": \"dfma\", \"ms\": %.3f, \"tfma_s\": %.3f, \"tflops\": %.3f}\n", ms,
                        fma / ms * 1e-9, 2 * fma / ms * 1e-9);
    }
    for (int rep = 0; rep < 2; ++rep) {
        const int iters = 8192, threads = 256, blocks = sms * 8;
        CK(cudaEventRecord(e0));
        k_ffma<8><<<blocks, threads>>>((float *)dout, iters, 1.0000001f, 1e-9f);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        const double fma = (double)blocks * threads * 8.0 * iters;
        const float ms = time_ms(e0, e1);
        if (rep) printf("{\"bench\": \"ffma\", \"ms\": %.3f, \"tfma_s\": %.3f, \"tflops\": %.3f}\n", ms,
                        fma / ms * 1e-9, 2 * fma / ms * 1e-9);
    }
    for (int rep = 0; rep < 2; ++rep) {
        const int iters = 2048, threads = 256, blocks = sms * 8;
        CK(cudaEventRecord(e0));
        k_dmma<<<blocks, threads>>>(dout, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        const double fma = (double)blocks * (threads / 32) * 8.0 * iters * 256.0;
        const float ms = time_ms(e0, e1);
        if (rep) printf("{\"bench\": \"dmma_m8n8k4\", \"ms\": %.3f, \"tfma_s\": %.3f, \"tflops\": %.3f}\n",
                        ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
    }

    // ---- operand paths
    unsigned long long *cyc;
    uint32_t *bad;
    CK(cudaMalloc(&cyc, sizeof(unsigned long long) * sms));
    CK(cudaMalloc(&bad, 4));
    CK(cudaFuncSetAttribute(k_operand, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8));
    const int modes[] = {1, 2, 3, 5, 6, 7};
    const int steps[] = {16, 2, 1};
    for (int warps = 4; warps <= 16; warps *= 2)
        for (int mi = 0; mi < 6; ++mi)
            for (int si = 0; si < 3; ++si) {
                const int mode = modes[mi], step = steps[si];
                if (!(mode & 1) && si) continue;
                const int iters = 2048;
                CK(cudaMemset(cyc, 0, sizeof(unsigned long long) * sms));
                CK(cudaMemset(bad, 0, 4));
                k_operand<<<sms, warps * 32, 4096 * 8>>>(cyc, bad, dout, 64, mode, step);
                CK(cudaMemset(cyc, 0, sizeof(unsigned long long) * sms));
                CK(cudaEventRecord(e0));
                k_operand<<<sms, warps * 32, 4096 * 8>>>(cyc, bad, dout, iters, mode, step);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                unsigned long long h[256];
                uint32_t hb;
                CK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
                unsigned long long mx = 0;
                for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
                const double per_path = (double)warps * iters * 4.0 * 32 * 64;   // bytes per SM per path
                printf("{\"bench\": \"operand\", \"warps\": %d, \"tmem\": %d, \"lds\": %d, \"dfma\": %d, "
                       "\"col_step\": %d, \"cycles\": %llu, \"tmem_B_per_clk_sm\": %.1f, "
                       "\"lds_B_per_clk_sm\": %.1f, \"ms\": %.3f, \"odd_col_mismatch\": %u}\n",
                       warps, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, step, mx,
                       (mode & 1) ? per_path / mx : 0.0, (mode & 2) ? per_path / mx : 0.0,
                       time_ms(e0, e1), hb);
            }
    return 0;
}
