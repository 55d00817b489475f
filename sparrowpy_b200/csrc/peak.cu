// Live roofline denominators: the FMA-pipe peak of the device the bench runs on.
// SURVEY.md 8(d) asks for the exchange and bake kernels as a fraction of the *measured*
// FP64 / FP32 pipe, not of a data-sheet number; bench.py times this kernel with CUDA
// events next to the kernels it reports on.
#include "common.cuh"

namespace spb {

// kChains independent FMA chains per thread (enough to cover the pipe latency at 8 warps
// per scheduler), no memory traffic.
template <typename T, int kChains>
__global__ void __launch_bounds__(256) k_fma_peak(T *out, int64_t iters, T a, T b) {
    T acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc[k] = (T)(threadIdx.x + k);
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) acc[k] = fma(acc[k], a, b);
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += acc[k];
    if (s == (T)123.456) out[0] = s;          // never true: keeps the chains alive
}

}  // namespace spb

using namespace spb;

extern "C" int spb_fma_peak(int dtype, int64_t iters, int64_t n_blocks, void *scratch,
                            double *flops_launched, void *stream) {
    SPB_REQUIRE(scratch && flops_launched, "null pointer");
    SPB_REQUIRE(iters > 0 && n_blocks > 0 && n_blocks <= 1 << 20, "iters / n_blocks");
    SPB_REQUIRE(dtype == SPB_F64 || dtype == SPB_F32, "dtype");
    constexpr int kChains = 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SPB_F64)
        k_fma_peak<double, kChains><<<(unsigned)n_blocks, 256, 0, st>>>((double *)scratch, iters,
                                                                       1.0000001, 1e-9);
    else
        k_fma_peak<float, kChains><<<(unsigned)n_blocks, 256, 0, st>>>((float *)scratch, iters,
                                                                      1.0000001f, 1e-9f);
    *flops_launched = 2.0 * kChains * 256.0 * (double)n_blocks * (double)iters;
    return check_launch("k_fma_peak");
}
