// Stage 1 of the energy exchange with TMA-staged energy tiles (sm_100a).
//
//   G[c,j,b,t] = sum_{i -> j in class c} ff * E_prev[src(i), b, t - delay]
//   (reference RadiosityFast.py:1124-1143, one reflection order)
//
// A CTA owns a tile of kR = 8 neighbouring receiver patches of one BRDF class, one
// band and up to 1024 time bins.  Neighbouring receivers see almost the same
// senders with almost the same delays, so the pair list is stored as the UNION of
// the tile's senders: one record per (sender row, 32-bin delay bucket) holding the
// weight and the delay remainder of each of the 8 receivers.
//
// A producer warp streams, per record, the sender's energy window
// [t0 - dmin - 32, t0 - dmin + 1024) from HBM/L2 into a shared-memory ring with 1-D
// bulk async copies (cp.async.bulk -> SASS UBLKCP) that complete on mbarriers.
// Each of the 8 consumer warps owns a 128-bin time slice for ALL 8 receivers
// (32 accumulators per lane): per record it reads the staged window at a receiver's
// shift with conflict-free LDS and -- because neighbouring receivers mostly share
// the same delay bin -- reuses the loaded registers for every following receiver
// with the same shift.  That cuts the shared-memory operand traffic, the binding
// resource of this kernel (one shifted operand per FMA), from one
// load per FMA to one load per distinct shift.
#include "common.cuh"

namespace spb {

constexpr int kR = 8;                       // receivers per tile
constexpr int kTileT = 256;                 // granularity of T_pad (spb_exchange_layout)
constexpr int kWarpsT = 8;                  // consumer warps = time slices per CTA
constexpr int kSliceT = 128;                // time bins per consumer warp
constexpr int kChunksT = kSliceT / 32;
constexpr int kCtaT = kWarpsT * kSliceT;    // 1024 time bins per CTA
constexpr int kBucket = 32;                 // delay bucket (bins) of one record
constexpr int kWindow = kCtaT + kBucket;    // staged elements per record (max)
constexpr int kStages = 12;
template <typename T>
struct alignas(16) TileRecord {
    T w[kR];             // weight per receiver slot (0 = no pair)
    uint8_t rel[kR];     // delay - dmin, 0..31 (empty slots repeat the previous shift)
    int32_t src;         // sender row = patch * D + outgoing direction
    int32_t dmin;        // bits 0..23: bucket start (multiple of 32);
                         // bits 24..31: reload mask, bit s = slot s starts a new shift
};
static_assert(sizeof(TileRecord<double>) == 80, "record layout (f64)");
static_assert(sizeof(TileRecord<float>) == 48, "record layout (f32)");

template <typename T>
struct alignas(128) Stage {
    T window[kWindow];
    TileRecord<T> rec;
};

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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

template <typename T>
__global__ void __launch_bounds__((kWarpsT + 1) * 32, 2)
k_gather_tma(const T *__restrict__ e_prev, T *__restrict__ g,
             const int64_t *__restrict__ ent_ptr, const TileRecord<T> *__restrict__ recs,
             int64_t n_patches, int64_t n_alloc, int64_t n_blocks, int64_t n_dirs,
             int64_t b_lo, int64_t jb_lo, int64_t n_jb, int64_t n_classes, int64_t t_pad,
             int64_t ld, int64_t pad, int warps_t, const int32_t *__restrict__ cta_order) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage<T> *stages = reinterpret_cast<Stage<T> *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + sizeof(Stage<T>) * kStages);
    uint64_t *empty = full + kStages;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t n_local = n_classes * n_jb;
    const int64_t b = b_lo + blockIdx.x / n_local;
    // cta_order: launch position -> local tile, longest record lists first, so that the
    // short tiles fill the tail of the grid (the hardware hands out CTAs in index order)
    const int64_t pos = blockIdx.x % n_local;
    const int64_t loc = cta_order ? cta_order[pos] : pos;
    const int64_t c = loc / n_jb;
    const int64_t jb = jb_lo + (loc - c * n_jb);
    const int64_t tile = c * n_blocks + jb;
    const int64_t e0 = ent_ptr[tile], e1 = ent_ptr[tile + 1];
    if (e0 == e1) return;                         // no pairs: rows are never read
    // warps_t (<= 8) time slices per CTA: small problems use fewer slices per CTA and
    // more CTAs along time to fill the machine
    const int64_t t0 = (int64_t)blockIdx.y * warps_t * kSliceT;
    const int n_active = (int)min((int64_t)warps_t, (t_pad - t0) / kSliceT);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], n_active);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kWarpsT) {
        // ---------------- producer warp ----------------
        const T *band_base = e_prev + b * n_alloc * n_dirs * ld + pad + t0 - kBucket;
        const uint32_t win_bytes = (uint32_t)(sizeof(T) * (n_active * kSliceT + kBucket));
        const uint32_t tx_bytes = win_bytes + (uint32_t)sizeof(TileRecord<T>);
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t e = e0; e < e1; e += 32) {
            int32_t s = 0, dm = 0;
            if (e + lane < e1) { s = recs[e + lane].src; dm = recs[e + lane].dmin & 0xffffff; }
            const int cnt = (int)min((int64_t)32, e1 - e);
            for (int k = 0; k < cnt; ++k) {
                const int32_t sk = __shfl_sync(0xffffffffu, s, k);
                const int32_t dk = __shfl_sync(0xffffffffu, dm, k);
                if (lane == 0) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], tx_bytes);
                    bulk_g2s(stages[stage].window, band_base + (int64_t)sk * ld - dk,
                             win_bytes, &full[stage]);
                    bulk_g2s(&stages[stage].rec, recs + e + k, sizeof(TileRecord<T>),
                             &full[stage]);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < n_active) {
        // ------- consumer warps: a 128-bin time slice of all 8 receivers each -------
        T acc[kR][kChunksT];
#pragma unroll
        for (int s = 0; s < kR; ++s)
#pragma unroll
            for (int v = 0; v < kChunksT; ++v) acc[s][v] = T(0);
        const int slice = kBucket + warp * kSliceT + lane;
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t e = e0; e < e1; ++e) {
            mbar_wait(&full[stage], phase);
            const Stage<T> &st = stages[stage];
            const uint64_t rel = *reinterpret_cast<const uint64_t *>(st.rec.rel);
            const uint32_t reload = ((uint32_t)st.rec.dmin) >> 24;
            // Every slot is accumulated unconditionally (absent slots have w = 0, a
            // numerical no-op); the operands are re-read from the window only where
            // the host marked a change of shift.  x starts finite so that 0 * x = 0.
            T x[kChunksT];
#pragma unroll
            for (int v = 0; v < kChunksT; ++v) x[v] = T(0);
#pragma unroll
            for (int h = 0; h < kR; h += 4) {
                T w[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) w[s] = st.rec.w[h + s];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if ((reload >> (h + s)) & 1u) {         // warp-uniform
                        const int r = (int)((rel >> (8 * (h + s))) & 0xffu);
                        const T *row = st.window + (slice - r);
#pragma unroll
                        for (int v = 0; v < kChunksT; ++v) x[v] = row[32 * v];
                    }
#pragma unroll
                    for (int v = 0; v < kChunksT; ++v)
                        acc[h + s][v] = fma(w[s], x[v], acc[h + s][v]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
#pragma unroll
        for (int s = 0; s < kR; ++s) {
            const int64_t j = jb * kR + s;
            if (j < n_patches) {
                T *out = g + ((b * n_classes + c) * n_patches + j) * ld + pad + t0 +
                         warp * kSliceT + lane;
#pragma unroll
                for (int v = 0; v < kChunksT; ++v) out[32 * v] = acc[s][v];
            }
        }
    }
}

template <typename T>
int gather_tiled_t(const void *e_prev, void *g, const int64_t *ent_ptr, const void *recs,
                   const int32_t *cta_order, int64_t n_patches, int64_t n_alloc,
                   int64_t n_classes, int64_t n_dirs, int64_t b_lo, int64_t b_hi, int64_t j_lo,
                   int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad, cudaStream_t st) {
    const int64_t n_blocks = ceil_div(n_patches, kR);
    const int64_t jb_lo = j_lo / kR, jb_hi = ceil_div(j_hi, kR);
    const int64_t n_jb = jb_hi - jb_lo;
    const int64_t n_cta = n_classes * n_jb * (b_hi - b_lo);
    if (n_cta == 0) return 0;
    SPB_REQUIRE(n_cta <= 2147483647LL, "too many tiles for one launch");
    const size_t smem = sizeof(Stage<T>) * kStages + 2 * kStages * sizeof(uint64_t);
    // per launch: the attribute is per device, and a process may drive several
    SPB_CUDA(cudaFuncSetAttribute(k_gather_tma<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    // 2 CTAs fit per SM (296 in flight).  Fewer time slices per CTA leave consumer warps
    // idle (the ring is sized for the full window and registers cap residency at 2
    // CTAs), so the window is only narrowed when the grid would not even fill the
    // machine once (measured: 8 GPUs on C4, 1250 CTAs, 4 slices: 8.9 ms vs 5.4 ms).
    int warps_t = kWarpsT;
    while (warps_t > 2 && n_cta * ceil_div(t_pad, (int64_t)warps_t * kSliceT) < 296)
        warps_t /= 2;
    dim3 grid((unsigned)n_cta, (unsigned)ceil_div(t_pad, (int64_t)warps_t * kSliceT));
    k_gather_tma<T><<<grid, (kWarpsT + 1) * 32, smem, st>>>(
        (const T *)e_prev, (T *)g, ent_ptr, (const TileRecord<T> *)recs, n_patches, n_alloc,
        n_blocks, n_dirs, b_lo, jb_lo, n_jb, n_classes, t_pad, ld, pad, warps_t, cta_order);
    return check_launch("k_gather_tma");
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_tile_geometry(int dtype, int64_t *receivers_per_tile, int64_t *delay_bucket,
                      int64_t *record_bytes) {
    SPB_REQUIRE(dtype == SPB_F64 || dtype == SPB_F32, "dtype");
    *receivers_per_tile = kR;
    *delay_bucket = kBucket;
    *record_bytes = dtype == SPB_F64 ? sizeof(TileRecord<double>) : sizeof(TileRecord<float>);
    return 0;
}

int spb_exchange_gather_tiled(const void *e_prev, void *g, const int64_t *ent_ptr,
                              const void *recs, const int32_t *cta_order,
                              int64_t n_patches, int64_t n_alloc,
                              int64_t n_classes, int64_t n_dirs, int64_t n_bands,
                              int64_t b_lo, int64_t b_hi, int64_t j_lo, int64_t j_hi,
                              int64_t t_pad, int64_t ld, int64_t pad, int dtype,
                              void *stream) {
    SPB_REQUIRE(e_prev && g && ent_ptr, "null pointer");
    SPB_REQUIRE(0 <= j_lo && j_lo <= j_hi && j_hi <= n_patches, "receiver range");
    SPB_REQUIRE(0 <= b_lo && b_lo <= b_hi && b_hi <= n_bands, "band range");
    SPB_REQUIRE(n_alloc >= n_patches, "n_alloc < n_patches");
    SPB_REQUIRE(j_lo == j_hi || j_lo % kR == 0,
                "j_lo must be a multiple of the receiver tile (8)");
    SPB_REQUIRE(t_pad % kTileT == 0 && ld == pad + t_pad, "layout (use spb_exchange_layout)");
    SPB_REQUIRE(pad % kBucket == 0 && pad >= 2 * kBucket, "pad (use spb_exchange_layout)");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SPB_F64)
        return gather_tiled_t<double>(e_prev, g, ent_ptr, recs, cta_order, n_patches, n_alloc,
                                      n_classes,
                                      n_dirs, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
    if (dtype == SPB_F32)
        return gather_tiled_t<float>(e_prev, g, ent_ptr, recs, cta_order, n_patches, n_alloc,
                                     n_classes,
                                     n_dirs, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
    return fail(-1, "invalid argument", "dtype");
}

}  // extern "C"
