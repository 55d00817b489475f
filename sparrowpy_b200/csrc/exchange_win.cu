// Stage 1 of the energy exchange, register-window variant (sm_100a, FP64).
//
//   G[c,j,b,t] = sum_{i -> j in class c} ff * E_prev[src(i), b, t - delay]
//   (reference RadiosityFast.py:1124-1143, one reflection order)
//
// Why another kernel: k_gather_tma (exchange_tma.cu) needs one shared-memory operand
// per FMA whenever the receivers of a tile have different delay bins (scenes whose
// patches are larger than a time bin, e.g. the street canyon), and the shared-memory
// path moves 16 FP64 operands per clock per SM -- a quarter of the FP64 pipe
// (DESIGN.md 3.1).  Here a lane owns 8 CONSECUTIVE time bins instead of 4 interleaved
// ones and loads, once per sender row, the window of 8 + W consecutive energies that
// covers every delay of the tile's 8 receivers (delays of neighbouring receivers
// differ by at most the tile's diameter / (c dt), W <= 10 bins).  Each receiver then
// takes its operands from that register window at its own offset: 18 loads feed 64
// FMAs instead of 64.  The offset is a warp-uniform switch, every case a straight
// run of 8 DFMAs on statically indexed registers.
//
// A lane's window is 16-byte chunks at a 64-byte lane stride, which would be a 4-way
// bank conflict in a linear row; the sender row is therefore staged with 16-byte
// cp.async (LDGSTS) into an XOR-swizzled layout (chunk bits 0-1 ^= bits
// 3-4, the classic 64-byte swizzle) in which any 8 consecutive lanes hit 8 different
// 16-byte bank groups, for every window offset.  cp.async completion is tracked by
// the stage's mbarrier (cp.async.mbarrier.arrive.noinc).
//
// Records: one per (tile, sender row, delay window): {w[8], rel[8], src, dbase} with
// dbase even, rel = delay - dbase in [0, W], rel = 255 for an empty slot
// (exchange.build_window_records).
#include "common.cuh"
#include "win_dispatch.cuh"

namespace spb {
namespace win {

constexpr int kR = 8;                  // receivers per tile
constexpr int kCtaT = 2048;            // most time bins one CTA covers
constexpr int kMaxW = 10;              // widest delay window of a record
constexpr int kStages = 12;
constexpr int kRecPad = 128;           // bytes reserved for the staged record

struct alignas(16) WinRecord {
    double w[kR];        // weight per receiver slot
    uint8_t rel[kR];     // delay - dbase (0..W), 255 = no pair in this slot
    int32_t src;         // sender row = patch * D + outgoing direction
    int32_t dbase;       // even
};
static_assert(sizeof(WinRecord) == 80, "record layout");

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread
// have landed (.noinc: the arrival is part of the barrier's initial count)
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 16-byte chunk index -> swizzled chunk index.  LT = 8 bins per lane: lane stride 4
// chunks, bits 0-1 ^= bits 3-4; LT = 4: lane stride 2 chunks, bit 0 ^= bit 3.
template <int LT>
__device__ __forceinline__ int swz(int v) {
    return v ^ ((v >> 3) & (LT == 8 ? 3 : 1));
}

// acc[k] += w * win[k + W - rel]   (rel warp-uniform)
template <int W, int LT>
__device__ __forceinline__ void accumulate(double (&acc)[LT], double w,
                                           const double (&win)[LT + W], unsigned rel) {
#define SPB_WIN_CASE(R)                                                     \
    case R:                                                                 \
        if constexpr (R <= W) {                                             \
            _Pragma("unroll") for (int k = 0; k < LT; ++k)                  \
                acc[k] = fma(w, win[k + (W - R < 0 ? 0 : W - R)], acc[k]);  \
        }                                                                   \
        break;
    switch (rel) {
        SPB_WIN_CASE(0)
        SPB_WIN_CASE(1)
        SPB_WIN_CASE(2)
        SPB_WIN_CASE(3)
        SPB_WIN_CASE(4)
        SPB_WIN_CASE(5)
        SPB_WIN_CASE(6)
        SPB_WIN_CASE(7)
        SPB_WIN_CASE(8)
        SPB_WIN_CASE(9)
        SPB_WIN_CASE(10)
        default: break;        // empty slot
    }
#undef SPB_WIN_CASE
}

// dynamic shared memory: kStages x [window (win_stride bytes) | record (kRecPad)] | barriers
//
// No dedicated producer warp (the register file is per SM sub-partition: a ninth warp
// would cap every warp at 168 registers, and the accumulators plus the window need
// ~200): every warp copies its share of the sender row of record r + kAhead while it
// works on record r.  full[s] counts one cp.async arrival per thread, empty[s] one
// arrival per warp; a stage is refilled kStages - kAhead records after this warp
// released it, so the warps are only loosely coupled.
constexpr int kAhead = 8;              // prefetch distance in records (< kStages)

template <int W, int LT>
__global__ void __launch_bounds__(kCtaT / LT, 1)
k_gather_win(const double *__restrict__ e_prev, double *__restrict__ g,
             const int64_t *__restrict__ ent_ptr, const WinRecord *__restrict__ recs,
             int64_t n_patches, int64_t n_alloc, int64_t n_blocks, int64_t n_dirs, int64_t b_lo,
             int64_t jb_lo, int64_t n_jb, int64_t n_classes, int64_t t_pad, int64_t ld,
             int64_t pad, int n_warps, int win_stride, const int32_t *__restrict__ cta_order) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int stage_stride = win_stride + kRecPad;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stage_stride * kStages);
    uint64_t *empty = full + kStages;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t n_local = n_classes * n_jb;
    const int64_t b = b_lo + blockIdx.x / n_local;
    const int64_t pos = blockIdx.x % n_local;
    const int64_t loc = cta_order ? cta_order[pos] : pos;     // longest tiles first
    const int64_t c = loc / n_jb;
    const int64_t jb = jb_lo + (loc - c * n_jb);
    const int64_t tile = c * n_blocks + jb;
    const int64_t e0 = ent_ptr[tile];
    const int n_rec = (int)(ent_ptr[tile + 1] - e0);
    if (n_rec == 0) return;                       // no pairs: rows are never read
    constexpr int kWarpT = 32 * LT;               // time bins per warp
    const int64_t t0 = (int64_t)blockIdx.y * n_warps * kWarpT;
    const int n_active = (int)min((int64_t)n_warps, (t_pad - t0) / kWarpT);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], n_active * 32);   // one cp.async arrival per thread
            mbar_init(&empty[s], n_active);       // one release per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= n_active) return;

    // ---- copy side: this thread's 16-byte chunks of a sender-row window ----
    const double *band_base = e_prev + b * n_alloc * n_dirs * ld + pad + t0 - W;
    const int n_thr = n_active * 32;
    const int tid = threadIdx.x;
    const int n_chunks = (n_active * kWarpT + W) / 2;            // 16-byte chunks per row
    const int n_mine = (n_chunks - tid + n_thr - 1) / n_thr;     // chunks tid, tid + n_thr, ...
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t my_dst = (uint32_t)(swz<LT>(tid) << 4);   // swz(tid + 32 m) = swz(tid) + 32 m
    const WinRecord *rec0 = recs + e0;
    // src / dbase of the records: read 32 at a time (lane l holds record base + l),
    // the next batch one batch ahead of its use
    int32_t m_src = 0, m_db = 0, nx_src = 0, nx_db = 0;
    if (lane < n_rec) { m_src = rec0[lane].src; m_db = rec0[lane].dbase; }
    if (32 + lane < n_rec) { nx_src = rec0[32 + lane].src; nx_db = rec0[32 + lane].dbase; }
    int p_rec = 0, p_stage = 0;
    uint32_t p_phase = 0;
    auto issue = [&]() {           // stage the window + record of record p_rec (warp-uniform)
        const int l = p_rec & 31;
        const int32_t sk = __shfl_sync(0xffffffffu, m_src, l);
        const int32_t dk = __shfl_sync(0xffffffffu, m_db, l);
        mbar_wait(&empty[p_stage], p_phase ^ 1);
        const char *src =
            reinterpret_cast<const char *>(band_base + (int64_t)sk * ld - dk) + (tid << 4);
        const uint32_t dst = smem0 + (uint32_t)p_stage * (uint32_t)stage_stride;
        for (int q = 0; q < n_mine; ++q)
            cp_async16(dst + my_dst + (uint32_t)(q * n_thr << 4), src + ((int64_t)q * n_thr << 4));
        if (tid < (int)(sizeof(WinRecord) / 16))
            cp_async16(dst + (uint32_t)win_stride + (tid << 4),
                       reinterpret_cast<const char *>(rec0 + p_rec) + (tid << 4));
        cp_async_arrive(&full[p_stage]);
        if (++p_stage == kStages) { p_stage = 0; p_phase ^= 1; }
        ++p_rec;
        if ((p_rec & 31) == 0) {   // next batch of record headers
            m_src = nx_src; m_db = nx_db;
            nx_src = nx_db = 0;
            if (p_rec + 32 + lane < n_rec) {
                nx_src = rec0[p_rec + 32 + lane].src;
                nx_db = rec0[p_rec + 32 + lane].dbase;
            }
        }
    };
    for (int p = 0; p < kAhead && p < n_rec; ++p) issue();

    // ---- compute side: 256 time bins (8 per lane) of all 8 receivers ----
    double acc[kR][LT];
#pragma unroll
    for (int s = 0; s < kR; ++s)
#pragma unroll
        for (int k = 0; k < LT; ++k) acc[s][k] = 0.0;
    const int vbase = warp * (kWarpT / 2) + lane * (LT / 2);   // first chunk of my window
    int stage = 0;
    uint32_t phase = 0;
    for (int r = 0; r < n_rec; ++r) {
        if (p_rec < n_rec) issue();
        mbar_wait(&full[stage], phase);
        const unsigned char *sw = smem_raw + (size_t)stage * stage_stride;
        double win[LT + W];
#pragma unroll
        for (int q = 0; q < (LT + W) / 2; ++q) {
            const double2 x =
                *reinterpret_cast<const double2 *>(sw + (swz<LT>(vbase + q) << 4));
            win[2 * q] = x.x;
            win[2 * q + 1] = x.y;
        }
        const WinRecord *rec = reinterpret_cast<const WinRecord *>(sw + win_stride);
        const uint64_t rel = *reinterpret_cast<const uint64_t *>(rec->rel);
#pragma unroll
        for (int s = 0; s < kR; ++s)
            accumulate<W, LT>(acc[s], rec->w[s], win, (unsigned)((rel >> (8 * s)) & 0xffu));
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
#pragma unroll
    for (int s = 0; s < kR; ++s) {
        const int64_t j = jb * kR + s;
        if (j < n_patches) {
            double2 *out = reinterpret_cast<double2 *>(
                g + ((b * n_classes + c) * n_patches + j) * ld + pad + t0 + warp * kWarpT +
                lane * LT);
#pragma unroll
            for (int k = 0; k < LT / 2; ++k)
                out[k] = make_double2(acc[s][2 * k], acc[s][2 * k + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------
// Variant 2 (window + 200, SPB_WIN_VARIANT=2): same records, same staging layout, two
// changes aimed at what the profile of variant 1 shows (profiles/r01_k_gather_win_c4_f64.txt:
// 316 instructions per (warp, record) for 58 DFMAs, branches = half of all stall samples):
//   * the row of record r is copied by ONE warp, warp r mod n_warps, with an unrolled
//     run of LDGSTS at immediate offsets; the fixed costs (barrier wait, header, address
//     arithmetic, loop control) are paid once per n_warps records and warp instead of
//     once per record and warp.  full[s] counts the 32 lanes of the copying warp.
//   * the per-(record, receiver) dispatch is one indirect branch through a jump table
//     (win_dispatch.cuh, `brx.idx`) instead of a compare tree.
// Variant 3 (window + 300, SPB_WIN_VARIANT=3) = variant 2 with the whole record in one
// chained dispatch (CHAIN = true): the dispatch of receiver s + 1 is duplicated into
// every case of receiver s, so there is no branch back to a join point and the next
// table lookup overlaps the DFMAs of the current case.
// ---------------------------------------------------------------------------------
template <int W, bool CHAIN>
__global__ void __launch_bounds__(kCtaT / 8, 1)
k_gather_win2(const double *__restrict__ e_prev, double *__restrict__ g,
              const int64_t *__restrict__ ent_ptr, const WinRecord *__restrict__ recs,
              int64_t n_patches, int64_t n_alloc, int64_t n_blocks, int64_t n_dirs,
              int64_t b_lo, int64_t jb_lo, int64_t n_jb, int64_t n_classes, int64_t t_pad,
              int64_t ld, int64_t pad, int n_warps, int win_stride,
              const int32_t *__restrict__ cta_order) {
    constexpr int LT = 8;
    constexpr int kWarpT = 32 * LT;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int stage_stride = win_stride + kRecPad;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)stage_stride * kStages);
    uint64_t *empty = full + kStages;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t n_local = n_classes * n_jb;
    const int64_t b = b_lo + blockIdx.x / n_local;
    const int64_t pos = blockIdx.x % n_local;
    const int64_t loc = cta_order ? cta_order[pos] : pos;     // longest tiles first
    const int64_t c = loc / n_jb;
    const int64_t jb = jb_lo + (loc - c * n_jb);
    const int64_t tile = c * n_blocks + jb;
    const int64_t e0 = ent_ptr[tile];
    const int n_rec = (int)(ent_ptr[tile + 1] - e0);
    if (n_rec == 0) return;
    const int64_t t0 = (int64_t)blockIdx.y * n_warps * kWarpT;
    const int n_active = (int)min((int64_t)n_warps, (t_pad - t0) / kWarpT);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 32);              // the 32 lanes of the copying warp
            mbar_init(&empty[s], n_active);       // one release per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= n_active) return;

    // ---- copy duty: rows of the records r = warp, warp + n_active, ... ----
    const double *band_base = e_prev + b * n_alloc * n_dirs * ld + pad + t0 - W;
    const int n_chunks = (n_active * kWarpT + W) / 2;            // 16-byte chunks per row
    const int n_full = n_chunks >> 5, n_tail = n_chunks & 31;
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t lane_dst = (uint32_t)(swz<LT>(lane) << 4);    // swz(lane + 32 q) = swz(lane) + 32 q
    const WinRecord *rec0 = recs + e0;
    int duty = warp;                              // next record whose row this warp copies
    int32_t h_src = 0, h_db = 0;                  // its header, loaded one duty ahead
    if (duty < n_rec) { h_src = rec0[duty].src; h_db = rec0[duty].dbase; }
    auto copy_row = [&]() {
        const int stage = duty % kStages;
        const uint32_t parity = (uint32_t)((duty / kStages) & 1);
        mbar_wait(&empty[stage], parity ^ 1);
        const char *src =
            reinterpret_cast<const char *>(band_base + (int64_t)h_src * ld - h_db) + (lane << 4);
        const uint32_t dst = smem0 + (uint32_t)stage * (uint32_t)stage_stride;
#pragma unroll 8
        for (int q = 0; q < n_full; ++q) cp_async16(dst + lane_dst + (q << 9), src + (q << 9));
        if (lane < n_tail) cp_async16(dst + lane_dst + (n_full << 9), src + (n_full << 9));
        if (lane < (int)(sizeof(WinRecord) / 16))
            cp_async16(dst + (uint32_t)win_stride + (lane << 4),
                       reinterpret_cast<const char *>(rec0 + duty) + (lane << 4));
        cp_async_arrive(&full[stage]);
        duty += n_active;
        if (duty < n_rec) { h_src = rec0[duty].src; h_db = rec0[duty].dbase; }
    };
    while (duty < n_rec && duty < kAhead) copy_row();

    // ---- compute: 256 time bins (8 per lane) of all 8 receivers ----
    double acc[kR][LT];
#pragma unroll
    for (int s = 0; s < kR; ++s)
#pragma unroll
        for (int k = 0; k < LT; ++k) acc[s][k] = 0.0;
    const int vbase = warp * (kWarpT / 2) + lane * (LT / 2);
    int stage = 0;
    uint32_t phase = 0;
    for (int r = 0; r < n_rec; ++r) {
        if (duty < n_rec && duty <= r + kAhead) copy_row();
        mbar_wait(&full[stage], phase);
        const unsigned char *sw = smem_raw + (size_t)stage * stage_stride;
        double win[LT + W];
#pragma unroll
        for (int q = 0; q < (LT + W) / 2; ++q) {
            const double2 x =
                *reinterpret_cast<const double2 *>(sw + (swz<LT>(vbase + q) << 4));
            win[2 * q] = x.x;
            win[2 * q + 1] = x.y;
        }
        const WinRecord *rec = reinterpret_cast<const WinRecord *>(sw + win_stride);
        const uint64_t rel = *reinterpret_cast<const uint64_t *>(rec->rel);
        if constexpr (CHAIN) {
            double w[kR];
#pragma unroll
            for (int s = 0; s < kR; s += 2) {
                const double2 x = *reinterpret_cast<const double2 *>(&rec->w[s]);
                w[s] = x.x;
                w[s + 1] = x.y;
            }
            accumulate_chain<W>(acc, w, win, (unsigned)rel, (unsigned)(rel >> 32));
        } else {
#pragma unroll
            for (int s = 0; s < kR; ++s)
                accumulate_brx<W, LT>(acc[s], rec->w[s], win,
                                      (unsigned)((rel >> (8 * s)) & 0xffu));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
#pragma unroll
    for (int s = 0; s < kR; ++s) {
        const int64_t j = jb * kR + s;
        if (j < n_patches) {
            double2 *out = reinterpret_cast<double2 *>(
                g + ((b * n_classes + c) * n_patches + j) * ld + pad + t0 + warp * kWarpT +
                lane * LT);
#pragma unroll
            for (int k = 0; k < LT / 2; ++k)
                out[k] = make_double2(acc[s][2 * k], acc[s][2 * k + 1]);
        }
    }
}

template <int W, int LT, int VARIANT = 1>
int launch(const double *e_prev, double *g, const int64_t *ent_ptr, const WinRecord *recs,
           const int32_t *cta_order, int64_t n_patches, int64_t n_alloc, int64_t n_classes,
           int64_t n_dirs, int64_t b_lo, int64_t b_hi, int64_t j_lo, int64_t j_hi, int64_t t_pad,
           int64_t ld, int64_t pad, cudaStream_t st) {
    const int64_t n_blocks = ceil_div(n_patches, kR);
    const int64_t jb_lo = j_lo / kR, jb_hi = ceil_div(j_hi, kR);
    const int64_t n_jb = jb_hi - jb_lo;
    const int64_t n_cta = n_classes * n_jb * (b_hi - b_lo);
    if (n_cta == 0) return 0;
    SPB_REQUIRE(n_cta <= 2147483647LL, "too many tiles for one launch");
    // time slices: as few CTAs along time as possible (a CTA stages each sender row
    // once for all its warps), equal shares; small grids use narrower CTAs so that the
    // machine is filled at least once
    constexpr int kWarpT = 32 * LT;
    constexpr int kMaxWarps = kCtaT / kWarpT;
    const int64_t m = t_pad / kWarpT;
    int64_t n_y = ceil_div(m, (int64_t)kMaxWarps);
    int n_warps = (int)ceil_div(m, n_y);
    while (n_warps > 1 && n_cta * n_y < 148) {
        n_warps = (n_warps + 1) / 2;
        n_y = ceil_div(m, (int64_t)n_warps);
    }
    const int win_bytes = (n_warps * kWarpT + W) * (int)sizeof(double);
    const int win_stride = (win_bytes + 511) / 512 * 512;
    const size_t smem = (size_t)(win_stride + kRecPad) * kStages + 2 * kStages * sizeof(uint64_t);
    dim3 grid((unsigned)n_cta, (unsigned)n_y);
    if constexpr (VARIANT >= 2) {
        static_assert(LT == 8, "variants 2 and 3 have 8 bins per lane");
        constexpr bool kChain = VARIANT == 3;
        SPB_CUDA(cudaFuncSetAttribute(k_gather_win2<W, kChain>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gather_win2<W, kChain><<<grid, n_warps * 32, smem, st>>>(
            e_prev, g, ent_ptr, recs, n_patches, n_alloc, n_blocks, n_dirs, b_lo, jb_lo, n_jb,
            n_classes, t_pad, ld, pad, n_warps, win_stride, cta_order);
        return check_launch("k_gather_win2");
    } else {
        SPB_CUDA(cudaFuncSetAttribute(k_gather_win<W, LT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gather_win<W, LT><<<grid, n_warps * 32, smem, st>>>(
            e_prev, g, ent_ptr, recs, n_patches, n_alloc, n_blocks, n_dirs, b_lo, jb_lo, n_jb,
            n_classes, t_pad, ld, pad, n_warps, win_stride, cta_order);
        return check_launch("k_gather_win");
    }
}

}  // namespace win
}  // namespace spb

using namespace spb;

extern "C" {

int spb_window_geometry(int dtype, int64_t *receivers_per_tile, int64_t *max_window,
                        int64_t *record_bytes) {
    SPB_REQUIRE(dtype == SPB_F64, "the register-window gather is FP64 only");
    *receivers_per_tile = win::kR;
    *max_window = win::kMaxW;
    *record_bytes = sizeof(win::WinRecord);
    return 0;
}

int spb_exchange_gather_window(const void *e_prev, void *g, const int64_t *ent_ptr,
                               const void *recs, const int32_t *cta_order, int64_t n_patches,
                               int64_t n_alloc, int64_t n_classes, int64_t n_dirs,
                               int64_t n_bands, int64_t b_lo, int64_t b_hi, int64_t j_lo,
                               int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad,
                               int64_t window, int dtype, void *stream) {
    SPB_REQUIRE(e_prev && g && ent_ptr, "null pointer");
    SPB_REQUIRE(dtype == SPB_F64, "the register-window gather is FP64 only");
    SPB_REQUIRE(0 <= j_lo && j_lo <= j_hi && j_hi <= n_patches, "receiver range");
    SPB_REQUIRE(0 <= b_lo && b_lo <= b_hi && b_hi <= n_bands, "band range");
    SPB_REQUIRE(n_alloc >= n_patches, "n_alloc < n_patches");
    SPB_REQUIRE(j_lo == j_hi || j_lo % win::kR == 0,
                "j_lo must be a multiple of the receiver tile (8)");
    SPB_REQUIRE(t_pad % 256 == 0 && ld == pad + t_pad, "layout (use spb_exchange_layout)");
    SPB_REQUIRE(pad % 32 == 0 && pad >= 64, "pad (use spb_exchange_layout)");
    cudaStream_t st = (cudaStream_t)stream;
    const double *ep = (const double *)e_prev;
    const win::WinRecord *r = (const win::WinRecord *)recs;
#define SPB_WIN_LAUNCH(W_, LT_, V_)                                                          \
    return win::launch<W_, LT_, V_>(ep, (double *)g, ent_ptr, r, cta_order, n_patches,       \
                                    n_alloc, n_classes, n_dirs, b_lo, b_hi, j_lo, j_hi,      \
                                    t_pad, ld, pad, st)
    // window + 100: tuning variant with 4 bins per lane and up to 16 warps per CTA
    // window + 200: variant 2 (row copied by one warp per record, jump-table dispatch)
    // window + 300: variant 3 (variant 2 with one chained dispatch per record)
    if (window == 4) SPB_WIN_LAUNCH(4, 8, 1);
    if (window == 10) SPB_WIN_LAUNCH(10, 8, 1);
    if (window == 104) SPB_WIN_LAUNCH(4, 4, 1);
    if (window == 110) SPB_WIN_LAUNCH(10, 4, 1);
    if (window == 204) SPB_WIN_LAUNCH(4, 8, 2);
    if (window == 210) SPB_WIN_LAUNCH(10, 8, 2);
    if (window == 304) SPB_WIN_LAUNCH(4, 8, 3);
    if (window == 310) SPB_WIN_LAUNCH(10, 8, 3);
#undef SPB_WIN_LAUNCH
    return fail(-1, "invalid argument", "window must be 4 or 10");
}

}  // extern "C"
