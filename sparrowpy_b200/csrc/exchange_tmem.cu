// Stage 1 of the energy exchange with the sender window held in TENSOR MEMORY
// (sm_100a, FP64).
//
//   G[c,j,b,t] = sum_{i -> j in class c} ff * E_prev[src(i), b, t - delay]
//   (reference RadiosityFast.py:1124-1143, one reflection order)
//
// Why: every FMA of this sum needs one operand from a *shifted* sender row, and the
// shift (the pair's delay bin) differs per (sender, receiver).  Read from shared memory
// that is one 8-byte operand per FMA at 128 B/clk/SM = a quarter of the FP64 pipe
// (k_gather_tma, measured 79 % of that roof).  Registers cannot be indexed dynamically,
// so a register window needs a branch per (record, receiver) (k_gather_win, no faster).
// Tensor memory can: tcgen05.ld takes a *dynamic column address* and returns consecutive
// columns of the thread's TMEM lane in statically named registers, and it is a separate
// datapath (measured on B200: 360-400 B/clk/SM with 8 warps, tools/microbench).  So the
// window lives in TMEM with time along the columns, and a delay is a column offset.
//
// Layout ("Toeplitz rows"): a CTA covers 8 receivers x 2048 time bins; TMEM lane R holds,
// for the record in flight, the contiguous energies E[T0(R) - dbase - H .. T0(R) - dbase
// + 16) of the sender row, T0(R) = first of the 16 consecutive bins the lane owns, H = the
// widest delay window of a record.  A receiver whose delay is dbase + rel reads its 16
// operands with two tcgen05.ld.32x32b.x16 at column 2 (H - rel).
//
// Pipeline per record (one (tile, sender row, delay window), 80 bytes, the records of
// exchange.build_window_records in the device encoding of exchange.device_window_records):
//   4 producer warps: (one per shared-memory stage, one working lane each) one 2-D TMA tensor
//                   load per band (rows of 16 doubles, SWIZZLE_128B) of the sender window,
//                   through the one of 8 phase-shifted tensor maps that makes the window
//                   start on a 16-byte chunk boundary of its row, + a bulk copy of the batch's
//                   records into a shared-memory ring that the consumers read directly
//   4 fill warps  : (one per TMEM lane quarter) read the lane's row from the swizzled
//                   stage with conflict-free LDS.128 (the swizzle makes the 128-byte lane
//                   pitch hit 8 different bank groups) at static addresses and write it with
//                   tcgen05.st into a ring of TMEM stages
//   8 consumer warps: (quarter q, receiver group g) 16 bins x 4 receivers per thread =
//                   64 FP64 accumulators; per receiver 2 tcgen05.ld + 16 DFMA.
// Quarters run decoupled (per-quarter mbarriers).  Registers are re-balanced with
// setmaxnreg (consumers 184, fill 104, producers 40).
//
// When the histogram is shorter than 2048 bins the four lane quarters are spread over
// several bands instead (same records, other energy rows), so no lane idles.
//
// Measured on B200, C4 (6.17 M records, 8.3 x 10^10 useful FMAs per launch), SPB_TMEM_DBG
// experiments: data movement + hand-overs alone 9.1 ms (103 GB from L2 = 11.3 TB/s, the
// chip's L2 throughput cap), + fill 10.5 ms, + consumers without fill 11.6 ms, everything
// 13.4 ms = 44 % of the measured FP64 FMA peak.  History: one producer warp 15.1 ms (its
// serial issue chain of ~1300 cycles per batch was the bound), per-record hand-overs 18.9 ms;
// a software-pipelined consumer loop (operands of the next receiver requested before the
// DFMAs of the current one, 208-216 consumer registers, fill warps at 56-72) measured
// 13.4-13.7 ms -- no gain, because the L2 -> SM window traffic, not the TMEM latency, is what
// the consumers wait for -- and was removed.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace spb {
namespace tmg {

constexpr int kR = 8;                          // receivers per tile
constexpr int kLaneT = 16;                     // consecutive time bins per TMEM lane
constexpr int kQuarterT = 32 * kLaneT;         // 512 bins per lane quarter
constexpr int kBoxArea = 20480;                // bytes reserved for the staged boxes of a record
constexpr int kBatch = 2;                      // records per pipeline hand-over
constexpr int kThreads = 512;   // 8 consumer + 4 fill + 4 producer warps; 512 x 128 regs
                                // at launch = the whole file, so setmaxnreg only re-deals it

struct alignas(16) WinRecord {
    double w[kR];        // weight per receiver slot
    uint8_t rel[kR];     // TMEM column offset 2 (delay - dbase); 0 (and w = 0) = no pair
                         // (device encoding, exchange.device_window_records)
    int32_t src;         // sender row = patch * D + outgoing direction
    int32_t dbase;       // even
};
static_assert(sizeof(WinRecord) == 80, "record layout");

template <int H>
struct Cfg {
    static constexpr int kRowD = kLaneT + H;           // doubles per TMEM lane row
    static constexpr int kChunks = kRowD / 2;          // 16-byte chunks per row
    static constexpr int kCols = 2 * kRowD;            // 32-bit TMEM columns per stage
    static constexpr int kTmemStages = 512 / kCols;
    static_assert(H % 2 == 0 && kRowD % 2 == 0, "rows are made of 16-byte chunks");
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
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// 16 consecutive 32-bit columns (8 doubles) of this thread's TMEM lane; asynchronous:
// tcgen05.wait::ld before the registers are read
__device__ __forceinline__ void tmem_ld16(uint32_t *r, uint32_t addr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,"
        "%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
          "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,"
        "%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(addr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
        "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,"
        "%14,%15,%16};" ::"r"(addr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
                 "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_store_row(uint32_t addr, const uint32_t *r) {
    // N 32-bit columns as the fewest power-of-two stores
    int done = 0;
    if constexpr (N >= 32) { tmem_st32(addr, r); done = 32; }
    if constexpr ((N - (N >= 32 ? 32 : 0)) >= 16) { tmem_st16(addr + done, r + done); done += 16; }
    if constexpr (((N % 16) >= 8)) { tmem_st8(addr + done, r + done); done += 8; }
    if constexpr (((N % 8) >= 4)) { tmem_st4(addr + done, r + done); done += 4; }
    static_assert(N % 4 == 0 && N < 64, "row width");
}

// ---- raw shared-memory / mbarrier accessors on 32-bit shared addresses (no generic
// address arithmetic in the inner loops)
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_f64x2(double &a, double &b, uint32_t addr) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint64_t v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// One record of the fill: the lane's row (C::kChunks 16-byte chunks) from the swizzled
// stage into the TMEM stage.  The producer picks, per record, the tensor map whose origin
// makes the lane's row start at chunk 0 of its tensor row, so the chunk addresses are
// static: p0 / p1 = address of the lane's tensor row / the next one with the row's swizzle
// key folded in (logical chunk cc of a row is at p ^ (cc << 4)).  The whole row goes through
// the registers at once: all LDS.128 are independent (one shared-memory latency per
// record), then the fewest power-of-two tcgen05.st.
template <int H>
__device__ __forceinline__ void fill_row(uint32_t p0, uint32_t p1, uint32_t taddr) {
    using C = Cfg<H>;
    uint32_t v[C::kCols];
#pragma unroll
    for (int i = 0; i < C::kChunks; ++i) {
        const uint4 x = lds128(i < 8 ? p0 ^ (uint32_t)(i << 4) : p1 ^ (uint32_t)((i - 8) << 4));
        v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    }
    tmem_store_row<C::kCols>(taddr, v);
}

// Tensor maps of the previous-order histogram seen as rows of 16 doubles, one per starting
// phase: map J has its origin at double 2 J, so a window that starts at double a is the box
// at row a >> 4 of map (a & 15) >> 1 and lands chunk-aligned in shared memory.
struct TmapSet {
    CUtensorMap m[8];
};

// Stage 2 fused into the stage-1 epilogue for scenes with one BRDF class and one direction
// (diffuse walls: E_k[j] = coef[b] * G[j]): instead of writing G, a consumer warp scales its
// accumulators, adds them to E_total and stores E_k into every rank's ping-pong buffer (the
// rank's own and, over NVLink, the peers'), so that k_mix and its exposed communication
// disappear and the stores of finished tiles overlap the tiles still running.
constexpr int kMaxPeers = 8;
constexpr int kEpiBytes = 8 * 4096;            // one 4 KB transpose buffer per consumer warp
struct FuseArgs {
    double *cur[kMaxPeers];    // E_k buffers of all ranks (n == 0: not fused, write G)
    double *total;             // E_total of this rank
    const double *coef;        // (1, 1, B)
    int n;
};

// Register deal per role (setmaxnreg); launch = 16 warps x 128.
struct Regs {
    static constexpr int kConsumer = 184;
    static constexpr int kFill = 104;
    static constexpr int kProducer = 40;
    // setmaxnreg.inc can only take what setmaxnreg.dec released inside the CTA (a pool that
    // starts empty): the re-deal must not need more registers than the launch allocated
    static_assert(8 * kConsumer + 4 * kFill + 4 * kProducer <= 16 * 128,
                  "register re-deal exceeds the launch allocation (the kernel would hang)");
};

template <int H, int B>
__global__ void __launch_bounds__(kThreads, 1)
k_gather_tmem(const __grid_constant__ TmapSet tmaps, const __grid_constant__ FuseArgs fuse,
              double *__restrict__ g,
              const int64_t *__restrict__ ent_ptr, const WinRecord *__restrict__ recs,
              const int32_t *__restrict__ cta_order, int64_t n_patches, int64_t n_alloc,
              int64_t n_blocks, int64_t n_dirs, int64_t b_lo, int64_t b_hi, int64_t jb_lo,
              int64_t n_jb, int64_t n_classes, int64_t t_pad, int64_t ld, int64_t pad, int qpb,
              int n_tchunks, int dbg) {
    // Every hand-over of the pipeline (producer -> fill -> consumers) moves a BATCH of B
    // records: the barrier waits, fences and arrivals are a serial chain of a few hundred
    // cycles per hand-over in every role, and per single record that chain, not any
    // bandwidth, bounded the kernel (measured: 480 clk per record with no data moved at all).
    using C = Cfg<H>;
    using R = Regs;
    constexpr int S = 8 / B;                           // shared-memory stages (of B records)
    constexpr int TS = 512 / (B * C::kCols);           // TMEM stages (of B rows per lane)
    // The records themselves go into a ring of S + TS batches that the consumers read
    // directly: when the producer refills slot n mod (S + TS) it has seen the fill warps
    // release the shared-memory stage of batch n - S, which they did after the consumers
    // released the TMEM stage of batch n - S - TS.
    constexpr int RS = S + TS;
    constexpr int kStage = B * kBoxArea;               // staged boxes of one batch
    constexpr int kRing = RS * B * (int)sizeof(WinRecord);
    static_assert(TS >= 2 && S >= 2, "double buffering");
    static_assert(kRing % 16 == 0, "ring alignment");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // layout: S x kStage | record ring | epilogue transpose buffers | barriers | tmem slot
    const uint32_t sm_stages = smem_u32(smem_raw);
    const uint32_t sm_recs = sm_stages + S * kStage;
    const uint32_t sm_epi = sm_recs + kRing;                          // 8 x 4 KB (fused epilogue)
    const uint32_t sm_full = sm_epi + kEpiBytes;                      // smem_full[S]
    const uint32_t sm_empty = sm_full + 8 * S;                        // smem_empty[S]
    const uint32_t tm_full = sm_empty + 8 * S;                        // tmem_full[TS][4]
    const uint32_t tm_empty = tm_full + 8 * TS * 4;                   // tmem_empty[TS][4]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(
        smem_raw + S * kStage + kRing + kEpiBytes + 16 * S + 64 * TS);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // ---- which tile, which bands / time chunk ----
    const int64_t pos = blockIdx.x;
    const int64_t loc = cta_order ? cta_order[pos] : pos;
    const int64_t c = loc / n_jb;
    const int64_t jb = jb_lo + (loc - c * n_jb);
    const int64_t tile = c * n_blocks + jb;
    const int64_t e0 = ent_ptr[tile];
    const int n_bat = (int)((ent_ptr[tile + 1] - e0) / B);   // record lists are padded to B
    const int bands_per_cta = 4 / qpb;
    const int64_t bg = blockIdx.y / n_tchunks;
    const int64_t tc = blockIdx.y - bg * n_tchunks;
    const int64_t band0 = b_lo + bg * bands_per_cta;
    const int64_t t_base = tc * (int64_t)kQuarterT * qpb;   // first bin of the CTA's chunk
    // quarter q: band band0 + q / qpb, bins [t_base + (q % qpb) * 512, + 512)
    auto q_band = [&](int q) { return band0 + q / qpb; };
    auto q_t0 = [&](int q) { return t_base + (int64_t)(q % qpb) * kQuarterT; };
    auto q_active = [&](int q) { return q_band(q) < b_hi && q_t0(q) < t_pad; };
    int n_act_q = 0, n_act_bands = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) n_act_q += q_active(q) ? 1 : 0;
    for (int s = 0; s < bands_per_cta; ++s) n_act_bands += (band0 + s < b_hi) ? 1 : 0;
    const int box_rows = 32 * qpb + 1;
    const int box_stride = (box_rows * 128 + 1023) & ~1023;
    if (n_bat == 0) {
        // no pairs.  Not fused: the G rows are never read.  Fused: E_k of these receivers is
        // zero and must overwrite what the ping-pong buffers hold from two orders ago
        if (fuse.n > 0) {
            const double2 zero = make_double2(0.0, 0.0);
            for (int idx = threadIdx.x; idx < 4 * kR * (kQuarterT / 2); idx += kThreads) {
                const int q = idx / (kR * (kQuarterT / 2));
                const int rem = idx - q * (kR * (kQuarterT / 2));
                const int64_t j = jb * kR + rem / (kQuarterT / 2);
                const int64_t t = q_t0(q) + 2 * (rem % (kQuarterT / 2));
                if (!q_active(q) || j >= n_patches || t >= t_pad) continue;
                const int64_t o = (q_band(q) * n_alloc + j) * ld + pad + t;
                for (int p = 0; p < fuse.n; ++p) *reinterpret_cast<double2 *>(fuse.cur[p] + o) = zero;
            }
        }
        return;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_full + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_empty + 8 * s),
                         "r"(n_act_q));
        }
        for (int s = 0; s < TS * 4; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tm_full + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tm_empty + 8 * s), "r"(2));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 12) {
        // ---------------- producer warps (one working lane each) ----------------
        // Producer warp p feeds shared-memory stage p: issuing the copies of one batch is a
        // serial chain of ~1300 cycles in a single thread (barrier wait, expect_tx, address
        // arithmetic, UTMALDG through the uniform datapath), which with ONE producer warp was
        // what bounded the whole kernel; four of them run these chains concurrently.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::kProducer));
        static_assert(S == 4, "one producer warp per shared-memory stage");
        const int p = warp - 12;
        const uint32_t tx_bytes =
            (uint32_t)(B * (n_act_bands * box_rows * 128 + (int)sizeof(WinRecord)));
        // first double of lane 0's row = a_base + band slot * a_band + src * ld - dbase
        const int64_t a_band = n_alloc * n_dirs * ld;
        const int64_t a_base = band0 * a_band + pad + t_base - H;
        const WinRecord *rp = recs + e0;
        const int n_rec = n_bat * B;
        const uint32_t full = sm_full + 8u * p, empty = sm_empty + 8u * p;
        const uint32_t st0 = sm_stages + (uint32_t)p * kStage;
        uint32_t phase = 0;
        for (int e = 0; e < n_rec; e += 32) {
            // (src, dbase) of the next 32 records, one per lane (all four warps read them)
            int32_t s = 0, db = 0;
            if (e + lane < n_rec) { s = rp[e + lane].src; db = rp[e + lane].dbase; }
            const int cnt = min(32, n_rec - e);
            for (int k = p * B; k < cnt; k += 4 * B) {     // batches e / B + p, + 4, ...
                if (lane == 0) {
                    mbar_wait_a(empty, phase ^ 1);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                                     full), "r"(tx_bytes) : "memory");
                }
#pragma unroll
                for (int r = 0; r < B; ++r) {
                    const int32_t sk = __shfl_sync(0xffffffffu, s, k + r);
                    const int32_t dk = __shfl_sync(0xffffffffu, db, k + r);
                    if (lane == 0) {
                        const uint32_t st = st0 + r * kBoxArea;
                        int64_t a0 = a_base + (int64_t)sk * ld - dk;
                        // every term but dbase + H is a multiple of 16: one map for all bands
                        const CUtensorMap *map = &tmaps.m[((uint32_t)a0 & 15u) >> 1];
                        for (int bs = 0; bs < n_act_bands; ++bs, a0 += a_band) {
                            const int32_t row0 = (int32_t)(a0 >> 4);       // tensor row (floor)
                            asm volatile(
                                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
                                "[%0], [%1, {%2, %3}], [%4];" ::"r"(st + bs * box_stride),
                                "l"(map), "r"(0), "r"(row0), "r"(full)
                                : "memory");
                        }
                    }
                }
                if (lane == 0) {
                    const int rs = ((e + k) / B) % RS;         // ring slot of this batch
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(sm_recs + (uint32_t)(rs * B * (int)sizeof(WinRecord))), "l"(rp + e + k),
                        "r"((uint32_t)(B * sizeof(WinRecord))), "r"(full)
                        : "memory");
                }
                phase ^= 1;
            }
        }
    } else if (warp >= 8) {
        // ---------------- fill warps: smem stage -> TMEM stage ----------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R::kFill));
        const int q = warp - 8;
        if (q_active(q)) {
            const int rb = (q % qpb) * 32 + lane;           // the lane's row inside its box
            // offsets of the lane's tensor row and the next one inside a box, swizzle key
            // folded in
            const uint32_t off0 = (uint32_t)((q / qpb) * box_stride + rb * 128 + ((rb & 7) << 4));
            const uint32_t off1 =
                (uint32_t)((q / qpb) * box_stride + (rb + 1) * 128 + (((rb + 1) & 7) << 4));
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
            uint32_t st = sm_stages, sfull = sm_full;            // smem_empty = sfull + 8 S
            uint32_t tfull = tm_full + 8 * q;                    // tmem_empty = tfull + 32 TS
            uint32_t tcol = trow;
            int ss = 0, ts = 0;
            uint32_t sphase = 0, tphase = 1;
            for (int e = n_bat; e > 0; --e) {
                mbar_wait_a(sfull, sphase);
                mbar_wait_a(tfull + 32 * TS, tphase);
                tc_fence_after();
                if (!(dbg & 2))
#pragma unroll
                for (int r = 0; r < B; ++r)
                    fill_row<H>(st + r * kBoxArea + off0, st + r * kBoxArea + off1,
                                               tcol + r * C::kCols);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_a(sfull + 8 * S);
                    mbar_arrive_a(tfull);
                }
                st += kStage; sfull += 8;
                if (++ss == S) { ss = 0; sphase ^= 1; st = sm_stages; sfull = sm_full; }
                tfull += 32; tcol += B * C::kCols;
                if (++ts == TS) { ts = 0; tphase ^= 1; tfull = tm_full + 8 * q; tcol = trow; }
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R::kConsumer));
        const int q = warp & 3, grp = warp >> 2;
        if (q_active(q)) {
            double acc[4][kLaneT];
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int k = 0; k < kLaneT; ++k) acc[s][k] = 0.0;
            const uint32_t col0 = tmem_base + ((uint32_t)(q * 32) << 16) + 2 * H;
            // this group's part of a record in the ring: w[4 grp ..] and the 4 column offsets
            // (2 x (delay - dbase), 0 for an empty slot) of its receivers
            const uint32_t rec0 = sm_recs + 32 * grp;
            uint32_t col = col0, rec = rec0;
            uint32_t tfull = tm_full + 8 * q;                          // tmem_empty = + 32 TS
            int ts = 0, rs = 0;
            uint32_t tphase = 0;
            // address of the 4 offset bytes of record r: base of the record + 64 + 4 grp
            // = (rec + r * 80) - 32 grp + 64 + 4 grp
            auto off_addr = [&](int r) { return rec + r * (int)sizeof(WinRecord) + 64 - 28 * grp; };
            {
                // Per receiver: one x32 load, 16 DFMAs; ptxas overlaps the next receiver's
                // load with the DFMAs as far as the registers allow, and the two consumer
                // warps of a scheduler overlap each other's TMEM latency.  Empty slots have
                // w = 0 and read the window at shift 0 (finite energies, a numerical no-op):
                // no branches.
                for (int n = n_bat; n > 0; --n) {
                    mbar_wait_a(tfull, tphase);
                    tc_fence_after();
#pragma unroll
                    for (int r = 0; r < B; ++r) {
                        double w[4];
                        lds_f64x2(w[0], w[1], rec + r * (int)sizeof(WinRecord));
                        lds_f64x2(w[2], w[3], rec + r * (int)sizeof(WinRecord) + 16);
                        const uint32_t off4 = lds32(off_addr(r));
                        const uint32_t cr = col + r * C::kCols;
                        if (!(dbg & 1))
#pragma unroll
                        for (int s2 = 0; s2 < 4; ++s2) {
                            uint32_t x[32];
                            const uint32_t adr = cr - ((off4 >> (8 * s2)) & 0xffu);
                            tmem_ld16(x, adr);
                            tmem_ld16(x + 16, adr + 16);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int k = 0; k < kLaneT; ++k)
                                acc[s2][k] = fma(w[s2],
                                                 __hiloint2double((int)x[2 * k + 1], (int)x[2 * k]),
                                                 acc[s2][k]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(tfull + 32 * TS);
                    tfull += 32; col += B * C::kCols; rec += B * (int)sizeof(WinRecord);
                    if (++ts == TS) { ts = 0; tphase ^= 1; tfull = tm_full + 8 * q; col = col0; }
                    if (++rs == RS) { rs = 0; rec = rec0; }
                }
            }
            // ---- epilogue: 16 consecutive bins per receiver row ----
            const int64_t b = q_band(q);
            const int64_t t0 = q_t0(q) + (int64_t)lane * kLaneT;
            if (fuse.n == 0) {
                if (t0 < t_pad) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int64_t j = jb * kR + grp * 4 + s;
                        if (j < n_patches) {
                            double2 *out = reinterpret_cast<double2 *>(
                                g + ((b * n_classes + c) * n_patches + j) * ld + pad + t0);
#pragma unroll
                            for (int k = 0; k < kLaneT / 2; ++k)
                                out[k] = make_double2(acc[s][2 * k], acc[s][2 * k + 1]);
                        }
                    }
                }
            } else {
                // fused stage 2.  The lanes own 16 consecutive bins each (128 bytes apart): the
                // receiver's 512 bins are transposed through a swizzled 4 KB buffer so that
                // every store instruction of the warp covers 512 contiguous bytes (full
                // NVLink / L2 lines), then stored to all ranks and added to E_total.
                const double cf = fuse.coef[b];
                const uint32_t epi = sm_epi + (uint32_t)warp * 4096;
                const int64_t tq = q_t0(q);                      // first bin of the quarter
                // chunk m = i * 32 + lane of the quarter = bins 2m, 2m + 1; E_total is read for
                // a whole receiver (8 independent loads per lane) BEFORE the transpose, so that
                // one memory latency is exposed per receiver instead of one per chunk
                auto row_of = [&](int s) { return (b * n_alloc + jb * kR + grp * 4 + s) * ld + pad + tq; };
                auto load_total = [&](int s, double2 (&tot)[8]) {
                    const bool valid = jb * kR + grp * 4 + s < n_patches;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int m = i * 32 + lane;
                        tot[i] = (valid && tq + 2 * m < t_pad)
                                     ? *reinterpret_cast<const double2 *>(fuse.total + row_of(s) + 2 * m)
                                     : make_double2(0.0, 0.0);
                    }
                };
                double2 tot[8];
                load_total(0, tot);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int64_t j = jb * kR + grp * 4 + s;
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < kLaneT / 2; ++k) {
                        // E_k = fma(coef, G, 0): the rounding of k_mix
                        const double e0v = fma(cf, acc[s][2 * k], 0.0);
                        const double e1v = fma(cf, acc[s][2 * k + 1], 0.0);
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(
                                         epi + (uint32_t)(lane * 128 + ((k ^ (lane & 7)) << 4))),
                                     "d"(e0v), "d"(e1v) : "memory");
                    }
                    __syncwarp();
                    double2 e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int m = i * 32 + lane, ln = m >> 3, ch = m & 7;
                        lds_f64x2(e[i].x, e[i].y, epi + (uint32_t)(ln * 128 + ((ch ^ (ln & 7)) << 4)));
                    }
                    const int64_t row = row_of(s);
                    if (j < n_patches) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int m = i * 32 + lane;
                            if (tq + 2 * m >= t_pad) continue;
                            tot[i].x += e[i].x;
                            tot[i].y += e[i].y;
                            *reinterpret_cast<double2 *>(fuse.total + row + 2 * m) = tot[i];
                        }
                    }
                    if (s + 1 < 4) load_total(s + 1, tot);      // in flight during the stores
                    if (j < n_patches) {
                        for (int p = 0; p < fuse.n; ++p) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int m = i * 32 + lane;
                                if (tq + 2 * m < t_pad)
                                    *reinterpret_cast<double2 *>(fuse.cur[p] + row + 2 * m) = e[i];
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled encode_fn() {
    static EncodeTiled fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult qres;
        void *p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
                cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiled)p;
    }
    return fn;
}

template <int H, int B>
int launch(const double *e_prev, double *g, const int64_t *ent_ptr, const WinRecord *recs,
           const int32_t *cta_order, int64_t n_patches, int64_t n_alloc, int64_t n_classes,
           int64_t n_dirs, int64_t n_bands, int64_t b_lo, int64_t b_hi, int64_t j_lo,
           int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad, const FuseArgs &fuse,
           cudaStream_t st) {
    using C = Cfg<H>;
    const int64_t n_blocks = ceil_div(n_patches, kR);
    const int64_t jb_lo = j_lo / kR, jb_hi = ceil_div(j_hi, kR);
    const int64_t n_jb = jb_hi - jb_lo;
    const int64_t n_tiles = n_classes * n_jb;
    if (n_tiles == 0 || b_hi == b_lo) return 0;
    SPB_REQUIRE(n_tiles <= 2147483647LL, "too many tiles for one launch");
    SPB_REQUIRE(ld % 16 == 0 && pad % 16 == 0, "row pitch must be a multiple of 16 bins");
    SPB_REQUIRE(pad >= H + 16, "pad smaller than the delay window");
    // quarters per band: 512-bin units of one band a CTA covers (the rest of its four
    // lane quarters go to further bands)
    const int64_t units = ceil_div(t_pad, (int64_t)kQuarterT);
    const int qpb = units >= 4 ? 4 : (units >= 2 ? 2 : 1);
    const int n_tchunks = (int)ceil_div(units, (int64_t)qpb);
    const int bands_per_cta = 4 / qpb;
    const int64_t n_bgroups = ceil_div(b_hi - b_lo, (int64_t)bands_per_cta);
    SPB_REQUIRE(n_bgroups * n_tchunks <= 65535, "too many band groups x time chunks");

    EncodeTiled encode = encode_fn();
    if (!encode) return fail(-2, "cuTensorMapEncodeTiled", "driver entry point not found");
    const int64_t total = n_bands * n_alloc * n_dirs * ld;             // doubles in e_prev
    SPB_REQUIRE(total / 16 <= 2147483647LL, "histogram too large for 32-bit tensor rows");
    TmapSet tmaps;
    for (int j = 0; j < 8; ++j) {
        cuuint64_t dims[2] = {16, (cuuint64_t)((total - 2 * j) / 16)};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {16, (cuuint32_t)(32 * qpb + 1)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmaps.m[j], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
                            (void *)(e_prev + 2 * j), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(-2, "cuTensorMapEncodeTiled", "encode failed");
    }

    constexpr int S = 8 / B, TS = 512 / (B * C::kCols);
    const size_t smem = (size_t)S * B * kBoxArea + (size_t)(S + TS) * B * sizeof(WinRecord) +
                        kEpiBytes + (2 * S + 8 * TS) * sizeof(uint64_t) + 16;
    SPB_CUDA(cudaFuncSetAttribute(k_gather_tmem<H, B>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char *dbg_env = getenv("SPB_TMEM_DBG");   // timing experiments: 1 = consumers idle,
    const int dbg = dbg_env ? atoi(dbg_env) : 0;    // 2 = fill idle (results are wrong)
    dim3 grid((unsigned)n_tiles, (unsigned)(n_bgroups * n_tchunks));
    k_gather_tmem<H, B><<<grid, kThreads, smem, st>>>(
        tmaps, fuse, g, ent_ptr, recs, cta_order, n_patches, n_alloc, n_blocks, n_dirs, b_lo, b_hi,
        jb_lo, n_jb, n_classes, t_pad, ld, pad, qpb, n_tchunks, dbg);
    return check_launch("k_gather_tmem");
}

}  // namespace tmg
}  // namespace spb

using namespace spb;

extern "C" {

int spb_tmem_batch(void) { return tmg::kBatch; }

int spb_window_geometry(int dtype, int64_t *receivers_per_tile, int64_t *max_window,
                        int64_t *record_bytes) {
    SPB_REQUIRE(dtype == SPB_F64, "the tensor-memory gather is FP64 only");
    *receivers_per_tile = tmg::kR;
    *max_window = 10;
    *record_bytes = sizeof(tmg::WinRecord);
    return 0;
}

static int gather_tmem_dispatch(const void *e_prev, void *g, const int64_t *ent_ptr,
                                const void *recs, const int32_t *cta_order, int64_t n_patches,
                                int64_t n_alloc, int64_t n_classes, int64_t n_dirs,
                                int64_t n_bands, int64_t b_lo, int64_t b_hi, int64_t j_lo,
                                int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad,
                                int64_t window, int dtype, const tmg::FuseArgs &fuse,
                                void *stream) {
    SPB_REQUIRE(e_prev && ent_ptr && (g || fuse.n > 0), "null pointer");
    SPB_REQUIRE(dtype == SPB_F64, "the tensor-memory gather is FP64 only");
    SPB_REQUIRE(0 <= j_lo && j_lo <= j_hi && j_hi <= n_patches, "receiver range");
    SPB_REQUIRE(0 <= b_lo && b_lo <= b_hi && b_hi <= n_bands, "band range");
    SPB_REQUIRE(n_alloc >= n_patches, "n_alloc < n_patches");
    SPB_REQUIRE(j_lo == j_hi || j_lo % tmg::kR == 0,
                "j_lo must be a multiple of the receiver tile (8)");
    SPB_REQUIRE(t_pad % 256 == 0 && ld == pad + t_pad, "layout (use spb_exchange_layout)");
    SPB_REQUIRE(((uintptr_t)e_prev & 15) == 0, "e_prev must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const double *ep = (const double *)e_prev;
    const tmg::WinRecord *r = (const tmg::WinRecord *)recs;
    if (window == 4)
        return tmg::launch<4, tmg::kBatch>(ep, (double *)g, ent_ptr, r, cta_order, n_patches, n_alloc,
                                           n_classes, n_dirs, n_bands, b_lo, b_hi, j_lo, j_hi, t_pad,
                                           ld, pad, fuse, st);
    if (window == 10)
        return tmg::launch<10, tmg::kBatch>(ep, (double *)g, ent_ptr, r, cta_order, n_patches, n_alloc,
                                            n_classes, n_dirs, n_bands, b_lo, b_hi, j_lo, j_hi, t_pad,
                                            ld, pad, fuse, st);
    return fail(-1, "invalid argument", "window must be 4 or 10");
}

int spb_exchange_gather_tmem(const void *e_prev, void *g, const int64_t *ent_ptr,
                             const void *recs, const int32_t *cta_order, int64_t n_patches,
                             int64_t n_alloc, int64_t n_classes, int64_t n_dirs,
                             int64_t n_bands, int64_t b_lo, int64_t b_hi, int64_t j_lo,
                             int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad,
                             int64_t window, int dtype, void *stream) {
    tmg::FuseArgs fuse = {};
    return gather_tmem_dispatch(e_prev, g, ent_ptr, recs, cta_order, n_patches, n_alloc,
                                n_classes, n_dirs, n_bands, b_lo, b_hi, j_lo, j_hi, t_pad, ld,
                                pad, window, dtype, fuse, stream);
}

int spb_exchange_order_fused(const void *e_prev, const uint64_t *cur_ptrs_h, int n_peers,
                             void *e_total, const void *coef, const int64_t *ent_ptr,
                             const void *recs, const int32_t *cta_order, int64_t n_patches,
                             int64_t n_alloc, int64_t n_bands, int64_t b_lo, int64_t b_hi,
                             int64_t j_lo, int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad,
                             int64_t window, int dtype, void *stream) {
    SPB_REQUIRE(cur_ptrs_h && e_total && coef, "null pointer");
    SPB_REQUIRE(n_peers >= 1 && n_peers <= tmg::kMaxPeers, "1..8 destination buffers");
    tmg::FuseArgs fuse = {};
    for (int p = 0; p < n_peers; ++p) {
        SPB_REQUIRE(cur_ptrs_h[p] && (cur_ptrs_h[p] & 15) == 0, "destination buffer alignment");
        fuse.cur[p] = (double *)(uintptr_t)cur_ptrs_h[p];
    }
    fuse.total = (double *)e_total;
    fuse.coef = (const double *)coef;
    fuse.n = n_peers;
    return gather_tmem_dispatch(e_prev, nullptr, ent_ptr, recs, cta_order, n_patches, n_alloc, 1,
                                1, n_bands, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, window, dtype,
                                fuse, stream);
}

}  // extern "C"
