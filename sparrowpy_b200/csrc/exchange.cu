// Energy exchange, initial energy and receiver collection kernels (sm_100a).
//
// Replaces the numba kernels `_energy_exchange_init_energy`, `_energy_exchange`
// and `_collect_receiver_energy` (reference RadiosityFast.py:1037-1185).  See
// include/sparrow_b200.h for the data layout and DESIGN.md for the roofline.
#include "common.cuh"

namespace spb {

std::string &last_error() {
    static thread_local std::string msg;
    return msg;
}

constexpr int kWarpsPerCta = 8;
constexpr int kChunks = 8;                 // 32-bin chunks per lane
constexpr int kTimeTile = 32 * kChunks;    // 256 time bins per warp

// ---------------------------------------------------------------------------
// initial energy: RadiosityFast.py:1037-1070
// ---------------------------------------------------------------------------
// e0 is (patch, direction, band); histogram rows are band-major:
// row(b, patch, dir) = (b * n_alloc + patch) * D + dir
template <typename T>
__global__ void k_init_scatter(T *__restrict__ e_total, T *__restrict__ e_prev,
                               const T *__restrict__ e0,
                               const int32_t *__restrict__ delay0, int64_t n_patches,
                               int64_t n_alloc, int64_t n_dirs, int64_t n_bands,
                               int64_t band_lo, int64_t n_samples, int64_t ld, int64_t pad) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_patches * n_dirs * n_bands) return;
    const int64_t b = idx % n_bands;
    const int64_t pd = idx / n_bands;               // patch * D + dir
    const int64_t patch = pd / n_dirs;
    const int32_t d = delay0[patch];
    if (d < 0 || d >= n_samples) return;   // out-of-range energy is dropped
    const T v = e0[idx];
    const int64_t row = (band_lo + b) * n_alloc * n_dirs + pd;
    e_total[row * ld + pad + d] += v;
    if (e_prev) e_prev[row * ld + pad + d] += v;
}

// ---------------------------------------------------------------------------
// stage 1: sparse delay-and-sum  (RadiosityFast.py:1124-1143, one order)
//   G[seg, b, t] = sum_q wgt_q * E_prev[src_q * B + b, t - dly_q]
// one warp per (band, segment); lanes run along time, kChunks chunks of 32 bins
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_gather(const T *__restrict__ e_prev, T *__restrict__ g,
         const int64_t *__restrict__ seg_ptr, const int32_t *__restrict__ src,
         const T *__restrict__ wgt, const int32_t *__restrict__ dly,
         int64_t n_patches, int64_t n_alloc, int64_t n_classes, int64_t n_dirs,
         int64_t b_lo, int64_t n_b, int64_t j_lo, int64_t n_j, int64_t ld, int64_t pad) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t n_local = n_classes * n_j;
    const int64_t widx = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    if (widx >= n_local * n_b) return;
    const int64_t b = b_lo + widx / n_local;
    const int64_t loc = widx % n_local;
    const int64_t c = loc / n_j;
    const int64_t seg = c * n_patches + j_lo + (loc - c * n_j);
    const int64_t p0 = seg_ptr[seg], p1 = seg_ptr[seg + 1];
    if (p0 == p1) return;                  // empty segment: row is never read
    const int64_t t0 = (int64_t)blockIdx.y * kTimeTile;

    T acc[kChunks];
#pragma unroll
    for (int v = 0; v < kChunks; ++v) acc[v] = T(0);
    const T *base = e_prev + b * n_alloc * n_dirs * ld + pad + t0 + lane;

    for (int64_t p = p0; p < p1; p += 32) {
        const int64_t q = p + lane;
        int32_t s = 0, d = 0;
        T w = T(0);
        if (q < p1) { s = src[q]; w = wgt[q]; d = dly[q]; }
        const int cnt = (int)min((int64_t)32, p1 - p);
        for (int k = 0; k < cnt; ++k) {
            const int32_t sk = __shfl_sync(0xffffffffu, s, k);
            const int32_t dk = __shfl_sync(0xffffffffu, d, k);
            const T wk = __shfl_sync(0xffffffffu, w, k);
            const T *row = base + (int64_t)sk * ld - dk;
#pragma unroll
            for (int v = 0; v < kChunks; ++v) acc[v] = fma(wk, row[32 * v], acc[v]);
        }
    }
    T *out = g + (b * n_classes * n_patches + seg) * ld + pad + t0 + lane;
#pragma unroll
    for (int v = 0; v < kChunks; ++v) out[32 * v] = acc[v];
}

// ---------------------------------------------------------------------------
// stage 2: BRDF contraction + accumulation
//   E_cur[j,d,b,t] = sum_c coef[c,d,b] * G[c,j,b,t];  E_total += E_cur
// one CTA per (patch, band, time slab); threads along time
// ---------------------------------------------------------------------------
constexpr int kMixDirs = 8;
constexpr int kMaxPeers = 8;

// Where stage 2 stores E_k: the local buffer (one GPU), every peer's copy of the
// buffer through NVLink P2P stores, or one NVSwitch multicast store (multimem.st)
// that the switch replicates into all ranks' copies.  With the peer modes the
// per-order all-gather of the sharded exchange disappears into this kernel.
enum : int { kStoreLocal = 0, kStorePeers = 1, kStoreMulticast = 2 };

template <typename T>
struct StoreTargets {
    T *ptr[kMaxPeers];      // kStoreLocal: ptr[0]; kStorePeers: all ranks' buffers
    T *mc;                  // kStoreMulticast: multicast address of the buffer
    int n;
};

// 16-byte vectors: every stage-2 thread owns 16/sizeof(T) consecutive time bins, so
// local, peer and multicast stores are all 128-bit
template <typename T>
struct alignas(16) Vec16 {
    static constexpr int kN = 16 / sizeof(T);
    T v[kN];
};

template <typename T>
__device__ __forceinline__ void multimem_store16(T *addr, const Vec16<T> &x) {
    const float *f = reinterpret_cast<const float *>(&x);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr),
                 "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3])
                 : "memory");
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_mix(const T *__restrict__ g, StoreTargets<T> cur, T *__restrict__ e_total,
      const int64_t *__restrict__ seg_ptr, const T *__restrict__ coef,
      int64_t n_patches, int64_t n_alloc, int64_t n_classes, int64_t n_dirs,
      int64_t n_bands, int64_t b_lo, int64_t j_lo, int64_t n_j, int64_t t_pad, int64_t ld,
      int64_t pad) {
    constexpr int kV = Vec16<T>::kN;
    const int64_t jb = blockIdx.x;
    const int64_t b = b_lo + jb / n_j;
    const int64_t j = j_lo + jb % n_j;
    const int64_t t = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * kV;
    if (t >= t_pad) return;
    for (int64_t d0 = 0; d0 < n_dirs; d0 += kMixDirs) {
        Vec16<T> acc[kMixDirs];
#pragma unroll
        for (int dd = 0; dd < kMixDirs; ++dd)
#pragma unroll
            for (int q = 0; q < kV; ++q) acc[dd].v[q] = T(0);
        for (int64_t c = 0; c < n_classes; ++c) {
            const int64_t seg = c * n_patches + j;
            if (seg_ptr[seg] == seg_ptr[seg + 1]) continue;   // CTA-uniform
            const Vec16<T> gv = *reinterpret_cast<const Vec16<T> *>(
                g + (b * n_classes * n_patches + seg) * ld + pad + t);
            const T *cf = coef + (c * n_dirs + d0) * n_bands + b;
#pragma unroll
            for (int dd = 0; dd < kMixDirs; ++dd)
                if (d0 + dd < n_dirs) {
                    const T w = cf[dd * n_bands];
#pragma unroll
                    for (int q = 0; q < kV; ++q) acc[dd].v[q] = fma(w, gv.v[q], acc[dd].v[q]);
                }
        }
#pragma unroll
        for (int dd = 0; dd < kMixDirs; ++dd) {
            if (d0 + dd < n_dirs) {
                const int64_t o = ((b * n_alloc + j) * n_dirs + d0 + dd) * ld + pad + t;
                if (MODE == kStoreLocal) {
                    *reinterpret_cast<Vec16<T> *>(cur.ptr[0] + o) = acc[dd];
                } else if (MODE == kStorePeers) {
                    for (int p = 0; p < cur.n; ++p)
                        *reinterpret_cast<Vec16<T> *>(cur.ptr[p] + o) = acc[dd];
                } else {
                    multimem_store16(cur.mc + o, acc[dd]);
                }
                Vec16<T> tot = *reinterpret_cast<Vec16<T> *>(e_total + o);
#pragma unroll
                for (int q = 0; q < kV; ++q) tot.v[q] += acc[dd].v[q];
                *reinterpret_cast<Vec16<T> *>(e_total + o) = tot;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// receiver collection: RadiosityFast.py:735-748, :1148-1185
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_collect_partial(const T *__restrict__ e_total, const int32_t *__restrict__ rdir,
                  const int32_t *__restrict__ shift, const T *__restrict__ scale,
                  int64_t n_patches, int64_t n_alloc, int64_t n_dirs, int64_t n_bands,
                  int64_t n_samples, int64_t ld, int64_t pad, T *__restrict__ partial,
                  int64_t n_split) {
    // blockIdx.y = band * R + receiver: CTAs that are launched together read the same rows
    // of the same band for different receivers, so the histogram is streamed from HBM once
    // per (band, split) and the other receivers hit in L2
    const int64_t n_rcv = gridDim.y / n_bands;
    const int64_t b = blockIdx.y / n_rcv, r = blockIdx.y - b * n_rcv;
    const int64_t rb = r * n_bands + b;        // output row: receiver * B + band
    const int64_t split = blockIdx.z;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // output bin
    const int64_t k_lo = n_patches * split / n_split;
    const int64_t k_hi = n_patches * (split + 1) / n_split;
    T acc = T(0);
    if (t < n_samples) {
        for (int64_t k = k_lo; k < k_hi; ++k) {
            const int64_t rk = r * n_patches + k;
            const T sc = scale[rk * n_bands + b];
            if (sc == T(0)) continue;          // invisible patch (CTA-uniform)
            int64_t ts = t - shift[rk];
            if (ts < 0) ts += n_samples;       // circular np.roll
            acc += e_total[((b * n_alloc + k) * n_dirs + rdir[rk]) * ld + pad + ts] * sc;
        }
        partial[((split * gridDim.y) + rb) * n_samples + t] = acc;
    }
}

// ---------------------------------------------------------------------------
// receiver collection for MANY receivers of a diffuse scene (D = 1): the histogram row of a
// (band, patch) is staged ONCE in shared memory (cp.async ring, the receivers' scale and shift
// riding along) and applied to kRcv receivers from there -- k_collect_partial re-reads the
// row from L2/HBM for every receiver.  256 threads; a thread owns kBins output bins
// (t = chunk*256*kBins + tid + 256 q) of kRcv receivers: kRcv*kBins accumulators and one
// conflict-free shared-memory operand per FMA (the 128 B/clk/SM shared-memory path is the
// roof: a quarter of the FP64 pipe).  The row is stored TWICE back to back, so the circular
// read E[(t - shift) mod T] is the plain read row2[t - shift + T]: per receiver one address,
// per bin an immediate offset.  Same partial-sum layout as k_collect_partial.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async_16(uint32_t smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(gmem) : "memory");
}
template <int kBytes>
__device__ __forceinline__ void cp_async_small(uint32_t smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem), "l"(gmem),
                 "n"(kBytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_pending(int n) {   // n = groups left in flight
    if (n <= 0) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    else if (n == 1) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    else asm volatile("cp.async.wait_group 2;\n" ::: "memory");
}

template <typename T, int kRcv, int kBins>
__global__ void __launch_bounds__(256)
k_collect_staged(const T *__restrict__ e_total, const int32_t *__restrict__ shift,
                 const T *__restrict__ scale, int n_receivers, int64_t n_patches,
                 int64_t n_alloc, int n_bands, int n_samples, int64_t ld, int64_t pad,
                 T *__restrict__ partial, int n_split, int n_rgroups, int n_chunks,
                 int n_stages) {
    // blockIdx.x = ((band * n_split + split) * n_chunks + chunk) * n_rgroups + rgroup: the CTAs
    // that read the same histogram rows are neighbours in launch order and meet in L2
    int64_t id = blockIdx.x;
    const int rg = (int)(id % n_rgroups); id /= n_rgroups;
    const int chunk = (int)(id % n_chunks); id /= n_chunks;
    const int split = (int)(id % n_split);
    const int b = (int)(id / n_split);
    const int r0 = rg * kRcv;
    const int tid = threadIdx.x;
    const int t0 = chunk * 256 * kBins + tid;
    const int64_t k_lo = n_patches * split / n_split;
    const int n_k = (int)(n_patches * (split + 1) / n_split - k_lo);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kVec = 16 / (int)sizeof(T);
    // one stage: [t_al - T, t_al) copy A | [t_al, t_al + T) copy B | ... up to t_al + row_elems
    // (read by tail threads only) | kRcv scales | kRcv byte offsets.  A multiple of 16 bytes.
    const int t_al = (n_samples + kVec - 1) / kVec * kVec;
    const int n_vec = t_al / kVec;                        // 16-byte pieces of a row
    const int row_elems = n_chunks * 256 * kBins;         // >= n_samples
    const int row2 = t_al + row_elems;
    const int stage_bytes = (row2 + kRcv) * (int)sizeof(T) + kRcv * 4;
    const bool aligned = t_al == n_samples;               // copy A starts on a 16-byte boundary
    const uint32_t smem0 = smem_addr(smem_raw);
    // bins past copy B are read by tail threads only (t >= n_samples, never stored); zero them
    // once so that no stale bit pattern is ever multiplied (the pieces of copy B may spill
    // kVec - 1 elements into this zone: finite values of the row's padding).  Scale and shift
    // slots of receivers past the last one stay zero for the whole kernel.
    for (int s = 0; s < n_stages; ++s) {
        T *row = reinterpret_cast<T *>(smem_raw + (size_t)s * stage_bytes);
        for (int i = 2 * t_al + tid; i < row2; i += 256) row[i] = T(0);
        if (tid < kRcv) {
            row[row2 + tid] = T(0);
            reinterpret_cast<int32_t *>(row + row2 + kRcv)[tid] = 0;
        }
    }
    __syncthreads();
    // the receiver factors of this thread's receiver (threads < kRcv): everything in the loop
    // is asynchronous -- a synchronous global load here would stall the whole CTA at the next
    // barrier for a DRAM latency per patch
    const bool has_rcv = tid < kRcv && r0 + tid < n_receivers;
    const T *my_scale = scale + ((int64_t)(r0 + (has_rcv ? tid : 0)) * n_patches + k_lo) * n_bands + b;
    const int32_t *my_shift = shift + (int64_t)(r0 + (has_rcv ? tid : 0)) * n_patches + k_lo;
    const T *band_rows = e_total + ((int64_t)b * n_alloc + k_lo) * ld + pad;

    auto issue = [&](int i, int stage) {       // patch k_lo + i -> stage
        const uint32_t so = smem0 + stage * stage_bytes;
        const T *src = band_rows + (int64_t)i * ld;
        for (int v = tid; v < n_vec; v += 256) {
            cp_async_16(so + (t_al + v * kVec) * (int)sizeof(T), src + (size_t)v * kVec);
            if (aligned) cp_async_16(so + v * 16, src + (size_t)v * kVec);
        }
        if (!aligned)                           // copy A element by element
            for (int v = tid; v < n_samples; v += 256)
                cp_async_small<(int)sizeof(T)>(so + (t_al - n_samples + v) * (int)sizeof(T),
                                               src + v);
        if (has_rcv) {
            cp_async_small<(int)sizeof(T)>(so + (row2 + tid) * (int)sizeof(T),
                                           my_scale + (int64_t)i * n_bands);
            cp_async_small<4>(so + (row2 + kRcv) * (int)sizeof(T) + tid * 4, my_shift + i);
        }
    };

    T acc[kRcv][kBins];
#pragma unroll
    for (int r = 0; r < kRcv; ++r)
#pragma unroll
        for (int q = 0; q < kBins; ++q) acc[r][q] = T(0);

    for (int i = 0; i < n_stages - 1; ++i) {
        if (i < n_k) issue(i, i);
        cp_async_commit();
    }
    int stage = 0;                             // stage of patch i
    int refill = n_stages - 1;                 // stage of patch i + n_stages - 1 (= of patch i-1)
    const int sh_scale = (int)sizeof(T);
    for (int i = 0; i < n_k; ++i) {
        cp_async_wait_pending(n_stages - 2);   // the copies of patch i have landed
        __syncthreads();                       // ... for every thread; patch i-1 is consumed
        if (i + n_stages - 1 < n_k) issue(i + n_stages - 1, refill);
        cp_async_commit();
        const unsigned char *st = smem_raw + (size_t)stage * stage_bytes;
        const unsigned char *mine = st + (size_t)(t0 + t_al) * sizeof(T);
        const T *sc = reinterpret_cast<const T *>(st) + row2;
        const int4 *sh4 = reinterpret_cast<const int4 *>(sc + kRcv);
        T c[kRcv];
        int sh[kRcv];
#pragma unroll
        for (int r = 0; r < kRcv; r += kVec) {
            const int4 raw = *reinterpret_cast<const int4 *>(sc + r);   // kVec scales
            const T *cv = reinterpret_cast<const T *>(&raw);
#pragma unroll
            for (int j = 0; j < kVec; ++j) c[r + j] = cv[j];
        }
#pragma unroll
        for (int r = 0; r < kRcv; r += 4) {
            const int4 o = sh4[r / 4];
            sh[r] = o.x; sh[r + 1] = o.y; sh[r + 2] = o.z; sh[r + 3] = o.w;
        }
#pragma unroll
        for (int r = 0; r < kRcv; ++r) {
            // row2[t_al - shift + t]: copy B for t >= shift, copy A (the wrapped bins) below
            const T *win = reinterpret_cast<const T *>(mine - sh[r] * sh_scale);
#pragma unroll
            for (int q = 0; q < kBins; ++q) acc[r][q] += win[256 * q] * c[r];
        }
        stage = stage + 1 == n_stages ? 0 : stage + 1;
        refill = refill + 1 == n_stages ? 0 : refill + 1;
    }
    cp_async_wait_pending(0);
#pragma unroll
    for (int r = 0; r < kRcv; ++r) {
        if (r0 + r >= n_receivers) continue;
        const int64_t rb = (int64_t)(r0 + r) * n_bands + b;
        T *dst = partial + ((int64_t)split * n_receivers * n_bands + rb) * n_samples;
#pragma unroll
        for (int q = 0; q < kBins; ++q) {
            const int t = t0 + 256 * q;
            if (t < n_samples) dst[t] = acc[r][q];
        }
    }
}

template <typename T>
__global__ void k_collect_reduce(const T *__restrict__ partial, T *__restrict__ mono,
                                 int64_t n_out, int64_t n_split) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    T acc = T(0);
    for (int64_t s = 0; s < n_split; ++s) acc += partial[s * n_out + o];
    mono[o] = acc;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_collect_patchwise(const T *__restrict__ e_total, const int32_t *__restrict__ rdir,
                    const int32_t *__restrict__ shift, const T *__restrict__ scale,
                    int64_t n_patches, int64_t n_alloc, int64_t n_dirs, int64_t n_bands,
                    int64_t n_samples, int64_t ld, int64_t pad, T *__restrict__ out) {
    const int64_t rkb = blockIdx.x;            // (receiver * N + patch) * B + band
    const int64_t b = rkb % n_bands;
    const int64_t rk = rkb / n_bands;
    const int64_t k = rk % n_patches;
    const int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (t >= n_samples) return;
    int64_t ts = t - shift[rk];
    if (ts < 0) ts += n_samples;
    out[rkb * n_samples + t] =
        e_total[((b * n_alloc + k) * n_dirs + rdir[rk]) * ld + pad + ts] *
        scale[rk * n_bands + b];
}

// ---------------------------------------------------------------------------
// typed launchers
// ---------------------------------------------------------------------------
template <typename T>
int scatter_t(void *e_total, void *e_prev, const void *e0, const int32_t *delay0,
              int64_t n_patches, int64_t n_alloc, int64_t n_dirs, int64_t n_bands,
              int64_t band_lo, int64_t n_samples, int64_t ld, int64_t pad, cudaStream_t st) {
    const int64_t n_in = n_patches * n_dirs * n_bands;
    if (n_in == 0) return 0;
    k_init_scatter<T><<<(unsigned)ceil_div(n_in, 256), 256, 0, st>>>(
        (T *)e_total, (T *)e_prev, (const T *)e0, delay0, n_patches, n_alloc, n_dirs, n_bands,
        band_lo, n_samples, ld, pad);
    return check_launch("k_init_scatter");
}

template <typename T>
int init_t(void *e_total, void *e_prev, const void *e0, const int32_t *delay0,
           int64_t n_patches, int64_t n_alloc, int64_t n_dirs, int64_t n_bands,
           int64_t n_samples, int64_t ld, int64_t pad, cudaStream_t st) {
    const int64_t n_rows = n_alloc * n_dirs * n_bands;
    SPB_CUDA(cudaMemsetAsync(e_total, 0, sizeof(T) * n_rows * ld, st));
    if (e_prev) SPB_CUDA(cudaMemsetAsync(e_prev, 0, sizeof(T) * n_rows * ld, st));
    return scatter_t<T>(e_total, e_prev, e0, delay0, n_patches, n_alloc, n_dirs, n_bands, 0,
                        n_samples, ld, pad, st);
}

template <typename T>
int gather_t(const void *e_prev, void *g, const int64_t *seg_ptr, const int32_t *src,
             const void *wgt, const int32_t *dly, int64_t n_patches, int64_t n_alloc,
             int64_t n_classes, int64_t n_dirs, int64_t b_lo, int64_t b_hi, int64_t j_lo,
             int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad, cudaStream_t st) {
    const int64_t n_j = j_hi - j_lo, n_b = b_hi - b_lo;
    const int64_t n_work = n_classes * n_j * n_b;
    if (n_work == 0) return 0;
    dim3 grid((unsigned)ceil_div(n_work, kWarpsPerCta), (unsigned)(t_pad / kTimeTile));
    k_gather<T><<<grid, kWarpsPerCta * 32, 0, st>>>(
        (const T *)e_prev, (T *)g, seg_ptr, src, (const T *)wgt, dly, n_patches, n_alloc,
        n_classes, n_dirs, b_lo, n_b, j_lo, n_j, ld, pad);
    return check_launch("k_gather");
}

template <typename T>
int mix_t(const void *g, const StoreTargets<T> &cur, int mode, void *e_total,
          const int64_t *seg_ptr, const void *coef, int64_t n_patches, int64_t n_alloc,
          int64_t n_classes, int64_t n_dirs, int64_t n_bands, int64_t b_lo, int64_t b_hi,
          int64_t j_lo, int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad,
          cudaStream_t st) {
    const int64_t n_j = j_hi - j_lo, n_b = b_hi - b_lo;
    if (n_j * n_b == 0) return 0;
    SPB_REQUIRE(n_j * n_b <= 2147483647LL, "too many (patch, band) rows");
    dim3 grid((unsigned)(n_j * n_b), (unsigned)ceil_div(t_pad, 256 * Vec16<T>::kN));
#define SPB_MIX_LAUNCH(MODE)                                                                \
    k_mix<T, MODE><<<grid, 256, 0, st>>>((const T *)g, cur, (T *)e_total, seg_ptr,          \
                                         (const T *)coef, n_patches, n_alloc, n_classes,    \
                                         n_dirs, n_bands, b_lo, j_lo, n_j, t_pad, ld, pad)
    if (mode == kStoreLocal) SPB_MIX_LAUNCH(kStoreLocal);
    else if (mode == kStorePeers) SPB_MIX_LAUNCH(kStorePeers);
    else SPB_MIX_LAUNCH(kStoreMulticast);
#undef SPB_MIX_LAUNCH
    return check_launch("k_mix");
}

template <typename T>
int mix_dispatch(const void *g, void *e_cur, const uint64_t *peer_ptrs_h, int n_peers,
                 void *cur_mc, int mode, void *e_total, const int64_t *seg_ptr,
                 const void *coef, int64_t n_patches, int64_t n_alloc, int64_t n_classes,
                 int64_t n_dirs, int64_t n_bands, int64_t b_lo, int64_t b_hi, int64_t j_lo,
                 int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad, cudaStream_t st) {
    StoreTargets<T> cur;
    for (int p = 0; p < kMaxPeers; ++p) cur.ptr[p] = nullptr;
    cur.mc = (T *)cur_mc;
    cur.n = 1;
    cur.ptr[0] = (T *)e_cur;
    if (mode == kStorePeers) {
        cur.n = n_peers;
        for (int p = 0; p < n_peers; ++p) cur.ptr[p] = (T *)(uintptr_t)peer_ptrs_h[p];
    }
    return mix_t<T>(g, cur, mode, e_total, seg_ptr, coef, n_patches, n_alloc, n_classes,
                    n_dirs, n_bands, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
}

}  // namespace spb

using namespace spb;

template <typename T, int kRcv, int kBins>
static int collect_staged_t(const void *e_total, const int32_t *shift, const void *scale,
                            int64_t n_receivers, int64_t n_patches, int64_t n_alloc,
                            int64_t n_bands, int64_t n_samples, int64_t ld, int64_t pad,
                            void *mono, void *partial, int64_t n_split, int n_stages,
                            cudaStream_t st) {
    const int64_t n_chunks = ceil_div(n_samples, 256 * kBins);
    const int64_t n_rgroups = ceil_div(n_receivers, kRcv);
    constexpr int64_t kVec = 16 / (int64_t)sizeof(T);
    const int64_t row2 = round_up(n_samples, kVec) + n_chunks * 256 * kBins;   // doubled row
    const int64_t stage_bytes = (row2 + kRcv) * (int64_t)sizeof(T) + kRcv * 4;
    if (n_stages == 0) n_stages = 3 * stage_bytes <= 100 * 1024 ? 3 : 2;
    SPB_REQUIRE(n_stages >= 2 && n_stages <= 4, "n_stages must be 2..4");
    const int64_t smem = n_stages * stage_bytes;
    SPB_REQUIRE(smem <= 220 * 1024, "histogram rows too long for shared-memory staging");
    const int64_t n_cta = n_bands * n_split * n_chunks * n_rgroups;
    SPB_REQUIRE(n_cta <= 2147483647LL, "grid too large");
    auto kern = k_collect_staged<T, kRcv, kBins>;
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)n_cta, 256, (size_t)smem, st>>>(
        (const T *)e_total, shift, (const T *)scale, (int)n_receivers, n_patches, n_alloc,
        (int)n_bands, (int)n_samples, ld, pad, (T *)partial, (int)n_split, (int)n_rgroups,
        (int)n_chunks, n_stages);
    int rc = check_launch("k_collect_staged");
    if (rc) return rc;
    const int64_t n_out = n_receivers * n_bands * n_samples;
    k_collect_reduce<T><<<(unsigned)ceil_div(n_out, 256), 256, 0, st>>>(
        (const T *)partial, (T *)mono, n_out, n_split);
    return check_launch("k_collect_reduce");
}

extern "C" {

int spb_version(void) { return 100; }

const char *spb_last_error(void) { return last_error().c_str(); }

int spb_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    SPB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SPB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

int spb_exchange_layout(int64_t n_samples, int64_t max_delay, int dtype, int64_t *t_pad,
                        int64_t *pad) {
    SPB_REQUIRE(n_samples > 0 && max_delay >= 0, "n_samples > 0, max_delay >= 0");
    SPB_REQUIRE(dtype == SPB_F64 || dtype == SPB_F32, "dtype");
    *t_pad = round_up(n_samples, kTimeTile);
    // one spare 32-bin bucket in front: the tiled gather stages windows that start
    // 32 bins before the delay bucket of a record (exchange_tma.cu)
    *pad = round_up(max_delay + 1, 32) + 32;
    return 0;
}

int spb_exchange_init(void *e_total, void *e_prev, const void *e0, const int32_t *delay0,
                      int64_t n_patches, int64_t n_alloc, int64_t n_dirs, int64_t n_bands,
                      int64_t n_samples, int64_t ld, int64_t pad, int dtype, void *stream) {
    SPB_REQUIRE(e_total && e0 && delay0, "null pointer");
    SPB_REQUIRE(ld >= pad + n_samples, "ld < pad + n_samples");
    SPB_REQUIRE(n_alloc >= n_patches, "n_alloc < n_patches");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SPB_F64)
        return init_t<double>(e_total, e_prev, e0, delay0, n_patches, n_alloc, n_dirs, n_bands,
                              n_samples, ld, pad, st);
    if (dtype == SPB_F32)
        return init_t<float>(e_total, e_prev, e0, delay0, n_patches, n_alloc, n_dirs, n_bands,
                             n_samples, ld, pad, st);
    return fail(-1, "invalid argument", "dtype");
}

int spb_exchange_scatter(void *e_total, void *e_prev, const void *e0, const int32_t *delay0,
                         int64_t n_patches, int64_t n_alloc, int64_t n_dirs,
                         int64_t n_bands_src, int64_t band_lo, int64_t n_bands_total,
                         int64_t n_samples, int64_t ld, int64_t pad, int dtype, void *stream) {
    SPB_REQUIRE(e_total && e0 && delay0, "null pointer");
    SPB_REQUIRE(ld >= pad + n_samples, "ld < pad + n_samples");
    SPB_REQUIRE(n_alloc >= n_patches, "n_alloc < n_patches");
    SPB_REQUIRE(band_lo >= 0 && band_lo + n_bands_src <= n_bands_total, "band window");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SPB_F64)
        return scatter_t<double>(e_total, e_prev, e0, delay0, n_patches, n_alloc, n_dirs,
                                 n_bands_src, band_lo, n_samples, ld, pad, st);
    if (dtype == SPB_F32)
        return scatter_t<float>(e_total, e_prev, e0, delay0, n_patches, n_alloc, n_dirs,
                                n_bands_src, band_lo, n_samples, ld, pad, st);
    return fail(-1, "invalid argument", "dtype");
}

#define SPB_CHECK_RANGES()                                                                  \
    SPB_REQUIRE(0 <= j_lo && j_lo <= j_hi && j_hi <= n_patches, "receiver range");         \
    SPB_REQUIRE(0 <= b_lo && b_lo <= b_hi && b_hi <= n_bands, "band range");               \
    SPB_REQUIRE(n_alloc >= n_patches, "n_alloc < n_patches");                              \
    SPB_REQUIRE(t_pad % kTimeTile == 0 && ld == pad + t_pad, "layout (use spb_exchange_layout)")

int spb_exchange_gather(const void *e_prev, void *g, const int64_t *seg_ptr,
                        const int32_t *src, const void *wgt, const int32_t *dly,
                        int64_t n_patches, int64_t n_alloc, int64_t n_classes, int64_t n_dirs,
                        int64_t n_bands, int64_t b_lo, int64_t b_hi, int64_t j_lo,
                        int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad, int dtype,
                        void *stream) {
    SPB_REQUIRE(e_prev && g && seg_ptr, "null pointer");
    SPB_CHECK_RANGES();
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SPB_F64)
        return gather_t<double>(e_prev, g, seg_ptr, src, wgt, dly, n_patches, n_alloc,
                                n_classes, n_dirs, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
    if (dtype == SPB_F32)
        return gather_t<float>(e_prev, g, seg_ptr, src, wgt, dly, n_patches, n_alloc,
                               n_classes, n_dirs, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
    return fail(-1, "invalid argument", "dtype");
}

int spb_exchange_mix(const void *g, void *e_cur, void *e_total, const int64_t *seg_ptr,
                     const void *coef, int64_t n_patches, int64_t n_alloc, int64_t n_classes,
                     int64_t n_dirs, int64_t n_bands, int64_t b_lo, int64_t b_hi,
                     int64_t j_lo, int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad,
                     int dtype, void *stream) {
    SPB_REQUIRE(g && e_cur && e_total && seg_ptr && coef, "null pointer");
    SPB_CHECK_RANGES();
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SPB_F64)
        return mix_dispatch<double>(g, e_cur, nullptr, 1, nullptr, kStoreLocal, e_total,
                                    seg_ptr, coef, n_patches, n_alloc, n_classes, n_dirs,
                                    n_bands, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
    if (dtype == SPB_F32)
        return mix_dispatch<float>(g, e_cur, nullptr, 1, nullptr, kStoreLocal, e_total,
                                   seg_ptr, coef, n_patches, n_alloc, n_classes, n_dirs,
                                   n_bands, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
    return fail(-1, "invalid argument", "dtype");
}

int spb_exchange_mix_fused(const void *g, const uint64_t *cur_ptrs_h, int n_peers,
                           void *cur_multicast, void *e_total, const int64_t *seg_ptr,
                           const void *coef, int64_t n_patches, int64_t n_alloc,
                           int64_t n_classes, int64_t n_dirs, int64_t n_bands, int64_t b_lo,
                           int64_t b_hi, int64_t j_lo, int64_t j_hi, int64_t t_pad,
                           int64_t ld, int64_t pad, int dtype, void *stream) {
    SPB_REQUIRE(g && e_total && seg_ptr && coef, "null pointer");
    SPB_REQUIRE(cur_multicast || (cur_ptrs_h && n_peers >= 1 && n_peers <= kMaxPeers),
                "need a multicast address or 1..8 peer pointers");
    SPB_CHECK_RANGES();
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = cur_multicast ? kStoreMulticast : kStorePeers;
    if (dtype == SPB_F64)
        return mix_dispatch<double>(g, nullptr, cur_ptrs_h, n_peers, cur_multicast, mode,
                                    e_total, seg_ptr, coef, n_patches, n_alloc, n_classes,
                                    n_dirs, n_bands, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
    if (dtype == SPB_F32)
        return mix_dispatch<float>(g, nullptr, cur_ptrs_h, n_peers, cur_multicast, mode,
                                   e_total, seg_ptr, coef, n_patches, n_alloc, n_classes,
                                   n_dirs, n_bands, b_lo, b_hi, j_lo, j_hi, t_pad, ld, pad, st);
    return fail(-1, "invalid argument", "dtype");
}

int spb_energy_exchange(const void *e0, const int32_t *delay0, const int64_t *seg_ptr,
                        const int32_t *src, const void *wgt, const int32_t *dly,
                        const int64_t *ent_ptr, const void *recs, const void *coef,
                        int64_t n_patches, int64_t n_classes, int64_t n_dirs,
                        int64_t n_bands, int64_t n_samples, int64_t t_pad, int64_t pad,
                        int64_t max_order, void *e_total, void *e_a, void *e_b, void *g,
                        int dtype, void *stream) {
    const int64_t ld = pad + t_pad;
    const int64_t n = n_patches;
    int rc = spb_exchange_init(e_total, max_order >= 1 ? e_a : nullptr, e0, delay0, n, n,
                               n_dirs, n_bands, n_samples, ld, pad, dtype, stream);
    if (rc) return rc;
    if (max_order < 1) return 0;
    SPB_REQUIRE(e_a && e_b && g, "null workspace");
    // e_b's pre-roll must be zero as well (e_a was zeroed by init)
    const size_t esz = dtype == SPB_F64 ? 8 : 4;
    SPB_CUDA(cudaMemsetAsync(e_b, 0, esz * n * n_dirs * n_bands * ld, (cudaStream_t)stream));
    void *prev = e_a, *cur = e_b;
    for (int64_t k = 0; k < max_order; ++k) {
        if (recs)
            rc = spb_exchange_gather_tiled(prev, g, ent_ptr, recs, nullptr, n, n, n_classes, n_dirs,
                                           n_bands, 0, n_bands, 0, n, t_pad, ld, pad, dtype,
                                           stream);
        else
            rc = spb_exchange_gather(prev, g, seg_ptr, src, wgt, dly, n, n, n_classes, n_dirs,
                                     n_bands, 0, n_bands, 0, n, t_pad, ld, pad, dtype, stream);
        if (rc) return rc;
        rc = spb_exchange_mix(g, cur, e_total, seg_ptr, coef, n, n, n_classes, n_dirs, n_bands,
                              0, n_bands, 0, n, t_pad, ld, pad, dtype, stream);
        if (rc) return rc;
        void *tmp = prev; prev = cur; cur = tmp;
    }
    return 0;
}

int spb_collect_mono(const void *e_total, const int32_t *rdir, const int32_t *shift,
                     const void *scale, int64_t n_receivers, int64_t n_patches,
                     int64_t n_alloc, int64_t n_dirs, int64_t n_bands, int64_t n_samples,
                     int64_t ld, int64_t pad, void *mono, void *partial, int64_t n_split,
                     int dtype, void *stream) {
    SPB_REQUIRE(e_total && rdir && shift && scale && mono && partial, "null pointer");
    SPB_REQUIRE(n_split >= 1 && n_split <= 65535, "n_split");
    SPB_REQUIRE(n_receivers * n_bands <= 65535, "n_receivers * n_bands > 65535: batch the receivers");
    if (n_receivers == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div(n_samples, 256), (unsigned)(n_receivers * n_bands),
              (unsigned)n_split);
    const int64_t n_out = n_receivers * n_bands * n_samples;
    if (dtype == SPB_F64) {
        k_collect_partial<double><<<grid, 256, 0, st>>>(
            (const double *)e_total, rdir, shift, (const double *)scale, n_patches, n_alloc,
            n_dirs, n_bands, n_samples, ld, pad, (double *)partial, n_split);
        int rc = check_launch("k_collect_partial");
        if (rc) return rc;
        k_collect_reduce<double><<<(unsigned)ceil_div(n_out, 256), 256, 0, st>>>(
            (const double *)partial, (double *)mono, n_out, n_split);
        return check_launch("k_collect_reduce");
    }
    if (dtype == SPB_F32) {
        k_collect_partial<float><<<grid, 256, 0, st>>>(
            (const float *)e_total, rdir, shift, (const float *)scale, n_patches, n_alloc,
            n_dirs, n_bands, n_samples, ld, pad, (float *)partial, n_split);
        int rc = check_launch("k_collect_partial");
        if (rc) return rc;
        k_collect_reduce<float><<<(unsigned)ceil_div(n_out, 256), 256, 0, st>>>(
            (const float *)partial, (float *)mono, n_out, n_split);
        return check_launch("k_collect_reduce");
    }
    return fail(-1, "invalid argument", "dtype");
}

int spb_collect_mono_staged(const void *e_total, const int32_t *shift, const void *scale,
                            int64_t n_receivers, int64_t n_patches, int64_t n_alloc,
                            int64_t n_bands, int64_t n_samples, int64_t ld, int64_t pad,
                            void *mono, void *partial, int64_t n_split, int n_stages,
                            int dtype, void *stream) {
    SPB_REQUIRE(e_total && shift && scale && mono && partial, "null pointer");
    SPB_REQUIRE(n_split >= 1 && n_split <= 65535, "n_split");
    SPB_REQUIRE(n_samples >= 1 && n_samples < (1 << 30), "n_samples");
    SPB_REQUIRE(dtype == SPB_F64 || dtype == SPB_F32, "dtype");
    if (n_receivers == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // 8 receivers x 4 bins per thread, two CTAs per SM: the fastest of the shapes measured
    // (16x4, 8x8, two sets of 8x4 in a 512-thread CTA: profiles/r02_sweep_collect_*.jsonl)
    if (dtype == SPB_F64)
        return collect_staged_t<double, 8, 4>(e_total, shift, scale, n_receivers, n_patches,
                                              n_alloc, n_bands, n_samples, ld, pad, mono, partial,
                                              n_split, n_stages, st);
    return collect_staged_t<float, 8, 4>(e_total, shift, scale, n_receivers, n_patches, n_alloc,
                                         n_bands, n_samples, ld, pad, mono, partial, n_split,
                                         n_stages, st);
}

int spb_collect_patchwise(const void *e_total, const int32_t *rdir, const int32_t *shift,
                          const void *scale, int64_t n_receivers, int64_t n_patches,
                          int64_t n_alloc, int64_t n_dirs, int64_t n_bands, int64_t n_samples,
                          int64_t ld, int64_t pad, void *out, int dtype, void *stream) {
    SPB_REQUIRE(e_total && rdir && shift && scale && out, "null pointer");
    const int64_t rows = n_receivers * n_patches * n_bands;
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SPB_REQUIRE(rows <= 2147483647LL, "too many (receiver, patch, band) rows");
    dim3 grid((unsigned)rows, (unsigned)ceil_div(n_samples, 256));
    if (dtype == SPB_F64)
        k_collect_patchwise<double><<<grid, 256, 0, st>>>(
            (const double *)e_total, rdir, shift, (const double *)scale, n_patches, n_alloc,
            n_dirs, n_bands, n_samples, ld, pad, (double *)out);
    else if (dtype == SPB_F32)
        k_collect_patchwise<float><<<grid, 256, 0, st>>>(
            (const float *)e_total, rdir, shift, (const float *)scale, n_patches, n_alloc,
            n_dirs, n_bands, n_samples, ld, pad, (float *)out);
    else
        return fail(-1, "invalid argument", "dtype");
    return check_launch("k_collect_patchwise");
}

}  // extern "C"
