/*
 * sparrow_b200.h -- C ABI of the B200-native DirectionalRadiosityFast hot path.
 *
 * The reference (sparrow-acoustics/sparrowpy v1.0.1) has no FFI layer; its
 * operator boundary is the set of numba-jitted array functions that the class
 * methods call by name ("Seam B").  Every entry point below
 * replaces one of those functions and cites it.  A maintainer binds them with
 * ctypes.
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions across the boundary;
 *   - every pointer is a DEVICE pointer owned by the caller (e.g. a torch tensor's
 *     data_ptr()) unless its name ends in _h (host); nothing is allocated behind
 *     the caller's back -- scratch space is passed in explicitly;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*), no
 *     implicit synchronisation;
 *   - return value 0 = success, negative = error; spb_last_error() gives the
 *     message of the calling thread's last failure;
 *   - dtype: SPB_F64 = 0 (reference precision), SPB_F32 = 1 (histograms in fp32).
 *
 * Energy histogram layout in HBM ("padded rows", band-major):
 *   E[row][PAD + t],  row = (band * n_alloc + patch) * D + direction,
 *   row stride LD = PAD + T_pad elements.  Bands are independent through the whole
 *   recursion, so each band is one contiguous block (per-band all-gathers can
 *   overlap the other bands' kernels).  The PAD leading elements of every row
 *   are a zero pre-roll (t < 0), PAD >= the largest pair delay, so a delayed read
 *   E[row][PAD + t - delay] never needs a bounds test.  T_pad >= T is the tile
 *   multiple the caller got from spb_exchange_layout().
 */
#ifndef SPARROW_B200_H
#define SPARROW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPB_F64 0
#define SPB_F32 1

int spb_version(void);
const char *spb_last_error(void);
/* number of SMs / device name of the current device (sanity + grid sizing) */
int spb_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------------------
 * Energy exchange (reference RadiosityFast.py:1037-1145, `_energy_exchange` and
 * `_energy_exchange_init_energy`).
 *
 * Factored form of form_factors_tilde (RadiosityFast.py:1234-1272): a directed
 * visible pair q = (sender i -> receiver j) carries
 *     weight ff_q, delay bin dly_q, sender row src_q = i*D + out_dir(i,j),
 *     class  c_q = brdf_index[wall(i)] * S + in_dir(i,j)
 * and the dense tensor is tilde[i,j,d,b] = ff_q * coef[c_q, d, b] with
 *     coef[c,d,b] = exp(-air[b]) * brdf[c,d,b].
 * Pairs are stored receiver-major: segment s = c*N + j owns entries
 * seg_ptr[s] .. seg_ptr[s+1].
 *
 * One reflection order = gather (stage 1) followed by mix (stage 2):
 *     G[b,c,j,t]     = sum_{q in seg(c,j)} ff_q * E_prev[b, src_q, t - dly_q]
 *     E_cur[b,j,d,t] = sum_c coef[c,d,b] * G[b,c,j,t];   E_total += E_cur
 * ------------------------------------------------------------------------- */

/* Tile geometry the kernels expect: T_pad (multiple of the 256-bin time tile) and
 * PAD (multiple of 32, > max_delay + 32).  LD = PAD + T_pad. */
int spb_exchange_layout(int64_t n_samples, int64_t max_delay, int dtype,
                        int64_t *t_pad, int64_t *pad);

/* `_energy_exchange_init_energy` (RadiosityFast.py:1037-1070): zero e_total and
 * e_prev (all n_alloc*D*B rows), then add e0[i,d,b] at bin delay0[i] of row
 * (b, i, d).  Energy whose bin is >= n_samples is dropped (the reference writes out
 * of bounds).  e0: [N, D, B]; delay0: [N] int32; e_prev may be 0.
 * n_alloc >= N is the number of patches the histogram buffers are allocated for
 * (> N only when shards are padded to equal size, see distributed.py). */
int spb_exchange_init(void *e_total, void *e_prev, const void *e0,
                      const int32_t *delay0, int64_t n_patches, int64_t n_alloc,
                      int64_t n_dirs, int64_t n_bands, int64_t n_samples, int64_t ld,
                      int64_t pad, int dtype, void *stream);

/* The scatter step of spb_exchange_init alone, into the band window
 * [band_lo, band_lo + n_bands_src) of buffers that hold n_bands_total bands and were
 * zeroed by the caller.  Used to batch several sources: source s occupies the bands
 * [s*B, (s+1)*B) -- sources are just more independent channels of the exchange.
 * e0: [N, D, n_bands_src]. */
int spb_exchange_scatter(void *e_total, void *e_prev, const void *e0,
                         const int32_t *delay0, int64_t n_patches, int64_t n_alloc,
                         int64_t n_dirs, int64_t n_bands_src, int64_t band_lo,
                         int64_t n_bands_total, int64_t n_samples, int64_t ld,
                         int64_t pad, int dtype, void *stream);

/* Stage 1 of one order for receiver patches [j_lo, j_hi) and bands [b_lo, b_hi)
 * (everything on one GPU; a receiver shard, or one band of a pipelined schedule, on
 * several).  g: [B, C*N, LD], only rows of non-empty segments in the range are
 * written.  seg_ptr: [C*N + 1] int64; src, dly: int32; wgt: dtype.  CSR kernel. */
int spb_exchange_gather(const void *e_prev, void *g, const int64_t *seg_ptr,
                        const int32_t *src, const void *wgt, const int32_t *dly,
                        int64_t n_patches, int64_t n_alloc, int64_t n_classes,
                        int64_t n_dirs, int64_t n_bands, int64_t b_lo, int64_t b_hi,
                        int64_t j_lo, int64_t j_hi, int64_t t_pad, int64_t ld,
                        int64_t pad, int dtype, void *stream);

/* Stage 1 with TMA-staged energy tiles (the fast path).  Same result as
 * spb_exchange_gather; the pair list is stored per tile of R neighbouring receivers
 * (R, the delay bucket and the record size come from spb_tile_geometry) as the union
 * of the tile's senders: tile = class * ceil(N/R) + j / R owns records
 * ent_ptr[tile] .. ent_ptr[tile+1]; a record is
 *     { dtype w[R]; uint8 rel[R]; int32 src; int32 dmin_and_mask; }
 * = weights and (delay - dmin) of the R receivers for sender row src; dmin (low 24
 * bits, a multiple of the bucket) and a reload mask (high 8 bits: slot s starts a
 * new shift; empty slots have w = 0 and repeat the previous shift).  j_lo must be a
 * multiple of R.  cta_order (optional, may be 0): a permutation of the
 * n_classes * ceil((j_hi - j_lo)/R) local tiles (index = class * n_tile_cols + tile
 * column) giving the launch order, e.g. longest record lists first. */
int spb_tile_geometry(int dtype, int64_t *receivers_per_tile, int64_t *delay_bucket,
                      int64_t *record_bytes);
int spb_exchange_gather_tiled(const void *e_prev, void *g, const int64_t *ent_ptr,
                              const void *recs, const int32_t *cta_order,
                              int64_t n_patches, int64_t n_alloc,
                              int64_t n_classes, int64_t n_dirs, int64_t n_bands,
                              int64_t b_lo, int64_t b_hi, int64_t j_lo, int64_t j_hi,
                              int64_t t_pad, int64_t ld, int64_t pad, int dtype,
                              void *stream);

/* Window records of the tensor-memory gather below (80 bytes):
 *     { double w[R]; uint8 off[R]; int32 src; int32 dbase; }
 * one per (tile of R = 8 receivers, sender row, delay window [dbase, dbase + W]), dbase
 * even, off = 2 * (delay - dbase) = the receiver's tensor-memory column offset; a slot
 * without a pair has w = 0 and off = 0; W is one of the instantiated widths (4 or 10).
 * Every tile's list is padded to a multiple of spb_tmem_batch() records with null records
 * (all zero).  Tile numbering, ent_ptr, j_lo and cta_order as for
 * spb_exchange_gather_tiled. */
int spb_window_geometry(int dtype, int64_t *receivers_per_tile, int64_t *max_window,
                        int64_t *record_bytes);

/* Stage 1, tensor-memory variant (FP64 only; same result as spb_exchange_gather, window
 * records as described above).  The sender window of a record
 * is staged by a 2-D TMA tensor load into shared memory and from there into TENSOR
 * MEMORY with time along the TMEM columns, 16 consecutive bins (+ the delay window) per
 * TMEM lane; a receiver's delay then is a dynamic column address of tcgen05.ld, which
 * feeds the FMAs from a datapath three times wider than shared memory.
 * Replaces the inner loops of _energy_exchange (RadiosityFast.py:1121-1144).  e_prev
 * must be the base of the whole (n_bands * n_alloc * n_dirs, ld) histogram, 16-byte
 * aligned.  The pipeline hands records over in batches: every tile's record list must be
 * padded to a multiple of spb_tmem_batch() with null records (w = 0, rel = 255, src = 0,
 * dbase = 0). */
int spb_tmem_batch(void);
int spb_exchange_gather_tmem(const void *e_prev, void *g, const int64_t *ent_ptr,
                             const void *recs, const int32_t *cta_order,
                             int64_t n_patches, int64_t n_alloc, int64_t n_classes,
                             int64_t n_dirs, int64_t n_bands, int64_t b_lo, int64_t b_hi,
                             int64_t j_lo, int64_t j_hi, int64_t t_pad, int64_t ld,
                             int64_t pad, int64_t window, int dtype, void *stream);

/* One whole reflection order (stage 1 + stage 2, RadiosityFast.py:1121-1144) in ONE kernel for
 * scenes with a single BRDF class and a single direction (n_classes = n_dirs = 1, e.g. diffuse
 * walls): the tensor-memory gather scales its sums by coef[b] in its epilogue, adds them to
 * e_total and stores E_k into the n_peers buffers cur_ptrs_h[0..n_peers) (HOST array of device
 * addresses: this rank's ping-pong buffer and, for a receiver-sharded run, the peers' -- the
 * per-order exchange rides on the compute kernel, tile by tile).  Rows of receivers without
 * pairs are zeroed.  Same result, bit for bit, as spb_exchange_gather_tmem + spb_exchange_mix. */
int spb_exchange_order_fused(const void *e_prev, const uint64_t *cur_ptrs_h, int n_peers,
                             void *e_total, const void *coef, const int64_t *ent_ptr,
                             const void *recs, const int32_t *cta_order, int64_t n_patches,
                             int64_t n_alloc, int64_t n_bands, int64_t b_lo, int64_t b_hi,
                             int64_t j_lo, int64_t j_hi, int64_t t_pad, int64_t ld, int64_t pad,
                             int64_t window, int dtype, void *stream);

/* Stage 2 of one order for receiver patches [j_lo, j_hi), bands [b_lo, b_hi): BRDF
 * contraction, writes e_cur rows of those patches and accumulates them into
 * e_total.  coef: [C, D, B] in dtype. */
int spb_exchange_mix(const void *g, void *e_cur, void *e_total,
                     const int64_t *seg_ptr, const void *coef, int64_t n_patches,
                     int64_t n_alloc, int64_t n_classes, int64_t n_dirs, int64_t n_bands,
                     int64_t b_lo, int64_t b_hi, int64_t j_lo, int64_t j_hi,
                     int64_t t_pad, int64_t ld, int64_t pad, int dtype, void *stream);

/* Stage 2 fused with the per-order exchange of the receiver-sharded run: instead of
 * writing E_k into the local buffer and all-gathering it afterwards, the kernel
 * stores every E_k element of this rank's receivers straight into all ranks' copies
 * of the buffer -- with one NVSwitch multicast store (multimem.st) when
 * cur_multicast != 0, else with NVLink P2P stores to the n_peers buffers in
 * cur_ptrs_h (host array of device addresses, own rank included).  The buffers must
 * be symmetric allocations of identical size (e.g. torch symmetric memory); the
 * caller separates orders with a cross-rank barrier.  e_total stays local. */
int spb_exchange_mix_fused(const void *g, const uint64_t *cur_ptrs_h, int n_peers,
                           void *cur_multicast, void *e_total, const int64_t *seg_ptr,
                           const void *coef, int64_t n_patches, int64_t n_alloc,
                           int64_t n_classes, int64_t n_dirs, int64_t n_bands,
                           int64_t b_lo, int64_t b_hi, int64_t j_lo, int64_t j_hi,
                           int64_t t_pad, int64_t ld, int64_t pad, int dtype,
                           void *stream);

/* `_energy_exchange` (RadiosityFast.py:1073-1145) on one GPU: init + max_order
 * x (gather, mix).  e_a, e_b: ping-pong [B*N*D, LD]; g: [B*C*N, LD].  With recs != 0
 * stage 1 is the tiled TMA kernel, otherwise the CSR kernel.
 * max_order < 1 means "initial energy only" (RadiosityFast.py:550-555, :1119). */
int spb_energy_exchange(const void *e0, const int32_t *delay0,
                        const int64_t *seg_ptr, const int32_t *src, const void *wgt,
                        const int32_t *dly, const int64_t *ent_ptr, const void *recs,
                        const void *coef, int64_t n_patches, int64_t n_classes,
                        int64_t n_dirs, int64_t n_bands, int64_t n_samples,
                        int64_t t_pad, int64_t pad, int64_t max_order, void *e_total,
                        void *e_a, void *e_b, void *g, int dtype, void *stream);

/* ---------------------------------------------------------------------------
 * Receiver collection (reference RadiosityFast.py:686-752 `_collect_energy_patches`
 * + :1148-1185 `_collect_receiver_energy`):
 *   out[r,b,(t + shift[r,k]) mod T] += e_total[b, k, rdir[r,k], t] * scale[r,k,b]
 * scale = patch->receiver factor * exp(-air[b]*dist), shift = ceil-delay mod T
 * (the reference's np.roll is circular).  mono: [R, B, T] dense (no padding).
 * partial: scratch [n_split, R, B, T].
 * ------------------------------------------------------------------------- */
int spb_collect_mono(const void *e_total, const int32_t *rdir, const int32_t *shift,
                     const void *scale, int64_t n_receivers, int64_t n_patches,
                     int64_t n_alloc, int64_t n_dirs, int64_t n_bands, int64_t n_samples,
                     int64_t ld, int64_t pad, void *mono, void *partial, int64_t n_split,
                     int dtype, void *stream);

/* The same sum for MANY receivers of a diffuse scene (n_dirs = 1, so `rdir` is all zeros and
 * is not passed): every histogram row is staged once in shared memory and applied to a group
 * of 8 receivers from there (k_collect_staged) instead of being re-read per receiver -- the
 * reference loops over the receivers in Python (RadiosityFast.py:711).  n_stages: depth of the
 * cp.async ring, 0 = auto, else 2..4.  Fails (-1) when a row does not fit shared memory;
 * `partial` and the result layout are those of spb_collect_mono. */
int spb_collect_mono_staged(const void *e_total, const int32_t *shift, const void *scale,
                            int64_t n_receivers, int64_t n_patches, int64_t n_alloc,
                            int64_t n_bands, int64_t n_samples, int64_t ld, int64_t pad,
                            void *mono, void *partial, int64_t n_split, int n_stages,
                            int dtype, void *stream);

/* patch-wise variant (`collect_energy_receiver_patchwise`, RadiosityFast.py:660):
 * out: [R, N, B, T] dense. */
int spb_collect_patchwise(const void *e_total, const int32_t *rdir,
                          const int32_t *shift, const void *scale,
                          int64_t n_receivers, int64_t n_patches, int64_t n_alloc,
                          int64_t n_dirs, int64_t n_bands, int64_t n_samples, int64_t ld,
                          int64_t pad, void *out, int dtype, void *stream);

/* ---------------------------------------------------------------------------
 * Geometry baking.  All coordinates are FP64 (the reference's dtype); boolean and
 * integer outputs are bit-identical to the reference, floating-point outputs agree
 * to <= 1e-6 relative (measured ~1e-12).
 * ------------------------------------------------------------------------- */

/* Per-surface data that `_point_in_polygon` (geometry.py:614-686) derives from a
 * blocking polygon: rotation to the plane, rotated vertices, side normals.
 * surf_points: [M, 4, 3], surf_normals: [M, 3]; blockers: spb_blocker_bytes(M). */
size_t spb_blocker_bytes(int64_t m);
int spb_make_blockers(const double *surf_points, const double *surf_normals, int64_t m,
                      int nvert, void *blockers, void *stream);

/* `_check_patch2patch_visibility` (geometry.py:750-797): vis[i,j] (uint8, [N,N]) =
 * AND over all M surfaces of `_basic_visibility(c_i, c_j, surface)` for i < j, 0
 * elsewhere. */
int spb_visibility_p2p(const double *centers, int64_t n, const void *blockers, int64_t m,
                       uint8_t *vis, void *stream);

/* Same result as spb_visibility_p2p, evaluated hierarchically: the blockers are
 * grouped by plane (the patches of one wall), a group is decided with one evaluation
 * of the plane quantities, and only the members whose polygon the plane hit can touch
 * (found through a 2-D grid of cells over the wall's in-plane coordinates; bins along
 * the in-plane y axis for the rare query point whose +x ray grazes a horizontal edge)
 * are evaluated individually.  O(N^2 * walls) instead of O(N^3).  Tables from
 * sparrowpy_b200.bake.build_groups: groups: n_groups records of spb_group_bytes() bytes;
 * members: blocker indices; bin_ptr / bin_items: CSR of the y-bins and cells of all groups
 * (+ the per-bin first-strip index); strips: [lo, hi] pairs of the grazing y-ranges.
 * Two optional (nullable) accelerators that cannot change the result: own_in[i] (uint8, N,
 * from spb_visibility_own_in; 255 = unknown) = "centre i lies in polygon i", the one exact
 * polygon test nearly every pair repeats, memoised per patch; own_group[i] (int32, N; -1 =
 * none) = the group of blocker i, evaluated first for the pairs of patch i (the conjunction
 * over the blockers is order-independent, geometry.py:786-795). */
int spb_visibility_own_in(const double *centers, int64_t count, const void *blockers,
                          uint8_t *own_in, void *stream);
size_t spb_group_bytes(void);
int spb_visibility_p2p_grouped(const double *centers, int64_t n, const void *blockers,
                               const void *groups, int64_t n_groups, const int32_t *members,
                               const int32_t *bin_ptr, const int32_t *bin_items,
                               const double *strips, const uint8_t *own_in,
                               const int32_t *own_group, uint8_t *vis, void *stream);
/* Rows [row_lo, row_hi) of the same matrix into vis_rows ([row_hi - row_lo, N] uint8): the
 * unit of the bake when it is sharded over GPUs (pair tiles are independent). */
int spb_visibility_p2p_grouped_rows(const double *centers, int64_t n, const void *blockers,
                                    const void *groups, int64_t n_groups,
                                    const int32_t *members, const int32_t *bin_ptr,
                                    const int32_t *bin_items, const double *strips,
                                    const uint8_t *own_in, const int32_t *own_group,
                                    int64_t row_lo, int64_t row_hi, uint8_t *vis_rows,
                                    void *stream);
/* host twins of spb_make_blockers / spb_visibility_p2p_grouped (HOST pointers): the
 * same predicates compiled for the CPU; used by the CPU tests only */
int spb_make_blockers_host(const double *surf_points_h, const double *surf_normals_h,
                           int64_t m, void *blockers_h);
int spb_visibility_p2p_grouped_host(const double *centers_h, int64_t n, const void *blockers_h,
                                    const void *groups_h, int64_t n_groups,
                                    const int32_t *members_h, const int32_t *bin_ptr_h,
                                    const int32_t *bin_items_h, const double *strips_h,
                                    int64_t n_own, const int32_t *own_group_h, uint8_t *vis_h);

/* `_check_point2patch_visibility` (geometry.py:799-839) for a batch of points:
 * vis[r,j] ([R,N] uint8). */
int spb_visibility_pt2p(const double *points, int64_t n_points, const double *centers,
                        int64_t n, const void *blockers, int64_t m, uint8_t *vis,
                        void *stream);

/* `patch2patch_ff_universal` (universal.py:12-96): ff[p] for visible pair p =
 * (i, j), i < j.  The Stokes call handles every pair that shares no vertex and
 * flags the others (nusselt_flag[p] = 1); the Nusselt call integrates the flagged
 * pairs listed in todo (integration.py:116-289). */
int spb_form_factors_stokes(const double *pts, const double *areas, const int32_t *pairs,
                            int64_t n_pairs, double *ff, uint8_t *nusselt_flag, void *stream);
int spb_form_factors_nusselt(const double *pts, const double *normals, const int32_t *pairs,
                             const int64_t *todo, int64_t n_todo, double *ff, void *stream);

/* `_source2patch_energy_universal` (universal.py:98-147) + `_add_directional`
 * (RadiosityFast.py:988-1034): distance[j] (0 for invisible patches), e0[j,d,b];
 * energy[j,b] (before the BRDF, optional).  src: 3 doubles on the device. */
int spb_source_energy(const double *src, const double *centers, const double *pts,
                      const uint8_t *vis, const double *air, const int64_t *patch_to_wall,
                      const double *vi, int64_t n_in, const double *brdf,
                      const int64_t *brdf_index, int64_t n_out, int64_t n_bands, int64_t n,
                      double *distance, double *e0, double *energy, void *stream);
/* The same for a batch of `n_src` sources in ONE launch (extension, SURVEY.md 8f: the
 * reference allows one source, RadiosityFast.py:450-451): src (S,3), vis (S,N), distance (S,N),
 * e0 (S,N,D,B), energy (S,N,B) or NULL. */
int spb_source_energy_batch(const double *src, int64_t n_src, const double *centers,
                            const double *pts, const uint8_t *vis, const double *air,
                            const int64_t *patch_to_wall, const double *vi, int64_t n_in,
                            const double *brdf, const int64_t *brdf_index, int64_t n_out,
                            int64_t n_bands, int64_t n, double *distance, double *e0,
                            double *energy, void *stream);

/* Receiver side of `_collect_energy_patches` (RadiosityFast.py:711-748) for a batch
 * of receivers: factor[r,k] (universal.py:149-160), rdir[r,k] (:728-730), delay[r,k]
 * = ceil(dist/c/dt) (:1178-1179), shift = delay mod T, scale[r,k,b] = factor *
 * exp(-air[b] * dist). */
int spb_receiver_factors(const double *rcv, int64_t n_rcv, const double *centers,
                         const double *pts, const uint8_t *vis, const double *air,
                         const int64_t *patch_to_wall, const double *vo, int64_t n_out,
                         int64_t n_bands, int64_t n, double speed_of_sound, double dt,
                         int64_t n_samples, double *factor, int32_t *rdir, int32_t *delay,
                         int32_t *shift, double *scale, void *stream);

/* Per-pair tables behind form_factors_tilde: dist[p] (numpy 1-D norm model,
 * RadiosityFast.py:538-543), out_dir[2p+e] (:403-414), in_dir[2p+e] (:1386-1390);
 * directed entry 2p = lo->hi, 2p+1 = hi->lo. */
int spb_pair_geometry(const double *centers, const int64_t *patch_to_wall,
                      const int32_t *pairs, int64_t n_pairs, const double *vi, int64_t n_in,
                      const double *vo, int64_t n_out, double *dist, int32_t *out_dir,
                      int32_t *in_dir, void *stream);

/* int(dist / c / dt) (RadiosityFast.py:1067-1068, :1135-1136) */
int spb_delay_bins(const double *dist, int64_t count, double speed_of_sound, double dt,
                   int32_t *out, void *stream);

/* test probes of the device arithmetic model (x87 norm, `_basic_visibility`) */
int spb_probe_norms(const double *v, int64_t count, int dim, double *out, void *stream);
int spb_probe_basic_visibility(const double *a, const double *b, const void *blockers,
                               int64_t count, uint8_t *visible, uint8_t *in_a, uint8_t *in_b,
                               void *stream);

/* Measurement probe (no reference counterpart): launches a register-only FMA kernel of
 * `n_blocks` x 256 threads x 8 chains x `iters` FMAs in `dtype` and returns the flops it
 * will execute; the caller times it with CUDA events to obtain the FMA-pipe peak the
 * compute-bound kernels are reported against.  `scratch`: >= 8 device bytes. */
int spb_fma_peak(int dtype, int64_t iters, int64_t n_blocks, void *scratch,
                 double *flops_launched, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPARROW_B200_H */
