// Geometry baking kernels (sm_100a): visibility, form factors, point<->patch
// factors and the per-pair direction / distance tables.
//
// Compiled with --fmad=false: the integer/boolean outputs (visibility, direction
// indices, delay bins) must be bit-identical to the reference's, whose arithmetic
// is modelled in exact.cuh / x87.cuh.  Explicit fma() calls are the only FMAs.
//
// Replaces (reference file:line): geometry.py:750-909 (visibility),
// form_factor/universal.py:12-160 + integration.py:38-344 (form factors, point
// factors), RadiosityFast.py:403-414, :988-1034, :1277-1312, :1359-1390 (BRDF
// direction indices, directional source energy).
#include <vector>

#include "common.cuh"
#include "exact.cuh"
#include "vis_group.cuh"

namespace spb {

using exact::Blocker;
using exact::kBlockerDoubles;

#define SPB_PI 3.141592653589793

// ---------------------------------------------------------------------------
// blockers
// ---------------------------------------------------------------------------
__global__ void k_make_blockers(const double *__restrict__ pts,
                                const double *__restrict__ normals, int64_t m,
                                Blocker *__restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    double p[12], n[3];
    for (int k = 0; k < 12; ++k) p[k] = pts[12 * s + k];
    for (int k = 0; k < 3; ++k) n[k] = normals[3 * s + k];
    Blocker b;
    exact::make_blocker(p, n, b);
    out[s] = b;
}

// ---------------------------------------------------------------------------
// visibility: geometry.py:750-839
// ---------------------------------------------------------------------------
constexpr int kVisThreads = 128;
constexpr int kVisTile = 64;      // blockers staged in shared memory per step

// Per-blocker data of the fast path, rebuilt per CTA while a tile is staged: the
// view point A is the same for every thread of the CTA (row i / the evaluation
// point), so everything that depends on (A, blocker) only is computed once per
// (CTA, blocker) instead of once per thread.
struct BlockerHead {
    double n[3];
    double s0[3];
    double dA;          // DOT(A - S0, n), exactly as exact::blocked computes it
    double cop_a;       // 1.0 when |dA| <= eta (A lies in the blocker's plane)
    double clr_a;       // exact::ray_clearance of A's in-plane projection
    double pad_;
};
constexpr int kHeadDoubles = sizeof(BlockerHead) / sizeof(double);

__device__ __noinline__ bool blocked_slow(const double *A, const double *B, const double *v,
                                          double vlen, bool cull_ok, const Blocker *k) {
    return exact::blocked(A, B, v, vlen, cull_ok, *k);
}

// Shared loop: AND over all blockers of "not blocked", staged through smem.
// `first_a/first_b`: blocker indices tested first (or -1) -- only an ordering
// heuristic, the conjunction does not depend on order.
//
// The inner loop decides the common cases inline with the rules proved in
// exact::blocked (same arithmetic for every value the reference compares):
//   * neither end point in the blocker's plane: no plane hit (|dp| <= 1e-6) or hit
//     parameter clearly outside the open segment (rule (a));
//   * an end point in the plane with positive exact::ray_clearance: it is not "in the
//     surface"; if the other end is off the plane the hit lies within eta*|v|/|dp| of
//     that end point and misses the polygon too when the clearance exceeds that
//     distance (rule (b)); if both ends are in the plane there is no plane hit.
// Everything else goes to exact::blocked.
__device__ __forceinline__ bool visible_against_all(const double *A, const double *B,
                                                    bool active,
                                                    const Blocker *__restrict__ blockers,
                                                    int64_t m, int64_t first_a,
                                                    int64_t first_b, double *sm) {
    double v[3];
    exact::sub3(B, A, v);
    const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double vlen = sqrt(vv);
    const bool cull_ok = vv > 1e-6;            // see exact::blocked
    bool visible = active;
    if (visible && first_a >= 0)
        visible = !blocked_slow(A, B, v, vlen, cull_ok, blockers + first_a);
    if (visible && first_b >= 0)
        visible = !blocked_slow(A, B, v, vlen, cull_ok, blockers + first_b);
    const Blocker *tile = reinterpret_cast<const Blocker *>(sm);
    BlockerHead *heads = reinterpret_cast<BlockerHead *>(sm + kVisTile * kBlockerDoubles);
    for (int64_t s0 = 0; s0 < m; s0 += kVisTile) {
        if (!__syncthreads_or(visible)) break;          // whole CTA decided
        const int cnt = (int)min((int64_t)kVisTile, m - s0);
        const double *src = reinterpret_cast<const double *>(blockers + s0);
        for (int k = threadIdx.x; k < cnt * kBlockerDoubles; k += blockDim.x) sm[k] = src[k];
        if ((int)threadIdx.x < cnt) {
            const Blocker &bk = blockers[s0 + threadIdx.x];
            BlockerHead h;
            double wa[3];
            for (int c = 0; c < 3; ++c) { h.n[c] = bk.n[c]; h.s0[c] = bk.s0[c]; }
            exact::sub3(A, h.s0, wa);
            h.dA = exact::dot3(wa, h.n);
            h.cop_a = (fabs(h.dA) > exact::kEta) ? 0.0 : 1.0;
            h.clr_a = exact::ray_clearance(exact::dot3(bk.r0, A), exact::dot3(bk.r1, A), bk);
            h.pad_ = 0.0;
            heads[threadIdx.x] = h;
        }
        __syncthreads();
        if (visible) {
            for (int s = 0; s < cnt; ++s) {
                const BlockerHead &h = heads[s];
                double w[3];
                exact::sub3(B, h.s0, w);
                const double dB = exact::dot3(w, h.n);
                const double dp = exact::dot3(v, h.n);
                const bool cop_a = h.cop_a != 0.0;
                const bool cop_b = !(fabs(dB) > exact::kEta);
                const bool hit = fabs(dp) > 1e-6;
                bool decided = false;          // "not blocked" by the inline rules
                if (!cop_a && !cop_b) {
                    decided = !hit;
                } else {
                    double clr_b = 1.0;
                    if (cop_b) {
                        const Blocker &bk = tile[s];
                        clr_b = exact::ray_clearance(exact::dot3(bk.r0, B),
                                                     exact::dot3(bk.r1, B), bk);
                    }
                    const double clr_a = cop_a ? h.clr_a : 1.0;
                    if (clr_a > exact::kClearGuard && clr_b > exact::kClearGuard) {
                        // neither end point is in the surface
                        if (!hit) decided = true;
                        else if (cull_ok && (cop_a != cop_b)) {
                            const double slack = exact::kEta * vlen / fabs(dp) + 3e-9;
                            decided = (cop_a ? clr_a : clr_b) > slack;
                        }
                    }
                }
                if (!decided && hit && cull_ok) {           // rule (a)
                    const double u = -dB;
                    const bool outside = dp > 0 ? (u > 2e-3 * dp || u < -1.002 * dp)
                                                : (u < 2e-3 * dp || u > -1.002 * dp);
                    // only valid when neither end point is in the surface
                    decided = outside && !cop_a && !cop_b;
                }
                if (!decided && blocked_slow(A, B, v, vlen, cull_ok, tile + s)) {
                    visible = false;
                    break;
                }
            }
        }
    }
    return visible;
}

// one CTA per (row i, chunk of j); only j > i is evaluated (geometry.py:775-779)
__global__ void __launch_bounds__(kVisThreads)
k_vis_p2p(const double *__restrict__ centers, int64_t n,
          const Blocker *__restrict__ blockers, int64_t m, int64_t chunks_per_row,
          uint8_t *__restrict__ vis) {
    __shared__ double sm[kVisTile * (kBlockerDoubles + kHeadDoubles)];
    const int64_t i = blockIdx.x / chunks_per_row;
    const int64_t j0 = (blockIdx.x % chunks_per_row) * kVisThreads;
    if (j0 + kVisThreads - 1 <= i) return;               // chunk entirely at j <= i
    const int64_t j = j0 + threadIdx.x;
    const bool active = j > i && j < n;
    double A[3], B[3];
    for (int k = 0; k < 3; ++k) {
        A[k] = centers[3 * i + k];
        B[k] = active ? centers[3 * j + k] : 0.0;
    }
    const bool self = (m == n);   // blockers are the patches themselves: try i, j first
    const bool v = visible_against_all(A, B, active, blockers, m, self && active ? i : -1,
                                       self && active ? j : -1, sm);
    if (active) vis[i * n + j] = v ? 1 : 0;
}

// points x patches (geometry.py:799-839), vis_point = the point
__global__ void __launch_bounds__(kVisThreads)
k_vis_pt2p(const double *__restrict__ points, const double *__restrict__ centers, int64_t n,
           const Blocker *__restrict__ blockers, int64_t m, int64_t chunks_per_row,
           uint8_t *__restrict__ vis) {
    __shared__ double sm[kVisTile * (kBlockerDoubles + kHeadDoubles)];
    const int64_t r = blockIdx.x / chunks_per_row;
    const int64_t j = (blockIdx.x % chunks_per_row) * kVisThreads + threadIdx.x;
    const bool active = j < n;
    double A[3], B[3];
    for (int k = 0; k < 3; ++k) {
        A[k] = points[3 * r + k];
        B[k] = active ? centers[3 * j + k] : 0.0;
    }
    const bool v = visible_against_all(A, B, active, blockers, m, -1, -1, sm);
    if (active) vis[r * n + j] = v ? 1 : 0;
}

// hierarchical variant: one CTA per (row i, chunk of j), loop over blocker groups
// (vis_group.cuh); the group headers are staged in shared memory when they fit
constexpr int kMaxGroupsSmem = 128;

// memo for the grouped kernel: is centre i inside polygon i (vis_group.cuh, exact::Own)
__global__ void k_own_in(const double *__restrict__ centers, int64_t count,
                         const Blocker *__restrict__ blockers, uint8_t *__restrict__ own_in) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) own_in[i] = exact::own_in_polygon(centers + 3 * i, blockers[i]) ? 1 : 0;
}

// 4 resident CTAs per SM (128 registers, 268 bytes of spills): measured 79.4 ms against
// 92.3 ms with the 167 registers ptxas takes unbounded (C4, profiles/r02_sweep_vis_c4.jsonl)
__global__ void __launch_bounds__(kVisThreads, 4)
k_vis_p2p_grouped(const double *__restrict__ centers, int64_t n,
                  const Blocker *__restrict__ blockers, const exact::Group *__restrict__ groups,
                  int32_t n_groups, const int32_t *__restrict__ members,
                  const int32_t *__restrict__ bin_ptr, const int32_t *__restrict__ bin_items,
                  const double *__restrict__ strips, const uint8_t *__restrict__ own_in,
                  const int32_t *__restrict__ own_group, int64_t chunks_per_row, int64_t row_lo,
                  uint8_t *__restrict__ vis) {
    // rows [row_lo, row_lo + gridDim.x / chunks_per_row) of the matrix; vis holds those rows
    __shared__ exact::Group sg[kMaxGroupsSmem];
    const int64_t i = row_lo + blockIdx.x / chunks_per_row;
    const int64_t j0 = (blockIdx.x % chunks_per_row) * kVisThreads;
    if (j0 + kVisThreads - 1 <= i) return;
    const bool staged = n_groups <= kMaxGroupsSmem;
    if (staged) {
        const double *src = reinterpret_cast<const double *>(groups);
        double *dst = reinterpret_cast<double *>(sg);
        const int words = n_groups * (int)(sizeof(exact::Group) / sizeof(double));
        for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    const int64_t j = j0 + threadIdx.x;
    if (!(j > i && j < n)) return;
    double A[3], B[3], v[3];
    for (int k = 0; k < 3; ++k) { A[k] = centers[3 * i + k]; B[k] = centers[3 * j + k]; }
    exact::sub3(B, A, v);
    const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double vlen = sqrt(vv);
    const bool cull_ok = vv > 1e-6;
    // the conjunction over the groups does not depend on their order: the walls the two
    // patches lie on come first (they decide most invisible pairs: same wall, or the other
    // patch behind this one's wall), with the memoised "centre in its own polygon" values
    const exact::Own own = exact::make_own(i, j, own_in);
    const int32_t g_a = own_group ? own_group[i] : -1, g_b = own_group ? own_group[j] : -1;
    bool visible = true;
    for (int32_t step = -2; step < n_groups && visible; ++step) {      // one call site
        const int32_t g = step == -2 ? g_a : (step == -1 ? g_b : step);
        if (g < 0 || (step >= -1 && g == g_a) || (step >= 0 && g == g_b)) continue;
        const exact::Group &grp = staged ? sg[g] : groups[g];
        visible = !exact::group_blocked(A, B, v, vlen, cull_ok, grp, blockers, members, bin_ptr,
                                        bin_items, strips, own);
    }
    vis[(i - row_lo) * n + j] = visible ? 1 : 0;
}

// ---------------------------------------------------------------------------
// form factors (tolerance path, FP64)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double nrm3p(const double *v) {
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}
__device__ __forceinline__ double nrm2p(const double *v) {
    return sqrt(v[0] * v[0] + v[1] * v[1]);
}

// integration.py:608-651, point `idx` of a boundary sampled with n_div per edge
__device__ __forceinline__ void boundary_point(const double *el, int idx, int n_div,
                                               double *out) {
    const int e = idx / n_div, ii = idx - e * n_div;
    const double *p0 = el + 3 * e, *p1 = el + 3 * ((e + 1) & 3);
    for (int k = 0; k < 3; ++k) out[k] = p0[k] + (double)ii * (p1[k] - p0[k]) / (double)n_div;
}

__device__ __forceinline__ double boole(double x0, double x1, const double *y) {
    const double h = x1 - x0;                                       // integration.py:481-486
    return 2 * h / 45 * (7 * y[0] + 32 * y[1] + 12 * y[2] + 32 * y[3] + 7 * y[4]);
}

__device__ __forceinline__ double sgn(double v) { return (double)((v > 0) - (v < 0)); }

// integration.py:416-460 (+ :349-412): quadratic through three samples, closed form
__device__ double area_under_curve(const double ps[3][2]) {
    const double f[2] = {ps[2][0] - ps[0][0], ps[2][1] - ps[0][1]};
    const double nf = nrm2p(f);
    const double r0[2] = {f[0] / nf, f[1] / nf};
    const double r1[2] = {-f[1] / nf, f[0] / nf};
    double x[3] = {0, 0, 0}, y[3] = {0, 0, 0};
    for (int k = 1; k < 3; ++k) {
        const double c[2] = {ps[k][0] - ps[0][0], ps[k][1] - ps[0][1]};
        x[k] = exact::dot2(r0, c);
        y[k] = exact::dot2(r1, c);
    }
    if (fabs(x[2] - x[0]) < 1e-6) return 0.0;
    const double det = x[1] * x[2] * (x[1] - x[2]);
    const double c0 = (y[1] * x[2] - y[2] * x[1]) / det;
    const double c1 = (y[2] * x[1] * x[1] - y[1] * x[2] * x[2]) / det;
    return c0 * (x[2] * x[2] * x[2]) / 3 + c1 * (x[2] * x[2]) / 2;
}

// integration.py:116-230
__device__ double nusselt_analog(const double *o, const double *R /*rot(n_i)*/,
                                 const double *pj, double hand) {
    double sph[8][3], pln[8][2];
    for (int i = 0; i < 8; ++i) {
        double bp[3], d[3];
        boundary_point(pj, i, 2, bp);
        exact::sub3(bp, o, d);
        const double nd = nrm3p(d);
        for (int k = 0; k < 3; ++k) sph[i][k] = d[k] / nd;
        pln[i][0] = exact::dot3(R, sph[i]);
        pln[i][1] = exact::dot3(R + 3, sph[i]);
    }
    // _polygon_area of the 4 projected vertices (z = 0): |cross| = |2-D cross|
    double big = 0.0;
    for (int t = 0; t < 2; ++t) {
        const double ax = pln[2 * (t + 1)][0] - pln[0][0], ay = pln[2 * (t + 1)][1] - pln[0][1];
        const double bx = pln[2 * (t + 2)][0] - pln[0][0], by = pln[2 * (t + 2)][1] - pln[0][1];
        big += .5 * fabs(ax * by - ay * bx);
    }
    double curved = 0.0;
    for (int jj = 0; jj < 4; ++jj) {
        const int s0 = 2 * jj, s1 = 2 * jj + 1, s2 = (2 * jj + 2) & 7;
        const double cz = pln[s2][0] * pln[s0][1] - pln[s2][1] * pln[s0][0];
        if (fabs(cz) > 1e-6) {
            if (exact::dot2(pln[s2], pln[s0]) >= 1e-6) {
                const double ps[3][2] = {{pln[s0][0], pln[s0][1]}, {pln[s1][0], pln[s1][1]},
                                         {pln[s2][0], pln[s2][1]}};
                curved += area_under_curve(ps);
            } else {
                double mp[3], marc[3], a[3], b[3];
                for (int k = 0; k < 3; ++k) mp[k] = sph[s0][k] + (sph[s2][k] - sph[s0][k]) / 2;
                const double nm = nrm3p(mp);
                for (int k = 0; k < 3; ++k) marc[k] = mp[k] / nm;
                for (int k = 0; k < 3; ++k) {
                    a[k] = sph[s0][k] + (marc[k] - sph[s0][k]) / 2;
                    b[k] = marc[k] + (sph[s2][k] - marc[k]) / 2;
                }
                const double mp2[2] = {exact::dot3(R, mp), exact::dot3(R + 3, mp)};
                const double marc2[2] = {exact::dot3(R, marc), exact::dot3(R + 3, marc)};
                const double na = nrm3p(a), nb = nrm3p(b);
                for (int k = 0; k < 3; ++k) { a[k] = a[k] / na; b[k] = b[k] / nb; }
                const double a2[2] = {exact::dot3(R, a), exact::dot3(R + 3, a)};
                const double b2[2] = {exact::dot3(R, b), exact::dot3(R + 3, b)};
                const double d1[2] = {pln[s2][0] - pln[s0][0], pln[s2][1] - pln[s0][1]};
                const double d2[2] = {mp2[0] - marc2[0], mp2[1] - marc2[1]};
                const double lin = nrm2p(d1) * nrm2p(d2) / 2;
                const double ls[3][2] = {{pln[s0][0], pln[s0][1]}, {a2[0], a2[1]},
                                         {marc2[0], marc2[1]}};
                const double rs[3][2] = {{marc2[0], marc2[1]}, {b2[0], b2[1]},
                                         {pln[s2][0], pln[s2][1]}};
                const double left = area_under_curve(ls);
                const double right = area_under_curve(rs);
                curved += (lin * sgn(left) + left + right);
            }
        }
    }
    return big + hand * curved;
}

// universal.py:12-96, Stokes branch (integration.py:38-114); Nusselt pairs are flagged for
// the warp kernel.  16 lanes per pair (two pairs per warp): lane a owns boundary point a of
// patch i -- its 16 logarithms log|p_a - q_b| and the three inner Boole sums over the
// boundary of patch j -- and the outer Boole sum over the boundary of patch i is a weighted
// shuffle reduction over the 16 lanes (a point at a corner closes one segment and opens the
// next: weight 7 in both).  No per-thread tables, no local memory.
__global__ void __launch_bounds__(128, 4)
k_ff_stokes(const double *__restrict__ pts, const double *__restrict__ areas,
            const int32_t *__restrict__ pairs, int64_t n_pairs, double *__restrict__ ff,
            uint8_t *__restrict__ nusselt_flag) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int a = threadIdx.x & 15;
    const unsigned half_mask = 0xffffu << (threadIdx.x & 16);
    const bool live = (gid >> 4) < n_pairs;
    const int64_t p = live ? (gid >> 4) : n_pairs - 1;     // idle lanes shadow the last pair
    const int64_t i = pairs[2 * p], j = pairs[2 * p + 1];
    double pj[12];
    for (int k = 0; k < 12; ++k) pj[k] = pts[12 * j + k];
    // geometry.py:719-748: lane a compares vertex a / 4 of patch j with vertex a % 4 of patch i
    // (dynamic vertex indices address global memory, the register copies stay statically indexed)
    double dv[3];
    exact::sub3(pts + 12 * j + 3 * (a >> 2), pts + 12 * i + 3 * (a & 3), dv);
    const bool touch = __any_sync(half_mask, nrm3p(dv) < 1e-6);
    if (touch) {
        if (a == 0 && live) nusselt_flag[p] = 1;
        return;
    }
    double pa[3];
    boundary_point(pts + 12 * i, a, 4, pa);
    // Boole coefficient 2 h / 45 of segment s of a boundary in coordinate dim (0 where the
    // segment does not move in that coordinate, integration.py:86-95)
    auto seg_coef = [](const double *el, int s, int dim) {
        const double x0 = el[3 * s + dim], xl = el[3 * ((s + 1) & 3) + dim];
        return fabs(xl - x0) > 1e-3 ? 2 * ((x0 + 1.0 * (xl - x0) / 4.0) - x0) / 45 : 0.0;
    };
    double cj[4][3];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int dim = 0; dim < 3; ++dim) cj[s][dim] = seg_coef(pj, s, dim);
    // inner sums over the boundary of patch j, accumulated as the logarithms arrive: point b
    // is sample b % 4 of segment b / 4 and, at a corner, also sample 4 of the segment before
    double inner[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        double q[3], d[3];
        boundary_point(pj, b, 4, q);
        exact::sub3(pa, q, d);
        // log|p_a - q_b| as half the logarithm of the squared distance: the square root is a
        // fifth of the FP64 instructions of this kernel, which runs at the FP64 pipe's
        // instruction rate (tolerance path: differs from log(sqrt()) in the last bit only)
        const double yb = 0.5 * log(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const int s = b >> 2, k = b & 3;
        const double w = k == 0 ? 7.0 : (k == 2 ? 12.0 : 32.0);
#pragma unroll
        for (int dim = 0; dim < 3; ++dim) {
            double c = w * cj[s][dim];
            if (k == 0) c += 7.0 * cj[(s + 3) & 3][dim];
            inner[dim] += c * yb;
        }
    }
    // this lane's weight in the outer sum over the boundary of patch i: the same rule with a
    // (dynamic segment index: read from global memory)
    double contrib = 0.0;
    {
        const double *gi = pts + 12 * i;
        const int s = a >> 2, k = a & 3;
        const double w = k == 0 ? 7.0 : (k == 2 ? 12.0 : 32.0);
#pragma unroll
        for (int dim = 0; dim < 3; ++dim) {
            double c = w * seg_coef(gi, s, dim);
            if (k == 0) c += 7.0 * seg_coef(gi, (s + 3) & 3, dim);
            contrib += c * inner[dim];
        }
    }
    for (int off = 8; off > 0; off >>= 1) contrib += __shfl_xor_sync(half_mask, contrib, off);
    if (a == 0 && live) {
        nusselt_flag[p] = 0;
        ff[p] = fabs(contrib / (2 * SPB_PI * areas[i]));
    }
}

// integration.py:232-289 with :533-605: one warp per flagged pair, lanes over the
// regular surface samples of patch i, warp-shuffle reduction
__global__ void __launch_bounds__(128)
k_ff_nusselt(const double *__restrict__ pts, const double *__restrict__ normals,
             const int32_t *__restrict__ pairs, const int64_t *__restrict__ todo,
             int64_t n_todo, double *__restrict__ ff) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_todo) return;
    const int64_t p = todo[w];
    const int64_t i = pairs[2 * p], j = pairs[2 * p + 1];
    double pi[12], pj[12], ni[3], nj[3];
    for (int k = 0; k < 12; ++k) { pi[k] = pts[12 * i + k]; pj[k] = pts[12 * j + k]; }
    for (int k = 0; k < 3; ++k) { ni[k] = normals[3 * i + k]; nj[k] = normals[3 * j + k]; }
    double u[3], v[3];
    exact::sub3(pi + 3, pi, u);
    exact::sub3(pi + 9, pi, v);
    const double nu = nrm3p(u), nv = nrm3p(v);
    int npx = (int)rint(nu / nv * sqrt(64.0));
    int npz = (int)rint(nv / nu * sqrt(64.0));
    if (npz == 0) npz = 1;
    if (npx == 0) npx = 1;
    double R[9];
    exact::rotation_matrix(ni, R);
    double e0[3], e1[3], cr[3];
    exact::sub3(pj + 3, pj, e0);
    exact::sub3(pj + 6, pj + 3, e1);
    exact::cross3(e0, e1, cr);
    const double hand = sgn(exact::dot3(cr, nj));
    const double sstep = 1.0 / (npx * 2), sstepz = 1.0 / (npz * 2);
    const int total = npx * npz;
    double acc = 0.0;
    for (int idx = lane; idx < total; idx += 32) {
        const int ix = idx / npz, iz = idx - ix * npz;
        const double stop = 1 - 1.0 / npx, stopz = 1 - 1.0 / npz;
        double s = (npx > 1) ? (ix * (stop / (npx - 1))) : 0.0;
        if (npx > 1 && ix == npx - 1) s = stop;
        s += sstep;
        double t = (npz > 1) ? (iz * (stopz / (npz - 1))) : 0.0;
        if (npz > 1 && iz == npz - 1) t = stopz;
        t += sstepz;
        double o[3];
        for (int k = 0; k < 3; ++k) o[k] = s * u[k] + t * v[k] + pi[k];
        acc += nusselt_analog(o, R, pj, hand);
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) ff[p] = acc * (1 / (SPB_PI * total));
}

// ---------------------------------------------------------------------------
// point <-> patch factors: integration.py:295-344, geometry.py:688-715
// ---------------------------------------------------------------------------
__device__ __forceinline__ void sphere_tangent(const double *v0, const double *v1, double *out) {
    if (fabs(exact::dot3(v0, v1)) > 1e-10) {
        double d[3];
        exact::sub3(v1, v0, d);
        const double q = exact::dot3(d, v0) / exact::dot3(v0, v0);
        for (int k = 0; k < 3; ++k) out[k] = d[k] - q * v0[k];
        const double nn = nrm3p(out);
        for (int k = 0; k < 3; ++k) out[k] /= nn;
    } else {
        const double nn = nrm3p(v1);
        for (int k = 0; k < 3; ++k) out[k] = v1[k] / nn;
    }
}

__device__ double polygon_area4(const double *pts) {
    double area = 0.0;
    for (int t = 0; t < 2; ++t) {
        double a[3], b[3], c[3];
        exact::sub3(pts + 3 * (t + 1), pts, a);
        exact::sub3(pts + 3 * (t + 2), pts, b);
        exact::cross3(a, b, c);
        area += .5 * nrm3p(c);
    }
    return area;
}

__device__ double pt_solution(const double *point, const double *patch, bool receiver_mode) {
    const double source_area = receiver_mode ? polygon_area4(patch) : 4.0;
    double sph[4][3];
    for (int i = 0; i < 4; ++i) {
        double d[3];
        exact::sub3(patch + 3 * i, point, d);
        const double nn = nrm3p(d);
        for (int k = 0; k < 3; ++k) sph[i][k] = d[k] / nn;
    }
    double sum = 0.0;
    for (int i = 0; i < 4; ++i) {
        double v0[3], v1[3];
        sphere_tangent(sph[i], sph[(i + 3) & 3], v0);
        sphere_tangent(sph[i], sph[(i + 1) & 3], v1);
        sum += acos(exact::dot3(v0, v1));
    }
    return (sum - 2 * SPB_PI) / (SPB_PI * source_area);
}

// universal.py:98-147 + RadiosityFast.py:988-1034:
//   distance (x87 norm, feeds the delay bins), E0[j,d,b]
__global__ void __launch_bounds__(128)
k_source_energy(const double *__restrict__ src, const double *__restrict__ centers,
                const double *__restrict__ pts, const uint8_t *__restrict__ vis,
                const double *__restrict__ air, const int64_t *__restrict__ patch_to_wall,
                const double *__restrict__ vi, int64_t n_in, const double *__restrict__ brdf,
                const int64_t *__restrict__ brdf_index, int64_t n_out, int64_t n_bands,
                int64_t n, double *__restrict__ distance, double *__restrict__ e0,
                double *__restrict__ energy) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    // blockIdx.y = source of a batch: all per-source arrays are stacked along the first axis
    const int64_t is = blockIdx.y;
    src += 3 * is;
    vis += is * n;
    distance += is * n;
    e0 += is * n * n_out * n_bands;
    if (energy) energy += is * n * n_bands;
    const double s[3] = {src[0], src[1], src[2]};
    double c[3], patch[12];
    for (int k = 0; k < 3; ++k) c[k] = centers[3 * j + k];
    double dist = 0.0, g = 0.0;
    if (vis[j]) {
        for (int k = 0; k < 12; ++k) patch[k] = pts[12 * j + k];
        double d[3];
        exact::sub3(s, c, d);
        dist = exact::nrm3(d);
        g = pt_solution(s, patch, false);
    }
    distance[j] = dist;
    const int64_t w = patch_to_wall[j];
    const int sidx = exact::nearest_direction(s, c, vi + 3 * n_in * w, (int)n_in);
    const double *row = brdf + ((brdf_index[w] * n_in + sidx) * n_out) * n_bands;
    for (int64_t b = 0; b < n_bands; ++b) {
        const double e = vis[j] ? exp(-air[b] * dist) * g : 0.0;
        if (energy) energy[j * n_bands + b] = e;
        for (int64_t d = 0; d < n_out; ++d)
            e0[(j * n_out + d) * n_bands + b] = e * row[d * n_bands + b];
    }
}

// RadiosityFast.py:711-748 for a batch of receivers:
//   factor (universal.py:149-160), outgoing direction index (:728-730), distance
//   = numpy norm(axis=1) model, ceil delay (:1178-1179), scale = factor*exp(-air d)
__global__ void __launch_bounds__(128)
k_receiver_factors(const double *__restrict__ rcv, int64_t n_rcv,
                   const double *__restrict__ centers, const double *__restrict__ pts,
                   const uint8_t *__restrict__ vis, const double *__restrict__ air,
                   const int64_t *__restrict__ patch_to_wall, const double *__restrict__ vo,
                   int64_t n_out, int64_t n_bands, int64_t n, double speed_of_sound, double dt,
                   int64_t n_samples, double *__restrict__ factor, int32_t *__restrict__ rdir,
                   int32_t *__restrict__ delay, int32_t *__restrict__ shift,
                   double *__restrict__ scale) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rcv * n) return;
    const int64_t r = idx / n, k = idx - r * n;
    const double p[3] = {rcv[3 * r], rcv[3 * r + 1], rcv[3 * r + 2]};
    double c[3], patch[12];
    for (int q = 0; q < 3; ++q) c[q] = centers[3 * k + q];
    double f = 0.0;
    if (vis[idx]) {
        for (int q = 0; q < 12; ++q) patch[q] = pts[12 * k + q];
        f = pt_solution(p, patch, true);
    }
    factor[idx] = f;
    rdir[idx] = exact::nearest_direction(p, c, vo + 3 * n_out * patch_to_wall[k], (int)n_out);
    double d[3];
    exact::sub3(c, p, d);
    const double dist = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    const double bins = ceil(dist / speed_of_sound / dt);
    const int64_t dl = (int64_t)bins;
    delay[idx] = (int32_t)dl;
    shift[idx] = (int32_t)(((dl % n_samples) + n_samples) % n_samples);
    for (int64_t b = 0; b < n_bands; ++b) scale[idx * n_bands + b] = f * exp(-air[b] * dist);
}

// RadiosityFast.py:403-414 (outgoing index of the sender towards the receiver),
// :1386-1390 (incoming index on the SENDER's wall), :538-543 (numpy 1-D norm).
// Directed entry 2p is lo->hi, 2p+1 is hi->lo.
__global__ void __launch_bounds__(128)
k_pair_geometry(const double *__restrict__ centers, const int64_t *__restrict__ patch_to_wall,
                const int32_t *__restrict__ pairs, int64_t n_pairs,
                const double *__restrict__ vi, int64_t n_in, const double *__restrict__ vo,
                int64_t n_out, double *__restrict__ dist, int32_t *__restrict__ out_dir,
                int32_t *__restrict__ in_dir) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= 2 * n_pairs) return;
    const int64_t p = q >> 1;
    const int64_t lo = pairs[2 * p], hi = pairs[2 * p + 1];
    const int64_t i = (q & 1) ? hi : lo, j = (q & 1) ? lo : hi;   // sender, receiver
    double ci[3], cj[3];
    for (int k = 0; k < 3; ++k) { ci[k] = centers[3 * i + k]; cj[k] = centers[3 * j + k]; }
    const int64_t w = patch_to_wall[i];
    out_dir[q] = n_out > 1 ? exact::nearest_direction(cj, ci, vo + 3 * n_out * w, (int)n_out) : 0;
    in_dir[q] = n_in > 1 ? exact::nearest_direction(ci, cj, vi + 3 * n_in * w, (int)n_in) : 0;
    if (!(q & 1)) {
        double d[3];
        exact::sub3(ci, cj, d);
        dist[p] = sqrt(fma(d[2], d[2], fma(d[1], d[1], d[0] * d[0])));
    }
}

// delay bins: int(d / c / dt) (RadiosityFast.py:1067-1068, :1135-1136)
__global__ void k_delay_bins(const double *__restrict__ dist, int64_t count,
                             double speed_of_sound, double dt, int32_t *__restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const double bins = dist[k] / speed_of_sound / dt;
    out[k] = bins >= 2147483647.0 ? 2147483647 : (int32_t)bins;
}

// probes so that the tests can pin the device arithmetic model itself
__global__ void k_probe_norms(const double *__restrict__ v, int64_t count, int dim,
                              double *__restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    out[k] = dim == 3 ? exact::nrm3(v + 3 * k) : exact::nrm2(v + 2 * k);
}
__global__ void k_probe_basic_visibility(const double *__restrict__ A,
                                         const double *__restrict__ B,
                                         const Blocker *__restrict__ blockers, int64_t count,
                                         uint8_t *__restrict__ visible,
                                         uint8_t *__restrict__ in_a,
                                         uint8_t *__restrict__ in_b) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    double a[3], b[3], v[3];
    for (int q = 0; q < 3; ++q) { a[q] = A[3 * k + q]; b[q] = B[3 * k + q]; }
    exact::sub3(b, a, v);
    const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const Blocker blk = blockers[k];
    visible[k] = exact::blocked(a, b, v, sqrt(vv), vv > 1e-6, blk) ? 0 : 1;
    in_a[k] = exact::point_in_polygon(a, blk) ? 1 : 0;
    in_b[k] = exact::point_in_polygon(b, blk) ? 1 : 0;
}

}  // namespace spb

using namespace spb;

static int launch_vis_grouped(unsigned grid, cudaStream_t st, const double *centers, int64_t n,
                              const void *blockers, const void *groups, int64_t n_groups,
                              const int32_t *members, const int32_t *bin_ptr,
                              const int32_t *bin_items, const double *strips,
                              const uint8_t *own_in, const int32_t *own_group, int64_t chunks,
                              int64_t row_lo, uint8_t *vis) {
    k_vis_p2p_grouped<<<grid, kVisThreads, 0, st>>>(
        centers, n, (const Blocker *)blockers, (const exact::Group *)groups, (int32_t)n_groups,
        members, bin_ptr, bin_items, strips, own_in, own_group, chunks, row_lo, vis);
    return check_launch("k_vis_p2p_grouped");
}

extern "C" {

size_t spb_blocker_bytes(int64_t m) { return sizeof(Blocker) * (size_t)(m > 0 ? m : 0); }

int spb_make_blockers(const double *surf_points, const double *surf_normals, int64_t m,
                      int nvert, void *blockers, void *stream) {
    SPB_REQUIRE(nvert == 4, "only quadrilateral surfaces are supported (the reference tessellation produces quads)");
    SPB_REQUIRE(surf_points && surf_normals && blockers, "null pointer");
    if (m == 0) return 0;
    k_make_blockers<<<(unsigned)ceil_div(m, 128), 128, 0, (cudaStream_t)stream>>>(
        surf_points, surf_normals, m, (Blocker *)blockers);
    return check_launch("k_make_blockers");
}

int spb_visibility_p2p(const double *centers, int64_t n, const void *blockers, int64_t m,
                       uint8_t *vis, void *stream) {
    SPB_REQUIRE(centers && vis && (blockers || m == 0), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPB_CUDA(cudaMemsetAsync(vis, 0, (size_t)n * n, st));
    if (n < 2) return 0;
    const int64_t chunks = ceil_div(n, kVisThreads);
    SPB_REQUIRE(n * chunks <= 2147483647LL, "too many patches for one launch");
    k_vis_p2p<<<(unsigned)(n * chunks), kVisThreads, 0, st>>>(
        centers, n, (const Blocker *)blockers, m, chunks, vis);
    return check_launch("k_vis_p2p");
}

int spb_visibility_p2p_grouped(const double *centers, int64_t n, const void *blockers,
                               const void *groups, int64_t n_groups, const int32_t *members,
                               const int32_t *bin_ptr, const int32_t *bin_items,
                               const double *strips, const uint8_t *own_in,
                               const int32_t *own_group, uint8_t *vis, void *stream) {
    SPB_REQUIRE(centers && vis && blockers && groups && members && bin_ptr && bin_items &&
                strips, "null pointer");
    SPB_REQUIRE(n_groups >= 0 && n_groups <= 2147483647LL, "n_groups");
    cudaStream_t st = (cudaStream_t)stream;
    SPB_CUDA(cudaMemsetAsync(vis, 0, (size_t)n * n, st));
    if (n < 2) return 0;
    const int64_t chunks = ceil_div(n, kVisThreads);
    SPB_REQUIRE(n * chunks <= 2147483647LL, "too many patches for one launch");
    return launch_vis_grouped((unsigned)(n * chunks), st, centers, n, blockers, groups, n_groups,
                              members, bin_ptr, bin_items, strips, own_in, own_group, chunks, 0,
                              vis);
}

int spb_visibility_p2p_grouped_rows(const double *centers, int64_t n, const void *blockers,
                                    const void *groups, int64_t n_groups,
                                    const int32_t *members, const int32_t *bin_ptr,
                                    const int32_t *bin_items, const double *strips,
                                    const uint8_t *own_in, const int32_t *own_group,
                                    int64_t row_lo, int64_t row_hi, uint8_t *vis_rows,
                                    void *stream) {
    SPB_REQUIRE(centers && vis_rows && blockers && groups && members && bin_ptr && bin_items &&
                strips, "null pointer");
    SPB_REQUIRE(n_groups >= 0 && n_groups <= 2147483647LL, "n_groups");
    SPB_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= n, "row range");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = row_hi - row_lo;
    if (rows == 0) return 0;
    SPB_CUDA(cudaMemsetAsync(vis_rows, 0, (size_t)rows * n, st));
    const int64_t chunks = ceil_div(n, kVisThreads);
    SPB_REQUIRE(rows * chunks <= 2147483647LL, "too many rows for one launch");
    return launch_vis_grouped((unsigned)(rows * chunks), st, centers, n, blockers, groups,
                              n_groups, members, bin_ptr, bin_items, strips, own_in, own_group,
                              chunks, row_lo, vis_rows);
}

int spb_visibility_own_in(const double *centers, int64_t count, const void *blockers,
                          uint8_t *own_in, void *stream) {
    SPB_REQUIRE(centers && blockers && own_in, "null pointer");
    if (count == 0) return 0;
    k_own_in<<<(unsigned)ceil_div(count, 128), 128, 0, (cudaStream_t)stream>>>(
        centers, count, (const Blocker *)blockers, own_in);
    return check_launch("k_own_in");
}

size_t spb_group_bytes(void) { return sizeof(exact::Group); }

/* Host twins (HOST pointers) of make_blockers and the grouped visibility: the same
 * predicates compiled for the CPU, used by the CPU test-suite to check the grouping
 * logic and its Python table builder against the oracle without a GPU. */
int spb_make_blockers_host(const double *surf_points_h, const double *surf_normals_h,
                           int64_t m, void *blockers_h) {
    SPB_REQUIRE(surf_points_h && surf_normals_h && blockers_h, "null pointer");
    Blocker *out = (Blocker *)blockers_h;
    for (int64_t s = 0; s < m; ++s)
        exact::make_blocker(surf_points_h + 12 * s, surf_normals_h + 3 * s, out[s]);
    return 0;
}

int spb_visibility_p2p_grouped_host(const double *centers_h, int64_t n, const void *blockers_h,
                                    const void *groups_h, int64_t n_groups,
                                    const int32_t *members_h, const int32_t *bin_ptr_h,
                                    const int32_t *bin_items_h, const double *strips_h,
                                    int64_t n_own, const int32_t *own_group_h, uint8_t *vis_h) {
    SPB_REQUIRE(centers_h && vis_h && blockers_h && groups_h && strips_h, "null pointer");
    SPB_REQUIRE(n_own >= 0 && n_own <= n, "n_own");
    // memo of the first n_own centres (centre i <-> blocker i), 255 = not computed
    std::vector<uint8_t> own_in((size_t)n, 255);
    for (int64_t i = 0; i < n_own; ++i)
        own_in[i] = exact::own_in_polygon(centers_h + 3 * i, ((const Blocker *)blockers_h)[i]);
    const Blocker *blockers = (const Blocker *)blockers_h;
    const exact::Group *groups = (const exact::Group *)groups_h;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j) {
            uint8_t out = 0;
            if (j > i) {
                const double *A = centers_h + 3 * i, *B = centers_h + 3 * j;
                double v[3];
                exact::sub3(B, A, v);
                const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
                const exact::Own own = exact::make_own(i, j, n_own ? own_in.data() : nullptr);
                const int64_t g_a = own_group_h ? own_group_h[i] : -1;
                const int64_t g_b = own_group_h ? own_group_h[j] : -1;
                bool visible = true;
                for (int64_t step = -2; step < n_groups && visible; ++step) {
                    const int64_t g = step == -2 ? g_a : (step == -1 ? g_b : step);
                    if (g < 0 || (step >= -1 && g == g_a) || (step >= 0 && g == g_b)) continue;
                    visible = !exact::group_blocked(A, B, v, sqrt(vv), vv > 1e-6, groups[g],
                                                    blockers, members_h, bin_ptr_h, bin_items_h,
                                                    strips_h, own);
                }
                out = visible ? 1 : 0;
            }
            vis_h[i * n + j] = out;
        }
    return 0;
}

int spb_visibility_pt2p(const double *points, int64_t n_points, const double *centers,
                        int64_t n, const void *blockers, int64_t m, uint8_t *vis,
                        void *stream) {
    SPB_REQUIRE(points && centers && vis && (blockers || m == 0), "null pointer");
    if (n_points == 0 || n == 0) return 0;
    const int64_t chunks = ceil_div(n, kVisThreads);
    SPB_REQUIRE(n_points * chunks <= 2147483647LL, "too many points for one launch");
    k_vis_pt2p<<<(unsigned)(n_points * chunks), kVisThreads, 0, (cudaStream_t)stream>>>(
        points, centers, n, (const Blocker *)blockers, m, chunks, vis);
    return check_launch("k_vis_pt2p");
}

int spb_form_factors_stokes(const double *pts, const double *areas, const int32_t *pairs,
                            int64_t n_pairs, double *ff, uint8_t *nusselt_flag, void *stream) {
    SPB_REQUIRE(pts && areas && ff && nusselt_flag && (pairs || n_pairs == 0), "null pointer");
    if (n_pairs == 0) return 0;
    SPB_REQUIRE(n_pairs <= (2147483647LL * 8), "too many pairs for one launch");
    k_ff_stokes<<<(unsigned)ceil_div(n_pairs * 16, 128), 128, 0, (cudaStream_t)stream>>>(
        pts, areas, pairs, n_pairs, ff, nusselt_flag);
    return check_launch("k_ff_stokes");
}

int spb_form_factors_nusselt(const double *pts, const double *normals, const int32_t *pairs,
                             const int64_t *todo, int64_t n_todo, double *ff, void *stream) {
    SPB_REQUIRE(pts && normals && ff && (todo || n_todo == 0), "null pointer");
    if (n_todo == 0) return 0;
    k_ff_nusselt<<<(unsigned)ceil_div(n_todo * 32, 128), 128, 0, (cudaStream_t)stream>>>(
        pts, normals, pairs, todo, n_todo, ff);
    return check_launch("k_ff_nusselt");
}

int spb_source_energy(const double *src, const double *centers, const double *pts,
                      const uint8_t *vis, const double *air, const int64_t *patch_to_wall,
                      const double *vi, int64_t n_in, const double *brdf,
                      const int64_t *brdf_index, int64_t n_out, int64_t n_bands, int64_t n,
                      double *distance, double *e0, double *energy, void *stream) {
    SPB_REQUIRE(src && centers && pts && vis && air && patch_to_wall && vi && brdf &&
                brdf_index && distance && e0, "null pointer");
    if (n == 0) return 0;
    k_source_energy<<<(unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(
        src, centers, pts, vis, air, patch_to_wall, vi, n_in, brdf, brdf_index, n_out, n_bands,
        n, distance, e0, energy);
    return check_launch("k_source_energy");
}

int spb_source_energy_batch(const double *src, int64_t n_src, const double *centers,
                            const double *pts, const uint8_t *vis, const double *air,
                            const int64_t *patch_to_wall, const double *vi, int64_t n_in,
                            const double *brdf, const int64_t *brdf_index, int64_t n_out,
                            int64_t n_bands, int64_t n, double *distance, double *e0,
                            double *energy, void *stream) {
    SPB_REQUIRE(src && centers && pts && vis && air && patch_to_wall && vi && brdf &&
                brdf_index && distance && e0, "null pointer");
    SPB_REQUIRE(n_src >= 0 && n_src <= 65535, "at most 65535 sources per call");
    if (n == 0 || n_src == 0) return 0;
    dim3 grid((unsigned)ceil_div(n, 128), (unsigned)n_src);
    k_source_energy<<<grid, 128, 0, (cudaStream_t)stream>>>(
        src, centers, pts, vis, air, patch_to_wall, vi, n_in, brdf, brdf_index, n_out, n_bands,
        n, distance, e0, energy);
    return check_launch("k_source_energy");
}

int spb_receiver_factors(const double *rcv, int64_t n_rcv, const double *centers,
                         const double *pts, const uint8_t *vis, const double *air,
                         const int64_t *patch_to_wall, const double *vo, int64_t n_out,
                         int64_t n_bands, int64_t n, double speed_of_sound, double dt,
                         int64_t n_samples, double *factor, int32_t *rdir, int32_t *delay,
                         int32_t *shift, double *scale, void *stream) {
    SPB_REQUIRE(rcv && centers && pts && vis && air && patch_to_wall && vo && factor && rdir &&
                delay && shift && scale, "null pointer");
    SPB_REQUIRE(n_samples > 0 && speed_of_sound > 0 && dt > 0, "n_samples, c, dt > 0");
    if (n_rcv * n == 0) return 0;
    k_receiver_factors<<<(unsigned)ceil_div(n_rcv * n, 128), 128, 0, (cudaStream_t)stream>>>(
        rcv, n_rcv, centers, pts, vis, air, patch_to_wall, vo, n_out, n_bands, n,
        speed_of_sound, dt, n_samples, factor, rdir, delay, shift, scale);
    return check_launch("k_receiver_factors");
}

int spb_pair_geometry(const double *centers, const int64_t *patch_to_wall,
                      const int32_t *pairs, int64_t n_pairs, const double *vi, int64_t n_in,
                      const double *vo, int64_t n_out, double *dist, int32_t *out_dir,
                      int32_t *in_dir, void *stream) {
    SPB_REQUIRE(centers && patch_to_wall && dist && out_dir && in_dir, "null pointer");
    SPB_REQUIRE((n_in <= 1 || vi) && (n_out <= 1 || vo), "direction arrays");
    if (n_pairs == 0) return 0;
    k_pair_geometry<<<(unsigned)ceil_div(2 * n_pairs, 128), 128, 0, (cudaStream_t)stream>>>(
        centers, patch_to_wall, pairs, n_pairs, vi, n_in, vo, n_out, dist, out_dir, in_dir);
    return check_launch("k_pair_geometry");
}

int spb_delay_bins(const double *dist, int64_t count, double speed_of_sound, double dt,
                   int32_t *out, void *stream) {
    SPB_REQUIRE((dist && out) || count == 0, "null pointer");
    SPB_REQUIRE(speed_of_sound > 0 && dt > 0, "c, dt > 0");
    if (count == 0) return 0;
    k_delay_bins<<<(unsigned)ceil_div(count, 256), 256, 0, (cudaStream_t)stream>>>(
        dist, count, speed_of_sound, dt, out);
    return check_launch("k_delay_bins");
}

int spb_probe_norms(const double *v, int64_t count, int dim, double *out, void *stream) {
    SPB_REQUIRE(dim == 2 || dim == 3, "dim");
    if (count == 0) return 0;
    k_probe_norms<<<(unsigned)ceil_div(count, 128), 128, 0, (cudaStream_t)stream>>>(v, count, dim, out);
    return check_launch("k_probe_norms");
}

int spb_probe_basic_visibility(const double *a, const double *b, const void *blockers,
                               int64_t count, uint8_t *visible, uint8_t *in_a, uint8_t *in_b,
                               void *stream) {
    if (count == 0) return 0;
    k_probe_basic_visibility<<<(unsigned)ceil_div(count, 64), 64, 0, (cudaStream_t)stream>>>(
        a, b, (const Blocker *)blockers, count, visible, in_a, in_b);
    return check_launch("k_probe_basic_visibility");
}

}  // extern "C"
