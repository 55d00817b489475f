// Hierarchical (per-wall) visibility: the conjunction over all blocking patches
//     vis(A, B) = AND_k  not blocked(A, B, patch_k)           (geometry.py:786-795)
// evaluated group by group, where a group is a set of coplanar blockers with the
// bitwise-same normal (the patches of one wall).  For a group the plane quantities
// of `_basic_visibility` are (up to the tiny, measured deviation `plane_dev` of the
// members' first vertices from the common plane) the same for every member, so one
// evaluation decides the whole group in the common cases:
//
//   * both end points clearly off the plane: no plane hit, or a hit clearly outside
//     the open segment -> no member blocks (rule (a) of exact::blocked);
//   * otherwise the only members that can block are those whose polygon the plane hit
//     (or the in-plane end point) can be "in": every other member has
//     exact::ray_clearance > the distance the member's own hit can differ from the
//     group's.  The members are listed in a 2-D grid of cells over the in-plane (x, y)
//     coordinates, so those candidates are found by scanning ONE cell (a handful of
//     members); each candidate is then evaluated with the exact per-blocker predicate
//     exact::blocked.  The cell lists hold every member whose expanded bounding box
//     touches the cell.  That misses the one kind of candidate that lies far from its
//     polygon: a point LEFT of an axis-aligned rectangle within ~1e-3 of the height of
//     one of its horizontal edges (its +x ray grazes the edge, see ray_clearance).  The
//     y-ranges where that can happen ("strips", merged per group) are tabulated; a query
//     point inside a strip scans the whole 1-D bin along y instead (all members of the
//     band), as does every query of a group without cells.
//
// Whenever a bound needed for these arguments does not hold (end point within the
// deviation of the eta threshold, grazing segment, huge segment) the group is simply
// evaluated member by member.  The result is therefore identical to the brute-force
// conjunction; tests/native/exact_selftest.cpp and tests/test_visgroup_cpu.py check
// that on the host, the GPU tests on the device.
#pragma once
#include "exact.cuh"

namespace spb {
namespace exact {

struct Group {
    double n[3];         // common plane normal (bitwise equal for all members)
    double s0[3];        // first vertex of the first member
    double r0[3];        // rotation rows (functions of n only)
    double r1[3];
    double plane_dev;    // max_k |DOT(S0_k - s0, n)| over the members (0 for lattices)
    double y0;           // lower edge of bin 0 (in-plane y)
    double inv_bin_h;    // 1 / bin height
    double x0;           // left edge of cell column 0 (in-plane x)
    double inv_bin_w;    // 1 / cell width
    int32_t n_bins;
    int32_t bin_ptr0;    // this group's bins are bin_ptr[bin_ptr0 .. bin_ptr0 + n_bins]
    int32_t m0, m1;      // members[m0 .. m1) = blocker indices of the group
    int32_t n_bx;        // cell columns (0: no cells, always scan the 1-D bin)
    int32_t cell_ptr0;   // cell (bx, by) is bin_ptr[cell_ptr0 + by * n_bx + bx .. + 1]
    int32_t strip0;      // strips[2 * (strip0 + s) + {0, 1}] = [lo, hi] of strip s, sorted
    int32_t n_strips;
    int32_t sfirst0;     // bin_ptr[sfirst0 + bin] = first strip with hi >= lower edge of bin
    int32_t pad_;
};

constexpr double kGroupMargin = 1e-3;    // how far a member's hit may differ from the group's

// Memo of the one exact polygon test almost every pair needs twice: an end point that is the
// centre of a patch lies in that patch's own plane and polygon, so the member with the end
// point's own index always reaches point_in_polygon_sides (x87-emulated norms, ~1000
// instructions).  own_in[i] = point_in_polygon_sides(centre i in the frame of blocker i) is
// computed once per patch (own_in_polygon below) and handed to exact::blocked as a hint when
// the member's index equals the end point's index.  255 = not computed.
struct Own {
    int32_t a, b;        // index of end point A / B (= index of "its" blocker), -1: none
    int a_in, b_in;      // memoised point_in_polygon_sides, -1: not known
};
SPB_FN Own no_own() { return Own{-1, -1, -1, -1}; }
SPB_FN Own make_own(int64_t i, int64_t j, const uint8_t *own_in) {
    if (!own_in) return no_own();
    const int ai = own_in[i], bi = own_in[j];
    return Own{(int32_t)i, (int32_t)j, ai <= 1 ? ai : -1, bi <= 1 ? bi : -1};
}
SPB_FN bool own_in_polygon(const double *c, const Blocker &k) {
    return point_in_polygon_sides(dot3(k.r0, c), dot3(k.r1, c), &k);   // as in exact::blocked
}
SPB_FN bool blocked_member(const double *A, const double *B, const double *v, double vlen,
                           bool cull_ok, const Blocker *blockers, int32_t m, const Own &own) {
    return blocked(A, B, v, vlen, cull_ok, blockers[m], m == own.a ? own.a_in : -1,
                   m == own.b ? own.b_in : -1);
}

// every member of the group, one by one (always correct)
SPB_FN bool group_blocked_bruteforce(const double *A, const double *B, const double *v,
                                     double vlen, bool cull_ok, const Group &g,
                                     const Blocker *blockers, const int32_t *members,
                                     const Own &own) {
    for (int32_t q = g.m0; q < g.m1; ++q)
        if (blocked_member(A, B, v, vlen, cull_ok, blockers, members[q], own)) return true;
    return false;
}

// is qy inside one of the group's strips?  (sorted, disjoint; the scan starts at the first
// strip that reaches into the query's bin, so it looks at 1-3 strips for a lattice)
SPB_FN bool in_strip(const Group &g, int32_t bin, double qy, const int32_t *bin_ptr,
                     const double *strips) {
    for (int32_t s = bin_ptr[g.sfirst0 + bin]; s < g.n_strips; ++s) {
        const double lo = strips[2 * (g.strip0 + s)];
        if (lo > qy) return false;
        if (qy <= strips[2 * (g.strip0 + s) + 1]) return true;
    }
    return false;
}

// members whose polygon a point within kGroupMargin of (qx, qy) could be "in"
SPB_FN bool group_blocked_near(const double *A, const double *B, const double *v, double vlen,
                               bool cull_ok, const Group &g, double qx, double qy,
                               const Blocker *blockers, const int32_t *bin_ptr,
                               const int32_t *bin_items, const double *strips, const Own &own) {
    const double fb = (qy - g.y0) * g.inv_bin_h;
    if (!(fb > -1.0) || !(fb < (double)g.n_bins + 1.0)) return false;   // outside every band
    int32_t bin = (int32_t)floor(fb);
    bin = bin < 0 ? 0 : (bin >= g.n_bins ? g.n_bins - 1 : bin);
    int32_t list = g.bin_ptr0 + bin;                    // 1-D bin: every member of the band
    if (g.n_bx > 0 && !in_strip(g, bin, qy, bin_ptr, strips)) {
        const double fx = (qx - g.x0) * g.inv_bin_w;    // NaN -> column 0 (nothing passes)
        int32_t bx = fx >= 1.0 ? (fx < (double)g.n_bx ? (int32_t)floor(fx) : g.n_bx - 1) : 0;
        list = g.cell_ptr0 + bin * g.n_bx + bx;
    }
    const int32_t p0 = bin_ptr[list], p1 = bin_ptr[list + 1];
    for (int32_t p = p0; p < p1; ++p) {
        const int32_t m = bin_items[p];
        if (ray_clearance(qx, qy, blockers[m]) > kGroupMargin + kClearGuard) continue;
        if (blocked_member(A, B, v, vlen, cull_ok, blockers, m, own)) return true;
    }
    return false;
}

// does any member of the group block the segment A-B?   v = B - A, vlen = |v|
SPB_FN bool group_blocked(const double *A, const double *B, const double *v, double vlen,
                          bool cull_ok, const Group &g, const Blocker *blockers,
                          const int32_t *members, const int32_t *bin_ptr,
                          const int32_t *bin_items, const double *strips, const Own &own) {
    double wa[3], w[3];
    sub3(A, g.s0, wa);
    sub3(B, g.s0, w);
    const double dA = dot3(wa, g.n);
    const double dB = dot3(w, g.n);
    const double dp = dot3(v, g.n);            // identical for every member
    const double dev = g.plane_dev + 1e-11;    // |dE_k - dE_group| <= dev for E in {A, B}
    const bool offA = fabs(dA) > kEta + dev, offB = fabs(dB) > kEta + dev;
    const bool inplA = fabs(dA) < kEta - dev, inplB = fabs(dB) < kEta - dev;
    if (!(offA || inplA) || !(offB || inplB) || !cull_ok || !(vlen < 1e4))
        return group_blocked_bruteforce(A, B, v, vlen, cull_ok, g, blockers, members, own);

    if (offA && offB) {
        // neither end point lies in any member's plane
        if (!(fabs(dp) > 1e-6)) return false;                       // no plane hit at all
        const double u = -dB;
        const bool outside = dp > 0 ? (u > 3e-3 * dp || u < -1.003 * dp)
                                    : (u < 3e-3 * dp || u > -1.003 * dp);
        // member hit parameters differ from the group's by <= dev/|dp| <= ~1e-6
        if (outside) return false;
        const double shift = dev * vlen / fabs(dp) + 1e-9;          // |hit_k - hit_group|
        if (!(shift < kGroupMargin))
            return group_blocked_bruteforce(A, B, v, vlen, cull_ok, g, blockers, members, own);
        const double fac = -(dB / dp);
        const double pt[3] = {(w[0] + g.s0[0]) + fac * v[0], (w[1] + g.s0[1]) + fac * v[1],
                              (w[2] + g.s0[2]) + fac * v[2]};
        return group_blocked_near(A, B, v, vlen, cull_ok, g, dot3(g.r0, pt), dot3(g.r1, pt),
                                  blockers, bin_ptr, bin_items, strips, own);
    }
    if (inplA && inplB) {
        // both end points in the plane: a member blocks only if an end point is in its
        // polygon -- unless there is a "plane hit" with |dp| > 1e-6 (dp ~ dB - dA <= 2e-6)
        if (fabs(dp) > 1e-6)
            return group_blocked_bruteforce(A, B, v, vlen, cull_ok, g, blockers, members, own);
        if (group_blocked_near(A, B, v, vlen, cull_ok, g, dot3(g.r0, A), dot3(g.r1, A), blockers,
                               bin_ptr, bin_items, strips, own))
            return true;
        return group_blocked_near(A, B, v, vlen, cull_ok, g, dot3(g.r0, B), dot3(g.r1, B),
                                  blockers, bin_ptr, bin_items, strips, own);
    }
    // exactly one end point E in the plane, the other clearly off it: a member blocks
    // only if E is in its polygon, or the plane hit -- within eta*|v|/|dp| of E -- is
    // (rule (b) of exact::blocked)
    if (!(fabs(dp) > 1e-6)) {
        // no plane hit: only "E in the polygon" can block
        const double *E = inplA ? A : B;
        return group_blocked_near(A, B, v, vlen, cull_ok, g, dot3(g.r0, E), dot3(g.r1, E),
                                  blockers, bin_ptr, bin_items, strips, own);
    }
    const double slack = (kEta + dev) * vlen / fabs(dp) + 3e-9;
    if (!(slack < kGroupMargin))
        return group_blocked_bruteforce(A, B, v, vlen, cull_ok, g, blockers, members, own);
    const double *E = inplA ? A : B;
    return group_blocked_near(A, B, v, vlen, cull_ok, g, dot3(g.r0, E), dot3(g.r1, E), blockers,
                              bin_ptr, bin_items, strips, own);
}

}  // namespace exact
}  // namespace spb
