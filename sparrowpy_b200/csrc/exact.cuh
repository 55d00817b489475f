// Device restatement of the reference's geometric predicates with its exact
// rounding model:
//   DOT = FMA chain (numba np.dot -> BLAS ddot), NRM = x87 80-bit norm (x87.cuh),
//   every other operation a separately rounded IEEE double op in source order.
// This header must be compiled with --fmad=false.
#pragma once
#include "x87.cuh"

// The predicates are plain C++: under nvcc they are device (and host) functions,
// under g++ they compile for the host, where tests/native/exact_selftest.cpp checks
// them against the CPU oracle on millions of cases.
#if defined(__CUDACC__)
#define SPB_FN __host__ __device__ __forceinline__
#define SPB_FN_NOINLINE __host__ __device__ __noinline__
#else
#define SPB_FN inline
#define SPB_FN_NOINLINE inline
#endif

namespace spb {
namespace exact {

constexpr double kEta = 1e-6;

SPB_FN double dot3(const double *a, const double *b) {
    double s = a[0] * b[0];
    s = fma(a[1], b[1], s);
    s = fma(a[2], b[2], s);
    return s;
}
SPB_FN double dot2(const double *a, const double *b) {
    double s = a[0] * b[0];
    s = fma(a[1], b[1], s);
    return s;
}
SPB_FN double nrm3(const double *v) { return x87::norm3(v[0], v[1], v[2]); }
SPB_FN double nrm2(const double *v) { return x87::norm2(v[0], v[1]); }
SPB_FN void sub3(const double *a, const double *b, double *o) {
    o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2];
}
SPB_FN void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// reference geometry.py:498-559 with the default target (0,0,1)
SPB_FN void rotation_matrix(const double *n, double R[9]) {
    if (n[0] == 0.0 && n[1] == 0.0 && n[2] == 1.0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    const double nn = nrm3(n);
    const double a[3] = {n[0] / nn, n[1] / nn, n[2] / nn};
    const double b[3] = {0.0, 0.0, 1.0};
    const double c = dot3(a, b);
    if (c != -1) {
        double v[3];
        cross3(a, b, v);
        const double s = nrm3(v);
        const double K[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        const double f = (1 - c) / (s * s);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double acc = K[3 * i] * K[j];               // kmat.dot(kmat)
                acc = fma(K[3 * i + 1], K[3 + j], acc);
                acc = fma(K[3 * i + 2], K[6 + j], acc);
                const double eye = (i == j) ? 1.0 : 0.0;
                R[3 * i + j] = (eye + K[3 * i + j]) + acc * f;
            }
    } else {
        R[0] = -1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = -1;
    }
}

// Everything about a blocking quadrilateral that `_point_in_polygon`
// (geometry.py:614-686) recomputes on every call but that depends on the polygon
// only.  Computed once per blocker with the same operations, so reusing it
// cannot change a result.
struct Blocker {
    double n[3];        // plane normal
    double s0[3];       // first vertex
    double r0[3];       // rows 0 and 1 of the rotation matrix
    double r1[3];
    double p2[4][2];    // vertices rotated into the plane
    double nl[4][2];    // side normals (-side_y, side_x) / NRM(side)
    double len[4];      // NRM(a1 - a0)
    double xmin, xmax, ymin, ymax;   // bounding box of p2
    double h;           // cull margin (see ray_clearance)
    double aa2d;        // 1.0 when p2 is an exactly axis-aligned rectangle
};
constexpr int kBlockerDoubles = sizeof(Blocker) / sizeof(double);

SPB_FN void make_blocker(const double *pts /*4x3*/, const double *n, Blocker &k) {
    double R[9];
    rotation_matrix(n, R);
    for (int c = 0; c < 3; ++c) {
        k.n[c] = n[c]; k.s0[c] = pts[c]; k.r0[c] = R[c]; k.r1[c] = R[3 + c];
    }
    for (int i = 0; i < 4; ++i) {
        k.p2[i][0] = dot3(R, pts + 3 * i);
        k.p2[i][1] = dot3(R + 3, pts + 3 * i);
    }
    double lmax = 0.0;
    k.xmin = k.p2[0][0]; k.xmax = k.p2[0][0]; k.ymin = k.p2[0][1]; k.ymax = k.p2[0][1];
    bool aa = true;
    for (int i = 0; i < 4; ++i) {
        const double *a1 = k.p2[(i + 1) % 4], *a0 = k.p2[i];
        const double side[2] = {a1[0] - a0[0], a1[1] - a0[1]};
        const double ns = nrm2(side);
        k.nl[i][0] = -side[1] / ns;
        k.nl[i][1] = side[0] / ns;
        k.len[i] = ns;
        lmax = fmax(lmax, ns);
        // sides alternate exactly horizontal / exactly vertical, none degenerate
        const bool horiz = side[1] == 0.0 && side[0] != 0.0;
        const bool vert = side[0] == 0.0 && side[1] != 0.0;
        aa = aa && ((i & 1) == 0 ? (horiz || vert) : true) && (horiz || vert);
        k.xmin = fmin(k.xmin, a0[0]);
        k.xmax = fmax(k.xmax, a0[0]);
        k.ymin = fmin(k.ymin, a0[1]);
        k.ymax = fmax(k.ymax, a0[1]);
    }
    // ... and form a proper rectangle: horizontal and vertical sides alternate, the two
    // vertical sides span exactly [ymin, ymax], sit at xmin and xmax, and run in
    // opposite directions
    int n_vert = 0;
    double dir_sum = 0.0;
    for (int i = 0; i < 4 && aa; ++i) {
        const double *a1 = k.p2[(i + 1) % 4], *a0 = k.p2[i], *a2 = k.p2[(i + 2) % 4];
        const bool h0 = (a1[1] - a0[1]) == 0.0, h1 = (a2[1] - a1[1]) == 0.0;
        if (h0 == h1) aa = false;
        if (!h0) {
            ++n_vert;
            dir_sum += (a1[1] > a0[1]) ? 1.0 : -1.0;
            aa = aa && fmin(a0[1], a1[1]) == k.ymin && fmax(a0[1], a1[1]) == k.ymax &&
                 (a0[0] == k.xmin || a0[0] == k.xmax);
        }
    }
    aa = aa && n_vert == 2 && dir_sum == 0.0 && k.xmin < k.xmax && k.ymin < k.ymax;
    k.aa2d = aa ? 1.0 : 0.0;
    // A side only counts when the ray hit b satisfies
    // |b-a0| + |b-a1| - |a1-a0| <= 1e-6, i.e. b lies within sqrt(1e-6*L/2) of the
    // side.  h = sqrt(5e-6*L) puts 2h^2/L >= 1e-5 = 10 x the tolerance between the
    // culled region and that band (the rounding error of the test is ~1e-15).
    k.h = sqrt(5e-6 * fmax(lmax, 1e-3)) + 1e-9;
}

// geometry.py:658-686: winding count over the four sides for a query point that is
// already rotated into the plane.  Out of line: it carries the x87 emulation and is
// reached for a tiny fraction of the (pair, blocker) combinations only.
SPB_FN_NOINLINE bool point_in_polygon_sides(double ptx, double pty, const Blocker *kp) {
    const Blocker &k = *kp;
    const double pt[2] = {ptx, pty};
    int count = 0;
    const double pt1[2] = {pt[0] + 1., pt[1] + 0.};
    const double v[2] = {pt1[0] - pt[0], pt1[1] - pt[1]};
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const double *a1 = k.p2[(i + 1) % 4], *a0 = k.p2[i];
        const double *nl = k.nl[i];
        const double dp = dot2(v, nl);                      // geometry.py:596-602
        if (!(fabs(dp) > 1e-6)) continue;
        const double w[2] = {pt1[0] - a1[0], pt1[1] - a1[1]};
        const double fac = -(dot2(nl, w) / dp);
        const double b[2] = {(w[0] + a1[0]) + fac * v[0], (w[1] + a1[1]) + fac * v[1]};
        if (b[0] > pt[0]) {
            const double ba0[2] = {b[0] - a0[0], b[1] - a0[1]};
            const double ba1[2] = {b[0] - a1[0], b[1] - a1[1]};
            if (fabs(nrm2(ba0) + nrm2(ba1) - k.len[i]) <= kEta) {
                const double bp[2] = {b[0] - pt[0], b[1] - pt[1]};
                const double d = dot2(bp, nl);
                if (d > 0) count += 1;
                else if (d < 0) count -= 1;
            }
        }
    }
    return count != 0;
}

// How far a point q in the polygon's plane (2-D coordinates) is from the region where
// the side loop of _point_in_polygon could return True.  For any s >= 0:
//     ray_clearance(q) > s   ==>   point_in_polygon_sides(q') is False for every q'
//                                  within distance s of q.
// Two rules, both consequences of how the loop works (geometry.py:658-686):
//  (ray)  the hit b of the +x ray keeps the y of the query point and has b_x > q_x,
//         and a side only counts when b lies within sqrt(1e-6*L/2) << h of it: a
//         point above/below the polygon's y-range or right of its x-range by more
//         than h registers no side at all -> count == 0.
//  (aa)   exactly axis-aligned rectangle, query point left of it and strictly inside
//         its y-range: the horizontal sides give dp = v.nl = (+-0)*1 + 0*(+-1) = 0 (no
//         hit); both vertical sides are hit in their interior, where
//         |b-a0| + |b-a1| - |a1-a0| is pure rounding (~1e-15 << 1e-6), with
//         d = (b_x - q_x) * nl_x of opposite signs (+1 and -1) -> count == 0.
SPB_FN double ray_clearance(double qx, double qy, const Blocker &k) {
    const double c_ray = fmax(fmax(qy - k.ymax, k.ymin - qy), qx - k.xmax) - k.h;
    const double c_aa = fmin(fmin(qy - k.ymin, k.ymax - qy), k.xmin - qx) - 1e-7;
    return k.aa2d != 0.0 ? fmax(c_ray, c_aa) : c_ray;
}
constexpr double kClearGuard = 2e-9;   // absorbs the rounding of ray_clearance itself

// geometry.py:645-686 for a point that passed the coplanarity test (:641)
SPB_FN bool point_in_polygon_2d(const double *p, const Blocker &k) {
    const double ptx = dot3(k.r0, p), pty = dot3(k.r1, p);
    if (ray_clearance(ptx, pty, k) > kClearGuard) return false;
    return point_in_polygon_sides(ptx, pty, &k);
}

// geometry.py:614-686
SPB_FN bool point_in_polygon(const double *p, const Blocker &k) {
    double d0[3];
    sub3(p, k.s0, d0);
    if (fabs(dot3(d0, k.n)) > kEta) return false;
    return point_in_polygon_2d(p, k);
}

// geometry.py:841-909 evaluated for one blocking surface, organised so that the
// common cases cost ~25 FP64 operations.  Only the ORDER of evaluation differs from
// the reference; the predicates are pure, so the value cannot:
//  * inA / inB start with the coplanarity test (geometry.py:641), whose dot products
//    dA, dB are needed anyway (dB is also the numerator of the plane hit parameter);
//  * first branch (neither end point in the surface, :881-893): the reference requires
//    hit exists AND hit in polygon AND (hit-A).(hit-B) < 0, the hit being B + fac*v
//    (+ rounding ~1e-13), fac = -(dB/dp).
//    (a) (hit-A).(hit-B) = fac(1+fac)|v|^2 is positive whenever fac lies outside
//        [-1, 0] by a margin; tested on products (no division) with thresholds
//        2e-3 / -1.002, which imply fac > 1e-3 resp. fac < -1.001 for the rounded
//        quotient as well.
//    (b) an end point E in {A, B} that lies in the surface's plane (|dE| <= eta) but
//        not in the polygon: in exact arithmetic dA = dB - dp, so the hit is
//        E -+ (dE/dp) v, at most eta*|v|/|dp| away from E.  If ray_clearance(E)
//        exceeds that distance (+3e-9 for rounding), "hit in polygon" is False.
//    v = B - A, vlen = |v| (any rounding), cull_ok = |v|^2 > 1e-6.
//    hint_a / hint_b: value of point_in_polygon_sides for the in-plane coordinates of A / B
//    with THIS blocker when it was computed before (0 / 1; -1 = not known) -- the function is
//    pure, so a memoised value is the value; see own_in in vis_group.cuh.
SPB_FN bool blocked(const double *A, const double *B, const double *v,
                                        double vlen, bool cull_ok, const Blocker &k,
                                        int hint_a = -1, int hint_b = -1) {
    double wa[3], w[3];
    sub3(A, k.s0, wa);
    sub3(B, k.s0, w);
    const double dA = dot3(wa, k.n);
    const double dB = dot3(w, k.n);
    const bool copA = !(fabs(dA) > kEta), copB = !(fabs(dB) > kEta);
    bool inA = false, inB = false;
    bool hit_misses = false;                    // rule (b)
    const double dp = dot3(v, k.n);
    if (copA | copB) {
        const double slack = kEta * vlen / fabs(dp) + 3e-9;     // inf/nan when dp == 0: no cull
        if (copA) {
            const double ax = dot3(k.r0, A), ay = dot3(k.r1, A);
            const double clr = ray_clearance(ax, ay, k);
            if (!(clr > kClearGuard))
                inA = hint_a >= 0 ? hint_a != 0 : point_in_polygon_sides(ax, ay, &k);
            else if (!copB) hit_misses = clr > slack;
        }
        if (copB) {
            const double bx = dot3(k.r0, B), by = dot3(k.r1, B);
            const double clr = ray_clearance(bx, by, k);
            if (!(clr > kClearGuard))
                inB = hint_b >= 0 ? hint_b != 0 : point_in_polygon_sides(bx, by, &k);
            else if (!copA) hit_misses = clr > slack;
        }
    }
    if (!inA && !inB) {
        if (!(fabs(dp) > 1e-6)) return false;               // geometry.py:599-604
        if (cull_ok) {
            if (hit_misses) return false;
            const double u = -dB;
            const bool outside = dp > 0 ? (u > 2e-3 * dp || u < -1.002 * dp)
                                        : (u < 2e-3 * dp || u > -1.002 * dp);
            if (outside) return false;
        }
        const double fac = -(dB / dp);
        const double pt[3] = {(w[0] + k.s0[0]) + fac * v[0], (w[1] + k.s0[1]) + fac * v[1],
                              (w[2] + k.s0[2]) + fac * v[2]};
        double pa[3], pb[3];
        sub3(pt, A, pa);
        sub3(pt, B, pb);
        if (!(dot3(pa, pb) < 0)) return false;
        return point_in_polygon(pt, k);
    }
    double d[3];
    if (inA && !inB && (sub3(B, A, d), dot3(k.n, d) < 0)) return true;      // :895-897
    if (!inA && inB && (sub3(A, B, d), dot3(k.n, d) < 0)) return true;      // :899-901
    return fabs(dA) < kEta && fabs(dB) < kEta;                              // :903-906
}

// numba `diff /= np.linalg.norm(diff)` followed by argmin of squared distances
// (RadiosityFast.py:1304-1310, :1386-1389): first minimum wins.
SPB_FN int nearest_direction(const double *to, const double *from,
                                        const double *dirs, int n_dirs) {
    double diff[3];
    sub3(to, from, diff);
    const double nn = nrm3(diff);
    diff[0] /= nn; diff[1] /= nn; diff[2] /= nn;
    int best = 0;
    double bestv = 0.0;
    for (int k = 0; k < n_dirs; ++k) {
        const double e0 = dirs[3 * k] - diff[0], e1 = dirs[3 * k + 1] - diff[1],
                     e2 = dirs[3 * k + 2] - diff[2];
        const double val = (e0 * e0 + e1 * e1) + e2 * e2;
        if (k == 0 || val < bestv) { bestv = val; best = k; }
    }
    return best;
}

}  // namespace exact
}  // namespace spb
