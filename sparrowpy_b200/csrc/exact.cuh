// Device restatement of the reference's geometric predicates with its exact
// rounding model:
//   DOT = FMA chain (numba np.dot -> BLAS ddot), NRM = x87 80-bit norm (x87.cuh),
//   every other operation a separately rounded IEEE double op in source order.
// This header must be compiled with --fmad=false.
#pragma once
#include "x87.cuh"

namespace spb {
namespace exact {

constexpr double kEta = 1e-6;

__device__ __forceinline__ double dot3(const double *a, const double *b) {
    double s = a[0] * b[0];
    s = fma(a[1], b[1], s);
    s = fma(a[2], b[2], s);
    return s;
}
__device__ __forceinline__ double dot2(const double *a, const double *b) {
    double s = a[0] * b[0];
    s = fma(a[1], b[1], s);
    return s;
}
__device__ __forceinline__ double nrm3(const double *v) { return x87::norm3(v[0], v[1], v[2]); }
__device__ __forceinline__ double nrm2(const double *v) { return x87::norm2(v[0], v[1]); }
__device__ __forceinline__ void sub3(const double *a, const double *b, double *o) {
    o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2];
}
__device__ __forceinline__ void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// reference geometry.py:498-559 with the default target (0,0,1)
__device__ inline void rotation_matrix(const double *n, double R[9]) {
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
    double xmax, ymin, ymax;   // bounding box of p2
    double h;           // cull margin (see point_in_polygon)
};
constexpr int kBlockerDoubles = sizeof(Blocker) / sizeof(double);

__device__ inline void make_blocker(const double *pts /*4x3*/, const double *n, Blocker &k) {
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
    k.xmax = k.p2[0][0]; k.ymin = k.p2[0][1]; k.ymax = k.p2[0][1];
    for (int i = 0; i < 4; ++i) {
        const double *a1 = k.p2[(i + 1) % 4], *a0 = k.p2[i];
        const double side[2] = {a1[0] - a0[0], a1[1] - a0[1]};
        const double ns = nrm2(side);
        k.nl[i][0] = -side[1] / ns;
        k.nl[i][1] = side[0] / ns;
        k.len[i] = ns;
        lmax = fmax(lmax, ns);
        k.xmax = fmax(k.xmax, a0[0]);
        k.ymin = fmin(k.ymin, a0[1]);
        k.ymax = fmax(k.ymax, a0[1]);
    }
    // A side only counts when the ray hit b satisfies
    // |b-a0| + |b-a1| - |a1-a0| <= 1e-6, i.e. b lies within sqrt(1e-6*L/2) of the
    // side.  h = 0.01*max(1, L) puts 2h^2/L >= 2e-4 >> 1e-6 between the culled
    // region and that band, far above any rounding error.
    k.h = 0.01 * fmax(1.0, lmax);
}

// geometry.py:614-686.  `culled` results are provably identical to the full
// evaluation: b keeps the y of the query point and b_x > pt_x, so a query point
// above/below the polygon's y-range or right of its x-range (by more than h)
// cannot register a hit on any side -> count == 0 -> False.
// point_in_polygon_2d is the part after the coplanarity test
// |DOT(p - S0, n)| <= eta (geometry.py:641).
__device__ inline bool point_in_polygon_2d(const double *p, const Blocker &k);

__device__ inline bool point_in_polygon(const double *p, const Blocker &k) {
    double d0[3];
    sub3(p, k.s0, d0);
    if (fabs(dot3(d0, k.n)) > kEta) return false;
    return point_in_polygon_2d(p, k);
}

__device__ inline bool point_in_polygon_2d(const double *p, const Blocker &k) {
    const double pt[2] = {dot3(k.r0, p), dot3(k.r1, p)};
    if (pt[1] > k.ymax + k.h || pt[1] < k.ymin - k.h || pt[0] > k.xmax + k.h) return false;
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

// geometry.py:841-909 evaluated for one blocking surface, organised so that the
// common cases cost ~20 FP64 operations.  Only the ORDER of evaluation differs from
// the reference; the predicates are pure, so the value cannot:
//  * inA / inB start with the coplanarity test (geometry.py:641), whose dot products
//    dA, dB are needed anyway (dB is also the numerator of the plane hit parameter);
//  * in the first branch (neither end point in the surface, :881-893) the reference
//    requires  hit exists  AND  hit in polygon  AND  (hit-A).(hit-B) < 0.  The hit is
//    B + fac*v (+ rounding ~1e-13) with fac = -(dB/dp), so (hit-A).(hit-B) =
//    fac(1+fac)|v|^2 is positive whenever fac lies outside [-1, 0] by a margin; that
//    margin test is done on products (no division) with thresholds 2e-3 / -1.002,
//    which imply fac > 1e-3 resp. fac < -1.001 for the rounded quotient as well.
//    v = B - A, cull_ok = |v|^2 > 1e-6.
__device__ __forceinline__ bool blocked(const double *A, const double *B, const double *v,
                                        bool cull_ok, const Blocker &k) {
    double wa[3], w[3];
    sub3(A, k.s0, wa);
    sub3(B, k.s0, w);
    const double dA = dot3(wa, k.n);
    const double dB = dot3(w, k.n);
    bool inA = false, inB = false;
    if (!(fabs(dA) > kEta)) inA = point_in_polygon_2d(A, k);
    if (!(fabs(dB) > kEta)) inB = point_in_polygon_2d(B, k);
    if (!inA && !inB) {
        const double dp = dot3(v, k.n);
        if (!(fabs(dp) > 1e-6)) return false;               // geometry.py:599-604
        if (cull_ok) {
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
__device__ inline int nearest_direction(const double *to, const double *from,
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
