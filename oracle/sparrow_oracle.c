/*
 * sparrow_oracle.c -- CPU restatement of sparrowpy's DirectionalRadiosityFast
 * hot path (reference: sparrow-acoustics/sparrowpy v1.0.1, /root/reference).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test
 * and bench.py's cpu_baseline / --impl reference leg may load it; the product
 * path (sparrowpy_b200/) never does.
 *
 * Parity status: PINNED.  Every function is checked bit-for-bit (integers,
 * booleans) or to <=1e-12 relative (floating point) against vectors produced by
 * the live reference's numba kernels in this container
 * (tests/golden/make_golden.py -> tests/golden/ npz files, tests/test_oracle_golden.py).
 *
 * Arithmetic model of the reference on x86-64 (pinned by
 * tests/golden/rounding_probes.npz):
 *   DOT  numba np.dot (n = 2, 3)        -> BLAS ddot  = FMA chain
 *   NRM  numba 1-D np.linalg.norm       -> BLAS dnrm2 = x87 80-bit sum of squares
 *                                          + fsqrt, rounded to double at the end
 *   everything else: separately rounded IEEE double operations, evaluated in
 *   source order, no contraction.  Compile with -ffp-contract=off.
 *
 * Each function cites the reference file:line it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SOR_PI 3.141592653589793

/* ------------------------------------------------------------------ */
/* rounding-model primitives                                           */
/* ------------------------------------------------------------------ */
static inline double dot3(const double *a, const double *b) {
    double s = a[0] * b[0];
    s = fma(a[1], b[1], s);
    s = fma(a[2], b[2], s);
    return s;
}
static inline double dot2(const double *a, const double *b) {
    double s = a[0] * b[0];
    s = fma(a[1], b[1], s);
    return s;
}
static inline double nrm3(const double *v) {
    long double s = (long double)v[0] * (long double)v[0];
    s += (long double)v[1] * (long double)v[1];
    s += (long double)v[2] * (long double)v[2];
    return (double)sqrtl(s);
}
static inline double nrm2(const double *v) {
    long double s = (long double)v[0] * (long double)v[0];
    s += (long double)v[1] * (long double)v[1];
    return (double)sqrtl(s);
}
static inline void sub3(const double *a, const double *b, double *o) {
    o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2];
}
static inline void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* exported probes so the tests can pin the primitives themselves */
double sor_dot3(const double *a, const double *b) { return dot3(a, b); }
double sor_dot2(const double *a, const double *b) { return dot2(a, b); }
double sor_nrm3(const double *v) { return nrm3(v); }
double sor_nrm2(const double *v) { return nrm2(v); }
/* numpy (not numba) 1-D norm used for distance_i_j, RadiosityFast.py:542 */
double sor_np_norm1d3(const double *v) {
    return sqrt(fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0])));
}
/* numpy norm(axis=1) used at RadiosityFast.py:745 */
double sor_np_norm_axis3(const double *v) {
    return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
}

/* ------------------------------------------------------------------ */
/* tessellation: geometry.py:290-448                                   */
/* ------------------------------------------------------------------ */
static void wall_grid(const double *wall /*4x3*/, double max_size, int64_t nums[3],
                      double size[3], int *x_idx, int *y_idx) {
    /* geometry.py:357-369 / :375-388 */
    for (int k = 0; k < 3; ++k) {
        double mx = wall[k], mn = wall[k];
        for (int v = 1; v < 4; ++v) {
            if (wall[3 * v + k] > mx) mx = wall[3 * v + k];
            if (wall[3 * v + k] < mn) mn = wall[3 * v + k];
        }
        size[k] = mx - mn;
        nums[k] = (int64_t)(size[k] / max_size);
    }
    *x_idx = 0; *y_idx = 1;
    if (nums[2] == 0) { *x_idx = 0; *y_idx = 1; }
    if (nums[1] == 0) { *x_idx = 0; *y_idx = 2; }
    if (nums[0] == 0) { *x_idx = 1; *y_idx = 2; }
}

/* geometry.py:341-371 summed over walls as in :317-321 */
int64_t sor_count_patches(const double *walls, int64_t n_walls, double max_size) {
    int64_t total = 0;
    for (int64_t w = 0; w < n_walls; ++w) {
        int64_t nums[3]; double size[3]; int xi, yi;
        wall_grid(walls + 12 * w, max_size, nums, size, &xi, &yi);
        total += nums[xi] * nums[yi];
    }
    return total;
}

/* geometry.py:373-410 per wall, assembled as in :326-334 */
void sor_create_patches(const double *walls, int64_t n_walls, double max_size,
                        double *patches /*N x4x3*/, int64_t *patch_to_wall) {
    int64_t p = 0;
    for (int64_t w = 0; w < n_walls; ++w) {
        const double *wall = walls + 12 * w;
        int64_t nums[3]; double size[3]; int xi, yi;
        wall_grid(wall, max_size, nums, size, &xi, &yi);
        double rsx = size[xi] / (double)nums[xi];
        double rsy = size[yi] / (double)nums[yi];
        double x_min = wall[xi], y_min = wall[yi];
        for (int v = 1; v < 4; ++v) {
            if (wall[3 * v + xi] < x_min) x_min = wall[3 * v + xi];
            if (wall[3 * v + yi] < y_min) y_min = wall[3 * v + yi];
        }
        for (int64_t ix = 0; ix < nums[xi]; ++ix)
            for (int64_t iy = 0; iy < nums[yi]; ++iy) {
                double *pt = patches + 12 * p;
                memcpy(pt, wall, 12 * sizeof(double));
                double x0 = x_min + (double)ix * rsx, x1 = x_min + (double)(ix + 1) * rsx;
                double y0 = y_min + (double)iy * rsy, y1 = y_min + (double)(iy + 1) * rsy;
                pt[0 + xi] = x0; pt[0 + yi] = y0;
                pt[3 + xi] = x1; pt[3 + yi] = y0;
                pt[9 + xi] = x0; pt[9 + yi] = y1;
                pt[6 + xi] = x1; pt[6 + yi] = y1;
                patch_to_wall[p] = w;
                ++p;
            }
    }
}

/* geometry.py:412-413 */
void sor_centers(const double *pts /*N x4x3*/, int64_t n, double *out /*N x3*/) {
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            const double *p = pts + 12 * i + k;
            double s = 0.0;
            s += p[0]; s += p[3]; s += p[6]; s += p[9];
            out[3 * i + k] = s / 4;
        }
}

/* geometry.py:420-440 */
static double polygon_area(const double *pts, int nv) {
    double area = 0.0;
    for (int t = 0; t < nv - 2; ++t) {
        double a[3], b[3], c[3];
        sub3(pts + 3 * (t + 1), pts, a);
        sub3(pts + 3 * (t + 2), pts, b);
        cross3(a, b, c);
        area += .5 * nrm3(c);
    }
    return area;
}
/* geometry.py:442-448 */
void sor_areas(const double *pts, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = polygon_area(pts + 12 * i, 4);
}

/* ------------------------------------------------------------------ */
/* visibility predicates: geometry.py:498-909                          */
/* ------------------------------------------------------------------ */
/* geometry.py:498-559, target direction (0,0,1) */
static void rotation_matrix(const double *n, double R[9]) {
    if (n[0] == 0.0 && n[1] == 0.0 && n[2] == 1.0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0;
        R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    double nn = nrm3(n);
    double a[3] = {n[0] / nn, n[1] / nn, n[2] / nn};
    double b[3] = {0.0, 0.0, 1.0};          /* n_out / norm(n_out) */
    double c = dot3(a, b);
    if (c != -1) {
        double v[3];
        cross3(a, b, v);
        double s = nrm3(v);
        double K[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        double K2[9];
        for (int i = 0; i < 3; ++i)         /* kmat.dot(kmat): dgemm, FMA over k */
            for (int j = 0; j < 3; ++j) {
                double acc = K[3 * i] * K[j];
                acc = fma(K[3 * i + 1], K[3 + j], acc);
                acc = fma(K[3 * i + 2], K[6 + j], acc);
                K2[3 * i + j] = acc;
            }
        double f = (1 - c) / (s * s);
        for (int i = 0; i < 9; ++i) {
            double eye = (i % 4 == 0) ? 1.0 : 0.0;
            R[i] = (eye + K[i]) + K2[i] * f;
        }
    } else {
        R[0] = -1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0;
        R[6] = 0; R[7] = 0; R[8] = -1;
    }
}

/* geometry.py:561-612 in 2-D with check_normal=False (as called at :668) */
static int project_to_line2(const double *origin, const double *point,
                            const double *plane_pt, const double *nl, double *out) {
    double v[2] = {point[0] - origin[0], point[1] - origin[1]};
    double dp = dot2(v, nl);
    if (!(fabs(dp) > 1e-6)) return 0;
    double w[2] = {point[0] - plane_pt[0], point[1] - plane_pt[1]};
    double fac = -(dot2(nl, w) / dp);
    out[0] = (w[0] + plane_pt[0]) + fac * v[0];
    out[1] = (w[1] + plane_pt[1]) + fac * v[1];
    return 1;
}

/* geometry.py:561-612 in 3-D with check_normal=False (as called at :883) */
static int project_to_plane3(const double *origin, const double *point,
                             const double *plane_pt, const double *n, double *out) {
    double v[3]; sub3(point, origin, v);
    double dp = dot3(v, n);
    if (!(fabs(dp) > 1e-6)) return 0;
    double w[3]; sub3(point, plane_pt, w);
    double fac = -(dot3(n, w) / dp);
    for (int k = 0; k < 3; ++k) out[k] = (w[k] + plane_pt[k]) + fac * v[k];
    return 1;
}

/* geometry.py:614-686 */
static int point_in_polygon(const double *p, const double *poly, int nv,
                            const double *n) {
    const double eta = 1e-6;
    double d0[3]; sub3(p, poly, d0);
    if (fabs(dot3(d0, n)) > eta) return 0;
    double R[9];
    rotation_matrix(n, R);
    double pt[2] = {dot3(R, p), dot3(R + 3, p)};
    double P2[16][2];
    for (int i = 0; i < nv; ++i) {
        P2[i][0] = dot3(R, poly + 3 * i);
        P2[i][1] = dot3(R + 3, poly + 3 * i);
    }
    int count = 0;
    for (int i = 0; i < nv; ++i) {
        const double *a1 = P2[(i + 1) % nv];
        const double *a0 = P2[i % nv];
        double side[2] = {a1[0] - a0[0], a1[1] - a0[1]};
        double ns = nrm2(side);
        double nl[2] = {-side[1] / ns, side[0] / ns};
        double pt1[2] = {pt[0] + 1., pt[1] + 0.};
        double b[2];
        if (project_to_line2(pt, pt1, a1, nl, b) && b[0] > pt[0]) {
            double ba0[2] = {b[0] - a0[0], b[1] - a0[1]};
            double ba1[2] = {b[0] - a1[0], b[1] - a1[1]};
            double a10[2] = {a1[0] - a0[0], a1[1] - a0[1]};
            if (fabs(nrm2(ba0) + nrm2(ba1) - nrm2(a10)) <= eta) {
                double bp[2] = {b[0] - pt[0], b[1] - pt[1]};
                double d = dot2(bp, nl);
                if (d > 0) count += 1;
                else if (d < 0) count -= 1;
            }
        }
    }
    return count != 0;
}

/* geometry.py:841-909 */
static int basic_visibility(const double *A, const double *B, const double *S,
                            int nv, const double *n) {
    const double eta = 1e-6;
    int visible = 1;
    int inA = point_in_polygon(A, S, nv, n);
    int inB = point_in_polygon(B, S, nv, n);
    double d[3];
    if (!inA && !inB) {
        double pt[3];
        if (project_to_plane3(A, B, S, n, pt)) {
            if (point_in_polygon(pt, S, nv, n)) {
                double pa[3], pb[3];
                sub3(pt, A, pa); sub3(pt, B, pb);
                if (dot3(pa, pb) < 0) visible = 0;
            }
        }
    } else if (inA && !inB && (sub3(B, A, d), dot3(n, d) < 0)) {
        visible = 0;
    } else if (!inA && inB && (sub3(A, B, d), dot3(n, d) < 0)) {
        visible = 0;
    } else {
        double da[3], db[3];
        sub3(A, S, da); sub3(B, S, db);
        if (fabs(dot3(da, n)) < eta && fabs(dot3(db, n)) < eta && (inA || inB))
            visible = 0;
    }
    return visible;
}

int sor_point_in_polygon(const double *p, const double *poly, int nv, const double *n) {
    return point_in_polygon(p, poly, nv, n);
}
int sor_basic_visibility(const double *A, const double *B, const double *S, int nv,
                         const double *n) {
    return basic_visibility(A, B, S, nv, n);
}

/* geometry.py:750-797.  vis is (N,N) uint8, upper triangle only.
 * Optional row range [row_lo,row_hi) for sampled timing/parity at large N. */
void sor_visibility_p2p_rows(const double *centers, const double *surf_normals,
                             const double *surf_points, int64_t n, int64_t m, int nv,
                             int64_t row_lo, int64_t row_hi, uint8_t *vis) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = row_lo; i < row_hi; ++i) {
        uint8_t *row = vis + (i - row_lo) * n;
        memset(row, 0, (size_t)n);
        for (int64_t j = i + 1; j < n; ++j) {
            int v = 1;
            for (int64_t s = 0; s < m && v; ++s)
                v = basic_visibility(centers + 3 * i, centers + 3 * j,
                                     surf_points + 3 * nv * s, nv,
                                     surf_normals + 3 * s);
            row[j] = (uint8_t)v;
        }
    }
}
void sor_visibility_p2p(const double *centers, const double *surf_normals,
                        const double *surf_points, int64_t n, int64_t m, int nv,
                        uint8_t *vis) {
    sor_visibility_p2p_rows(centers, surf_normals, surf_points, n, m, nv, 0, n, vis);
}

/* geometry.py:799-839 */
void sor_visibility_pt2p(const double *point, const double *centers,
                         const double *surf_normals, const double *surf_points,
                         int64_t n, int64_t m, int nv, uint8_t *vis) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        int v = 1;
        for (int64_t s = 0; s < m && v; ++s)
            v = basic_visibility(point, centers + 3 * i, surf_points + 3 * nv * s, nv,
                                 surf_normals + 3 * s);
        vis[i] = (uint8_t)v;
    }
}

/* ------------------------------------------------------------------ */
/* form factors: form_factor/integration.py, universal.py              */
/* ------------------------------------------------------------------ */
/* integration.py:608-651 */
static void sample_boundary(const double *el /*4x3*/, int npoints, double *pts) {
    int n_div = npoints - 1;
    for (int i = 0; i < 4; ++i) {
        const double *p0 = el + 3 * i, *p1 = el + 3 * ((i + 1) % 4);
        for (int ii = 0; ii < n_div; ++ii)
            for (int k = 0; k < 3; ++k)
                pts[3 * (i * n_div + ii) + k] =
                    p0[k] + (double)ii * (p1[k] - p0[k]) / (double)n_div;
    }
}

/* integration.py:462-486 */
static inline double boole(double x0, double x1, const double *y) {
    double h = x1 - x0;
    return 2 * h / 45 * (7 * y[0] + 32 * y[1] + 12 * y[2] + 32 * y[3] + 7 * y[4]);
}

/* integration.py:38-114 */
static double stokes_integration(const double *pi, const double *pj, double area_i) {
    double bi[16 * 3], bj[16 * 3];
    sample_boundary(pi, 5, bi);
    sample_boundary(pj, 5, bj);
    int conn[4][5];
    for (int i = 0; i < 4; ++i) {
        for (int k = 0; k < 4; ++k) conn[i][k] = (4 * i + k) % 16;
        conn[i][4] = (4 * i + 4) % 16;
    }
    double form[16][16];
    for (int a = 0; a < 16; ++a)            /* integration.py:30-34 */
        for (int b = 0; b < 16; ++b) {
            double d[3]; sub3(bi + 3 * a, bj + 3 * b, d);
            form[a][b] = log(nrm3(d));
        }
    double outer = 0.0;
    double inner[16][3];
    memset(inner, 0, sizeof(inner));
    for (int dim = 0; dim < 3; ++dim) {
        for (int a = 0; a < 16; ++a)
            for (int s = 0; s < 4; ++s) {
                const int *seg = conn[s];
                double xl = bj[3 * seg[4] + dim], x0 = bj[3 * seg[0] + dim];
                if (fabs(xl - x0) > 1e-3) {
                    double y[5];
                    for (int k = 0; k < 5; ++k) y[k] = form[a][seg[k]];
                    inner[a][dim] += boole(x0, bj[3 * seg[1] + dim], y);
                }
            }
        for (int s = 0; s < 4; ++s) {
            const int *seg = conn[s];
            double xl = bi[3 * seg[4] + dim], x0 = bi[3 * seg[0] + dim];
            if (fabs(xl - x0) > 1e-3) {
                double y[5];
                for (int k = 0; k < 5; ++k) y[k] = inner[seg[k]][dim];
                outer += boole(x0, bi[3 * seg[1] + dim], y);
            }
        }
    }
    return fabs(outer / (2 * SOR_PI * area_i));
}

/* integration.py:416-460 with :349-412 folded in (order 2, three samples).
 * The reference solves the 3x3 Vandermonde system with np.linalg.inv (LAPACK);
 * here it is solved in closed form -- tolerance path (<=1e-12 rel. observed). */
static double area_under_curve(const double ps[3][2]) {
    double f[2] = {ps[2][0] - ps[0][0], ps[2][1] - ps[0][1]};
    double nf = nrm2(f);
    double r0[2] = {f[0] / nf, f[1] / nf};
    double r1[2] = {-f[1] / nf, f[0] / nf};
    double x[3] = {0, 0, 0}, y[3] = {0, 0, 0};
    for (int k = 1; k < 3; ++k) {
        double c[2] = {ps[k][0] - ps[0][0], ps[k][1] - ps[0][1]};
        x[k] = dot2(r0, c);
        y[k] = dot2(r1, c);
    }
    if (fabs(x[2] - x[0]) < 1e-6) return 0.0;     /* integration.py:373-374 */
    /* quadratic through (x0,y0)=(0,0), (x1,y1), (x2,y2): y = c0 x^2 + c1 x + c2 */
    double det = x[1] * x[2] * (x[1] - x[2]);
    double c0 = (y[1] * x[2] - y[2] * x[1]) / det;
    double c1 = (y[2] * x[1] * x[1] - y[1] * x[2] * x[2]) / det;
    double c2 = 0.0;
    double out = 0.0;                              /* integration.py:406-412 */
    out += c0 * (x[2] * x[2] * x[2]) / 3; out -= c0 * (x[0] * x[0] * x[0]) / 3;
    out += c1 * (x[2] * x[2]) / 2;        out -= c1 * (x[0] * x[0]) / 2;
    out += c2 * x[2] / 1;                 out -= c2 * x[0] / 1;
    return out;
}

static inline double sign(double v) { return (v > 0) - (v < 0); }

/* integration.py:116-230 */
static double nusselt_analog(const double *o, const double *n_i, const double *pj,
                             const double *n_j) {
    double bp[8 * 3];
    sample_boundary(pj, 3, bp);
    double e0[3], e1[3], cr[3];
    sub3(pj + 3, pj, e0); sub3(pj + 6, pj + 3, e1);
    cross3(e0, e1, cr);
    double hand = sign(dot3(cr, n_j));
    double sph[8][3], pln[8][2], proj[8][3];
    for (int i = 0; i < 8; ++i) {
        double d[3]; sub3(bp + 3 * i, o, d);
        double nd = nrm3(d);
        for (int k = 0; k < 3; ++k) sph[i][k] = d[k] / nd;
    }
    double R[9];
    rotation_matrix(n_i, R);
    for (int i = 0; i < 8; ++i) {
        pln[i][0] = dot3(R, sph[i]);
        pln[i][1] = dot3(R + 3, sph[i]);
        proj[i][0] = pln[i][0]; proj[i][1] = pln[i][1]; proj[i][2] = 0.;
    }
    double quad[12];
    for (int i = 0; i < 4; ++i) memcpy(quad + 3 * i, proj[2 * i], 3 * sizeof(double));
    double big_poly = polygon_area(quad, 4);
    double curved = 0.0;
    for (int jj = 0; jj < 4; ++jj) {
        int s0 = 2 * jj, s1 = 2 * jj + 1, s2 = (2 * jj + 2) % 8;
        double cp[3];
        cross3(proj[s2], proj[s0], cp);
        if (nrm3(cp) > 1e-6) {
            if (dot2(pln[s2], pln[s0]) >= 1e-6) {
                double ps[3][2] = {{pln[s0][0], pln[s0][1]}, {pln[s1][0], pln[s1][1]},
                                   {pln[s2][0], pln[s2][1]}};
                curved += area_under_curve(ps);
            } else {
                double mp[3], marc[3], a[3], b[3];
                for (int k = 0; k < 3; ++k)
                    mp[k] = sph[s0][k] + (sph[s2][k] - sph[s0][k]) / 2;
                double nm = nrm3(mp);
                for (int k = 0; k < 3; ++k) marc[k] = mp[k] / nm;
                for (int k = 0; k < 3; ++k) {
                    a[k] = sph[s0][k] + (marc[k] - sph[s0][k]) / 2;
                    b[k] = marc[k] + (sph[s2][k] - marc[k]) / 2;
                }
                double mp2[2] = {dot3(R, mp), dot3(R + 3, mp)};
                double marc2[2] = {dot3(R, marc), dot3(R + 3, marc)};
                double na = nrm3(a);
                for (int k = 0; k < 3; ++k) a[k] = a[k] / na;
                double a2[2] = {dot3(R, a), dot3(R + 3, a)};
                double nb = nrm3(b);
                for (int k = 0; k < 3; ++k) b[k] = b[k] / nb;
                double b2[2] = {dot3(R, b), dot3(R + 3, b)};
                double d1[2] = {pln[s2][0] - pln[s0][0], pln[s2][1] - pln[s0][1]};
                double d2[2] = {mp2[0] - marc2[0], mp2[1] - marc2[1]};
                double lin = nrm2(d1) * nrm2(d2) / 2;
                double ls[3][2] = {{pln[s0][0], pln[s0][1]}, {a2[0], a2[1]},
                                   {marc2[0], marc2[1]}};
                double rs[3][2] = {{marc2[0], marc2[1]}, {b2[0], b2[1]},
                                   {pln[s2][0], pln[s2][1]}};
                double left = area_under_curve(ls);
                double right = area_under_curve(rs);
                curved += (lin * sign(left) + left + right);
            }
        }
    }
    return big_poly + hand * curved;
}

/* integration.py:533-605 for quadrilaterals, feeding :232-289 */
static double nusselt_integration(const double *pi, const double *pj,
                                  const double *n_i, const double *n_j, int nsamples) {
    double u[3], v[3];
    sub3(pi + 3, pi, u);
    sub3(pi + 9, pi, v);
    double nu = nrm3(u), nvv = nrm3(v);
    int npx = (int)rint(nu / nvv * sqrt(1.0 * nsamples));
    int npz = (int)rint(nvv / nu * sqrt(1.0 * nsamples));
    if (npz == 0) npz = 1;
    if (npx == 0) npx = 1;
    double out = 0.0;
    int count = 0;
    double sstep = 1.0 / (npx * 2), sstepz = 1.0 / (npz * 2);
    for (int i = 0; i < npx; ++i) {
        /* np.linspace(0, 1-1/npx, npx)[i] + sstep */
        double stop = 1 - 1.0 / npx;
        double s = (npx > 1) ? (0.0 + i * ((stop - 0.0) / (npx - 1))) : 0.0;
        if (npx > 1 && i == npx - 1) s = stop;
        s += sstep;
        for (int j = 0; j < npz; ++j) {
            double stopz = 1 - 1.0 / npz;
            double t = (npz > 1) ? (0.0 + j * ((stopz - 0.0) / (npz - 1))) : 0.0;
            if (npz > 1 && j == npz - 1) t = stopz;
            t += sstepz;
            double o[3];
            for (int k = 0; k < 3; ++k) o[k] = s * u[k] + t * v[k] + pi[k];
            out += nusselt_analog(o, n_i, pj, n_j);
            ++count;
        }
    }
    out *= 1 / (SOR_PI * count);
    return out;
}

/* geometry.py:719-748 */
static int coincidence_check(const double *p0, const double *p1) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double d[3]; sub3(p0 + 3 * i, p1 + 3 * j, d);
            if (nrm3(d) < 1e-6) return 1;
        }
    return 0;
}

/* universal.py:54-96 */
double sor_universal_form_factor(const double *pts_i, const double *n_i, double area_i,
                                 const double *pts_j, const double *n_j) {
    if (coincidence_check(pts_j, pts_i))
        return nusselt_integration(pts_i, pts_j, n_i, n_j, 64);
    return stokes_integration(pts_i, pts_j, area_i);
}
int sor_coincidence_check(const double *p0, const double *p1) {
    return coincidence_check(p0, p1);
}

/* universal.py:12-52, output per visible pair (the reference scatters these into a
 * dense (N,N) matrix at [i,j], i<j) */
void sor_ff_pairs(const double *pts, const double *normals, const double *areas,
                  const int32_t *pairs, int64_t n_pairs, double *ff) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t p = 0; p < n_pairs; ++p) {
        int64_t i = pairs[2 * p], j = pairs[2 * p + 1];
        ff[p] = sor_universal_form_factor(pts + 12 * i, normals + 3 * i, areas[i],
                                          pts + 12 * j, normals + 3 * j);
    }
}

/* geometry.py:688-715 */
static void sphere_tangent(const double *v0, const double *v1, double *out) {
    if (fabs(dot3(v0, v1)) > 1e-10) {
        double d[3]; sub3(v1, v0, d);
        double q = dot3(d, v0) / dot3(v0, v0);
        for (int k = 0; k < 3; ++k) out[k] = d[k] - q * v0[k];
        double nn = nrm3(out);
        for (int k = 0; k < 3; ++k) out[k] /= nn;
    } else {
        double nn = nrm3(v1);
        for (int k = 0; k < 3; ++k) out[k] = v1[k] / nn;
    }
}

/* integration.py:295-344; mode 0 = source, 1 = receiver */
double sor_pt_solution(const double *point, const double *patch /*4x3*/, int mode) {
    double source_area = (mode == 1) ? polygon_area(patch, 4) : 4.0;
    double sph[4][3];
    for (int i = 0; i < 4; ++i) {
        double d[3]; sub3(patch + 3 * i, point, d);
        double nn = nrm3(d);
        for (int k = 0; k < 3; ++k) sph[i][k] = d[k] / nn;
    }
    double sum = 0.0;
    for (int i = 0; i < 4; ++i) {
        double v0[3], v1[3];
        sphere_tangent(sph[i], sph[(i + 3) % 4], v0);
        sphere_tangent(sph[i], sph[(i + 1) % 4], v1);
        sum += acos(dot3(v0, v1));
    }
    double factor = sum - (4 - 2) * SOR_PI;
    return factor / (SOR_PI * source_area);
}

/* universal.py:98-147 */
void sor_source_energy(const double *src, const double *centers, const double *pts,
                       const uint8_t *vis, const double *air, int64_t n, int64_t n_bins,
                       double *energy /*N x B*/, double *distance /*N*/) {
    for (int64_t j = 0; j < n; ++j) {
        distance[j] = 0.0;
        for (int64_t b = 0; b < n_bins; ++b) energy[j * n_bins + b] = 0.0;
        if (!vis[j]) continue;
        double d[3]; sub3(src, centers + 3 * j, d);
        distance[j] = nrm3(d);
        double g = sor_pt_solution(src, pts + 12 * j, 0);
        for (int64_t b = 0; b < n_bins; ++b)
            energy[j * n_bins + b] = exp(-air[b] * distance[j]) * g;
    }
}

/* universal.py:149-160 */
void sor_receiver_factor(const double *rcv, const double *pts, const uint8_t *vis,
                         int64_t n, double *factor) {
    for (int64_t i = 0; i < n; ++i)
        factor[i] = vis[i] ? sor_pt_solution(rcv, pts + 12 * i, 1) : 0.0;
}

/* ------------------------------------------------------------------ */
/* BRDF direction lookup: RadiosityFast.py:1277-1312, :1359-1390       */
/* ------------------------------------------------------------------ */
/* argmin_k ||dirs[k] - unit(to - from)||^2, first minimum wins */
static int64_t nearest_direction(const double *to, const double *from,
                                 const double *dirs, int64_t n_dirs) {
    double diff[3]; sub3(to, from, diff);
    double nn = nrm3(diff);
    for (int k = 0; k < 3; ++k) diff[k] /= nn;
    int64_t best = 0; double bestv = 0.0;
    for (int64_t k = 0; k < n_dirs; ++k) {
        double e0 = dirs[3 * k] - diff[0], e1 = dirs[3 * k + 1] - diff[1],
               e2 = dirs[3 * k + 2] - diff[2];
        double v = (e0 * e0 + e1 * e1) + e2 * e2;
        if (k == 0 || v < bestv) { bestv = v; best = k; }
    }
    return best;
}
int64_t sor_nearest_direction(const double *to, const double *from, const double *dirs,
                              int64_t n_dirs) {
    return nearest_direction(to, from, dirs, n_dirs);
}

/* RadiosityFast.py:1277-1312: index of wall(i)'s outgoing direction towards pos_j */
void sor_receiver_dir_index(const double *pos_i, int64_t n, const double *pos_j,
                            const double *vo /*W x D x3*/, int64_t n_dirs,
                            const int64_t *wall_id, int64_t *out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = nearest_direction(pos_j, pos_i + 3 * i, vo + 3 * n_dirs * wall_id[i],
                                   n_dirs);
}

/* RadiosityFast.py:988-1034 */
void sor_add_directional(const double *energy0 /*N x B*/, const double *src,
                         const double *centers, const int64_t *patch_to_wall,
                         const double *vi /*W x S x3*/, int64_t n_in, int64_t n_out,
                         const double *brdf /*n_brdf x S x D x B*/,
                         const int64_t *brdf_index, int64_t n, int64_t n_bins,
                         double *out /*N x D x B*/) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t w = patch_to_wall[i];
        int64_t s = nearest_direction(src, centers + 3 * i, vi + 3 * n_in * w, n_in);
        const double *row = brdf + ((brdf_index[w] * n_in + s) * n_out) * n_bins;
        for (int64_t d = 0; d < n_out; ++d)
            for (int64_t b = 0; b < n_bins; ++b)
                out[(i * n_out + d) * n_bins + b] =
                    energy0[i * n_bins + b] * row[d * n_bins + b];
    }
}

/* ------------------------------------------------------------------ */
/* directed pair tables = factored form_factors_tilde                  */
/* RadiosityFast.py:1234-1272 (+ :403-414 for p2o, :538-543 for delay) */
/* ------------------------------------------------------------------ */
/* For visible pair row p = (i,j), i<j, directed entries are stored as
 * [2p] = i->j and [2p+1] = j->i (the order the exchange loop visits them,
 * RadiosityFast.py:1124-1131).
 *   tilde[(2p+e), d, b], out_dir[2p+e] (sender's outgoing index), delay[2p+e] */
void sor_pair_tables(const double *centers, const double *areas,
                     const int64_t *patch_to_wall, const int32_t *pairs,
                     const double *ff_pairs, int64_t n_pairs, const double *air,
                     const double *vi, int64_t n_in, const double *vo, int64_t n_out,
                     const double *brdf, const int64_t *brdf_index, int64_t n_bins,
                     double speed_of_sound, double dt, double *tilde, int64_t *out_dir,
                     int64_t *in_dir, int64_t *delay) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n_pairs; ++p) {
        int64_t lo = pairs[2 * p], hi = pairs[2 * p + 1];
        for (int e = 0; e < 2; ++e) {
            int64_t i = e == 0 ? lo : hi;       /* sender */
            int64_t j = e == 0 ? hi : lo;       /* receiver */
            /* :1252-1258 */
            double diff[3]; sub3(centers + 3 * i, centers + 3 * j, diff);
            double nn = nrm3(diff);
            for (int k = 0; k < 3; ++k) diff[k] /= nn;
            double ff = (i < j) ? ff_pairs[p] : (ff_pairs[p] * areas[j] / areas[i]);
            double distance = nrm3(diff);       /* quirk: norm of the unit vector */
            int64_t w = patch_to_wall[i];
            /* :1386-1390 incoming index on the SENDER's wall */
            int64_t s = nearest_direction(centers + 3 * i, centers + 3 * j,
                                          vi + 3 * n_in * w, n_in);
            const double *row = brdf + ((brdf_index[w] * n_in + s) * n_out) * n_bins;
            double *t = tilde + (2 * p + e) * n_out * n_bins;
            for (int64_t d = 0; d < n_out; ++d)
                for (int64_t b = 0; b < n_bins; ++b)
                    t[d * n_bins + b] =
                        (ff * exp(-air[b] * distance)) * row[d * n_bins + b];
            in_dir[2 * p + e] = s;
            /* :407-414 */
            out_dir[2 * p + e] = nearest_direction(centers + 3 * j, centers + 3 * i,
                                                   vo + 3 * n_out * w, n_out);
            /* :538-543 numpy 1-D norm; :1135-1136 */
            double dd[3]; sub3(centers + 3 * i, centers + 3 * j, dd);
            delay[2 * p + e] =
                (int64_t)(sor_np_norm1d3(dd) / speed_of_sound / dt);
        }
    }
}

/* ------------------------------------------------------------------ */
/* energy exchange: RadiosityFast.py:1037-1145                         */
/* ------------------------------------------------------------------ */
/* RadiosityFast.py:1037-1070.  Energy whose delay bin is >= T is dropped (the
 * reference writes out of bounds there). */
void sor_init_energy(const double *e0 /*N x D x B*/, const double *distance0,
                     int64_t n, int64_t n_dirs, int64_t n_bins, int64_t n_samples,
                     double speed_of_sound, double dt, double *etc /*N x D x B x T*/) {
    memset(etc, 0, sizeof(double) * (size_t)(n * n_dirs * n_bins * n_samples));
    for (int64_t i = 0; i < n; ++i) {
        int64_t delay = (int64_t)(distance0[i] / speed_of_sound / dt);
        if (delay >= n_samples) continue;
        for (int64_t c = 0; c < n_dirs * n_bins; ++c)
            etc[(i * n_dirs * n_bins + c) * n_samples + delay] += e0[i * n_dirs * n_bins + c];
    }
}

/* RadiosityFast.py:1073-1145.  Same loop nest and accumulation order as the
 * reference (pair-major, i->j then j->i), so results are bit-identical to it.
 * With n_threads > 1 the time axis is cut into slices, one per thread: every
 * element still receives its contributions in the reference order. */
void sor_energy_exchange(const double *e0, const double *distance0, const int32_t *pairs,
                         int64_t n_pairs, const double *tilde /*2P x D x B*/,
                         const int64_t *out_dir, const int64_t *delay, int64_t n,
                         int64_t n_dirs, int64_t n_bins, int64_t n_samples,
                         double speed_of_sound, double dt, int64_t max_order,
                         int n_threads, double *etc_total, double *work /*2 x N*D*B*T*/) {
    const int64_t T = n_samples, DB = n_dirs * n_bins;
    const int64_t tot = n * DB * T;
    sor_init_energy(e0, distance0, n, n_dirs, n_bins, T, speed_of_sound, dt, etc_total);
    if (max_order < 1) return;
    memcpy(work, etc_total, sizeof(double) * (size_t)tot);
    memset(work + tot, 0, sizeof(double) * (size_t)tot);
    if (n_threads < 1) n_threads = 1;
    for (int64_t k = 0; k < max_order; ++k) {
        double *cur = work + ((1 + k) % 2) * tot;
        const double *prev = work + (k % 2) * tot;
#pragma omp parallel num_threads(n_threads)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
            int tid = 0, nt = 1;
#endif
            int64_t t_lo = T * tid / nt, t_hi = T * (tid + 1) / nt;
            for (int64_t r = 0; r < n * DB; ++r)
                memset(cur + r * T + t_lo, 0, sizeof(double) * (size_t)(t_hi - t_lo));
            for (int64_t q = 0; q < 2 * n_pairs; ++q) {
                int64_t p = q >> 1;
                int64_t i = (q & 1) ? pairs[2 * p + 1] : pairs[2 * p];
                int64_t j = (q & 1) ? pairs[2 * p] : pairs[2 * p + 1];
                int64_t dl = delay[q];
                int64_t lo = t_lo > dl ? t_lo : dl;
                if (lo >= t_hi) continue;
                const double *tq = tilde + q * DB;
                for (int64_t d = 0; d < n_dirs; ++d)
                    for (int64_t b = 0; b < n_bins; ++b) {
                        double w = tq[d * n_bins + b];
                        const double *src = prev + ((i * n_dirs + out_dir[q]) * n_bins + b) * T - dl;
                        double *dst = cur + ((j * n_dirs + d) * n_bins + b) * T;
                        for (int64_t t = lo; t < t_hi; ++t) dst[t] += w * src[t];
                    }
            }
            for (int64_t r = 0; r < n * DB; ++r)
                for (int64_t t = t_lo; t < t_hi; ++t)
                    etc_total[r * T + t] += cur[r * T + t];
        }
    }
}

/* ------------------------------------------------------------------ */
/* receiver collection: RadiosityFast.py:686-752, :1148-1185           */
/* ------------------------------------------------------------------ */
/* One receiver.  out is (N, B, T) patch-wise; mono (B, T) = sum over patches in
 * patch order (np.sum(axis=1), RadiosityFast.py:592). */
void sor_collect_receiver(const double *etc /*N x D x B x T*/, const double *rcv,
                          const double *centers, const double *factor /*N*/,
                          const int64_t *dir_index /*N*/, const double *air,
                          int64_t n, int64_t n_dirs, int64_t n_bins, int64_t n_samples,
                          double speed_of_sound, double dt, double *patchwise /*or NULL*/,
                          double *mono /*B x T*/, int64_t *delays_out /*N or NULL*/) {
    const int64_t T = n_samples;
    memset(mono, 0, sizeof(double) * (size_t)(n_bins * T));
    for (int64_t k = 0; k < n; ++k) {
        double d[3]; sub3(centers + 3 * k, rcv, d);
        double dist = sor_np_norm_axis3(d);
        int64_t delay = (int64_t)ceil(dist / speed_of_sound / dt);
        if (delays_out) delays_out[k] = delay;
        int64_t shift = ((delay % T) + T) % T;       /* np.roll is circular */
        for (int64_t b = 0; b < n_bins; ++b) {
            const double *row = etc + ((k * n_dirs + dir_index[k]) * n_bins + b) * T;
            double att = exp(-air[b] * dist);
            for (int64_t t = 0; t < T; ++t) {
                double v = (row[t] * factor[k]) * att;
                int64_t tt = t + shift; if (tt >= T) tt -= T;
                if (patchwise) patchwise[(k * n_bins + b) * T + tt] = v;
                mono[b * T + tt] += v;
            }
        }
    }
}

int sor_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
