// Host check of sparrowpy_b200/csrc/exact.cuh (the visibility predicate with its
// result-preserving shortcuts) against the CPU oracle's literal restatement of
// `_basic_visibility` (oracle/sparrow_oracle.c), on random and degenerate cases.
// Built and run by tests/test_host_cpu.py (g++ -ffp-contract=off, x86-64).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../sparrowpy_b200/csrc/exact.cuh"

extern "C" int sor_basic_visibility(const double *A, const double *B, const double *S, int nv,
                                    const double *n);
extern "C" int sor_point_in_polygon(const double *p, const double *poly, int nv,
                                    const double *n);

using spb::exact::Blocker;

static long total = 0, bad = 0, n_blocked = 0, n_in = 0;

static void check(const double *A, const double *B, const double *S, const double *n) {
    Blocker k;
    spb::exact::make_blocker(S, n, k);
    double v[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    const double vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const bool mine = spb::exact::blocked(A, B, v, std::sqrt(vv), vv > 1e-6, k);
    const bool ref = !sor_basic_visibility(A, B, S, 4, n);
    const bool in_mine = spb::exact::point_in_polygon(A, k);
    const bool in_ref = sor_point_in_polygon(A, S, 4, n);
    ++total;
    n_blocked += ref;
    n_in += in_ref;
    if (mine != ref || in_mine != in_ref) {
        if (bad < 10)
            printf("MISMATCH blocked %d/%d in %d/%d A=(%a %a %a) B=(%a %a %a)\n", mine, ref,
                   in_mine, in_ref, A[0], A[1], A[2], B[0], B[1], B[2]);
        ++bad;
    }
}

int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 200000;
    std::mt19937_64 rng(2024);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    auto uni = [&](double lo, double hi) { return lo + (hi - lo) * ud(rng); };

    for (long it = 0; it < n; ++it) {
        // ---- tilted rectangle, random segment, end points sometimes in the plane ----
        double a[3], b[3], nn[3], o[3];
        for (int c = 0; c < 3; ++c) { a[c] = nd(rng); b[c] = nd(rng); o[c] = uni(-2, 2); }
        double na = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        for (int c = 0; c < 3; ++c) a[c] /= na;
        double ab = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
        for (int c = 0; c < 3; ++c) b[c] -= ab * a[c];
        double nb = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        for (int c = 0; c < 3; ++c) b[c] /= nb;
        nn[0] = a[1] * b[2] - a[2] * b[1];
        nn[1] = a[2] * b[0] - a[0] * b[2];
        nn[2] = a[0] * b[1] - a[1] * b[0];
        const double la = uni(0.3, 2.0), lb = uni(0.3, 2.0);
        double S[12];
        for (int c = 0; c < 3; ++c) {
            S[c] = o[c]; S[3 + c] = o[c] + la * a[c];
            S[6 + c] = o[c] + la * a[c] + lb * b[c]; S[9 + c] = o[c] + lb * b[c];
        }
        double A[3], B[3];
        for (int c = 0; c < 3; ++c) { A[c] = uni(-3, 3); B[c] = uni(-3, 3); }
        const int mode = (int)(it % 4);
        if (mode >= 1) {   // A in the plane, inside / around the rectangle (incl. left of it)
            const double s = uni(-1.5, 1.5), t = uni(-0.3, 1.3);
            for (int c = 0; c < 3; ++c) A[c] = o[c] + s * la * a[c] + t * lb * b[c];
        }
        if (mode == 3) {   // both in the plane
            const double s = uni(-1.5, 1.5), t = uni(-0.3, 1.3);
            for (int c = 0; c < 3; ++c) B[c] = o[c] + s * la * a[c] + t * lb * b[c];
        }
        check(A, B, S, nn);
        check(B, A, S, nn);

        // ---- axis-aligned lattice: patches of a wall, segments between lattice points ----
        static const double steps[] = {1.0, 0.5, 0.2, 1.0 / 3.0, 0.25, 0.3};
        const double h = steps[it % 6];
        const int axis = (int)(rng() % 3), o1 = (axis + 1) % 3, o2 = (axis + 2) % 3;
        const double sign = (rng() & 1) ? 1.0 : -1.0;
        const long i1 = (long)(rng() % 9) - 4, i2 = (long)(rng() % 9) - 4, i0 = (long)(rng() % 9) - 4;
        const long w1 = 1 + (long)(rng() % 3), w2 = 1 + (long)(rng() % 3);
        double P[12];
        for (int vtx = 0; vtx < 4; ++vtx) {
            P[3 * vtx + axis] = i0 * h;
            P[3 * vtx + o1] = (i1 + ((vtx == 1 || vtx == 2) ? w1 : 0)) * h;
            P[3 * vtx + o2] = (i2 + ((vtx >= 2) ? w2 : 0)) * h;
        }
        double N[3] = {0, 0, 0};
        N[axis] = sign;
        double LA[3], LB[3];
        for (int c = 0; c < 3; ++c) {
            LA[c] = ((long)(rng() % 17) - 8) * 0.5 * h;
            LB[c] = ((long)(rng() % 17) - 8) * 0.5 * h;
        }
        if (it % 3 == 0) LA[axis] = i0 * h;           // end point in the wall's plane
        if (it % 5 == 0) LB[axis] = i0 * h;
        check(LA, LB, P, N);
        check(LB, LA, P, N);
    }
    printf("checked %ld cases (%ld blocked, %ld end points in surface), %ld mismatches\n", total,
           n_blocked, n_in, bad);
    return bad ? 1 : 0;
}
