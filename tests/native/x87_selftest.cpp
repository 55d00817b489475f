// Host self-test of sparrowpy_b200/csrc/x87.cuh against the CPU's real x87
// `long double` (x86-64 only).  Built and run by tests/test_x87_cpu.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../sparrowpy_b200/csrc/x87.cuh"

static double ref3(double a, double b, double c) {
    volatile long double s = (long double)a * (long double)a;
    s += (long double)b * (long double)b;
    s += (long double)c * (long double)c;
    return (double)sqrtl(s);
}
static double ref2(double a, double b) {
    volatile long double s = (long double)a * (long double)a;
    s += (long double)b * (long double)b;
    return (double)sqrtl(s);
}

static long bad = 0, total = 0;
static void check3(double a, double b, double c) {
    double r = ref3(a, b, c), e = x87::norm3(a, b, c);
    ++total;
    if (memcmp(&r, &e, 8) != 0) {
        if (bad < 10) printf("MISMATCH3 %a %a %a ref=%a emu=%a\n", a, b, c, r, e);
        ++bad;
    }
}
static void check2(double a, double b) {
    double r = ref2(a, b), e = x87::norm2(a, b);
    ++total;
    if (memcmp(&r, &e, 8) != 0) {
        if (bad < 10) printf("MISMATCH2 %a %a ref=%a emu=%a\n", a, b, r, e);
        ++bad;
    }
}

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 2000000;
    std::mt19937_64 rng(12345);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::uniform_int_distribution<int> ex(-40, 40), lat(-4000, 4000), small(-8, 8);
    for (long i = 0; i < n; ++i) {
        // wide dynamic range, independent exponents
        check3(ldexp(nd(rng), ex(rng)), ldexp(nd(rng), ex(rng)), ldexp(nd(rng), ex(rng)));
        check2(ldexp(nd(rng), ex(rng)), ldexp(nd(rng), ex(rng)));
        // lattice differences (patch centres): multiples of 0.05, 0.125, 1/3
        check3(lat(rng) * 0.05, lat(rng) * 0.05, lat(rng) * 0.05);
        check3(lat(rng) * 0.125, lat(rng) * 0.125, lat(rng) * 0.125);
        check3(lat(rng) / 3.0, lat(rng) / 3.0, lat(rng) / 3.0);
        check2(lat(rng) * 0.1, lat(rng) * 0.1);
        // axis-aligned (two zero components) and tiny integers: exact roots, ties
        check3(small(rng), small(rng), small(rng));
        check3(0.0, lat(rng) * 0.1, 0.0);
        check2(small(rng), small(rng));
        // near-unit vectors (normalised directions)
        double a = nd(rng), b = nd(rng), c = nd(rng);
        double nn = std::sqrt(a * a + b * b + c * c);
        check3(a / nn, b / nn, c / nn);
    }
    // edge cases
    double edge[] = {0.0, 1.0, -1.0, 0x1p-1074, 0x1p-1022, 0x1.fffffffffffffp-1023,
                     0x1.fffffffffffffp+500, 0x1p+511, 3.0, 4.0, 1e-160, 1e150,
                     0x1.fffffffffffffp0, 0x1.0000000000001p0};
    for (double a : edge) for (double b : edge) {
        check2(a, b);
        for (double c : edge) check3(a, b, c);
    }
    printf("checked %ld cases, %ld mismatches\n", total, bad);
    return bad ? 1 : 0;
}
