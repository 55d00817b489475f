// Bit-exact emulation of the x87 80-bit Euclidean norm used by the reference.
//
// numba lowers 1-D np.linalg.norm to BLAS dnrm2; the OpenBLAS x86-64 kernel
// computes sum(v_i^2) and the square root on the x87 stack (64-bit significand,
// round-to-nearest-even after every operation) and rounds to double only at the
// end (pinned by tests/golden/rounding_probes.npz).  Results
// that feed integer outputs (visibility, delay bins, direction indices) must match
// that bit for bit, so the GPU does the same arithmetic with integers.
//
// Plain C++ (also compiled for the host by tests/x87_selftest.cpp, where it is
// compared against the CPU's real `long double`).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define X87_HD __host__ __device__ __forceinline__
#else
#define X87_HD inline
#endif

namespace x87 {

typedef unsigned __int128 u128;

// non-negative extended value: mant * 2^exp with bit 63 of mant set, or mant == 0
struct Ext {
    uint64_t mant;
    int exp;
};

X87_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

X87_HD uint64_t double_bits(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    return b;
}

// |v| = m * 2^e with m < 2^53 (m == 0 for zero)
X87_HD void decompose(double v, uint64_t &m, int &e) {
    const uint64_t b = double_bits(v) & 0x7fffffffffffffffULL;
    const int ef = (int)(b >> 52);
    const uint64_t frac = b & 0x000fffffffffffffULL;
    if (ef == 0) { m = frac; e = -1074; }
    else { m = frac | 0x0010000000000000ULL; e = ef - 1075; }
}

// round a 128-bit magnitude x (scaled by 2^exp, `sticky` = lower bits already lost)
// to a normalised 64-bit significand, round-to-nearest-even
X87_HD Ext round_u128(u128 x, int exp, bool sticky) {
    Ext r;
    const uint64_t hi = (uint64_t)(x >> 64), lo = (uint64_t)x;
    if (hi == 0) {
        if (lo == 0) { r.mant = 0; r.exp = 0; return r; }
        const int s = clz64(lo);               // exact: fits in 64 bits
        r.mant = lo << s; r.exp = exp - s;
        return r;
    }
    const int drop = 64 - clz64(hi);            // 1..64 low bits to drop
    uint64_t keep = (uint64_t)(x >> drop);
    const u128 rem = x & ((((u128)1) << drop) - 1);
    const u128 half = ((u128)1) << (drop - 1);
    int e = exp + drop;
    bool up = rem > half || (rem == half && (sticky || (keep & 1)));
    if (up) {
        keep += 1;
        if (keep == 0) { keep = 0x8000000000000000ULL; e += 1; }
    }
    r.mant = keep; r.exp = e;
    return r;
}

// (long double)v * (long double)v
X87_HD Ext square(double v) {
    uint64_t m; int e;
    decompose(v, m, e);
    return round_u128((u128)m * (u128)m, 2 * e, false);
}

// a + b, both non-negative
X87_HD Ext add(Ext a, Ext b) {
    if (a.mant == 0) return b;
    if (b.mant == 0) return a;
    if (a.exp < b.exp) { Ext t = a; a = b; b = t; }
    const int d = a.exp - b.exp;
    if (d > 66) return a;                       // b < ulp(a)/4
    // work at scale 2^(a.exp - 64): A = a.mant << 64 would overflow when added, so
    // use scale 2^(a.exp - 63): A has 127 bits, sum < 2^128
    const u128 A = ((u128)a.mant) << 63;
    u128 B;
    bool sticky = false;
    const int sh = 63 - d;                      // b.mant << sh at this scale
    if (sh >= 0) B = ((u128)b.mant) << sh;
    else {
        B = ((u128)b.mant) >> (-sh);
        sticky = (b.mant & ((((uint64_t)1) << (-sh)) - 1)) != 0;
    }
    return round_u128(A + B, a.exp - 63, sticky);
}

X87_HD u128 sq64(uint64_t r) { return (u128)r * (u128)r; }

// sqrtl
X87_HD Ext sqrt_ext(Ext x) {
    if (x.mant == 0) return x;
    // X = mant * 2^k with (exp - k) even, X in [2^126, 2^128)
    const bool odd = (x.exp & 1) != 0;
    const int k = odd ? 63 : 64;
    const u128 X = ((u128)x.mant) << k;
    const int e = (x.exp - k) / 2;              // exact: even
    // double-precision first guess, one exact-residual correction, then fix-up
    const double xd = ldexp((double)x.mant, k - 64);     // X / 2^64 (rounded)
    const double guess = sqrt(xd) * 4294967296.0;        // * 2^32 ~ sqrt(X)
    uint64_t r;
    if (guess >= 18446744073709549568.0) r = 0xfffffffffffff800ULL;
    else if (guess < 9223372036854775808.0) r = 0x8000000000000000ULL;
    else r = (uint64_t)guess;
    {
        const u128 r2 = sq64(r);
        const bool pos = X >= r2;
        const u128 diff = pos ? X - r2 : r2 - X;
        const double delta =
            (double)(uint64_t)(diff >> 64) * 18446744073709551616.0 + (double)(uint64_t)diff;
        const double corr = floor(delta / (2.0 * (double)r));
        if (pos) {
            const uint64_t room = 0xffffffffffffffffULL - r;
            r += (corr >= (double)room) ? room : (uint64_t)corr;
        } else {
            const uint64_t room = r - 0x8000000000000000ULL;
            r -= (corr >= (double)room) ? room : (uint64_t)corr;
        }
    }
    while (sq64(r) > X) --r;
    while (r != 0xffffffffffffffffULL && sq64(r + 1) <= X) ++r;
    const u128 rem = X - sq64(r);
    Ext out;
    out.exp = e;
    if (rem > (u128)r) {                        // fraction > 1/2 (ties impossible)
        if (r == 0xffffffffffffffffULL) { out.mant = 0x8000000000000000ULL; out.exp = e + 1; }
        else out.mant = r + 1;
    } else out.mant = r;
    if (!(out.mant >> 63)) {                    // root of a 127-bit X has 64 bits only
        const int s = clz64(out.mant);          // when X >= 2^126; keep it general
        out.mant <<= s; out.exp -= s;
    }
    return out;
}

// (double)ext, round-to-nearest-even
X87_HD double to_double(Ext x) {
    if (x.mant == 0) return 0.0;
    uint64_t m = x.mant >> 11;
    const uint64_t rem = x.mant & 0x7ffULL;
    int e = x.exp + 11;
    if (rem > 0x400ULL || (rem == 0x400ULL && (m & 1))) {
        m += 1;
        if (m == (1ULL << 53)) { m >>= 1; e += 1; }
    }
    return ldexp((double)m, e);
}

X87_HD double norm3(double a, double b, double c) {
    return to_double(sqrt_ext(add(add(square(a), square(b)), square(c))));
}

X87_HD double norm2(double a, double b) {
    return to_double(sqrt_ext(add(square(a), square(b))));
}

}  // namespace x87
