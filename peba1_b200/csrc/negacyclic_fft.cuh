// negacyclic_fft.cuh — folded negacyclic FP64 transform of a degree-<1024 real polynomial, written for a
// TEAM of 64 threads holding 8 complex points each (512 = 8 x 8 x 8).
//
// Replaces TFHE's FFT_Processor_nayuki / IntPolynomial_ifft / TorusPolynomial_fft (SURVEY.md §2.2 T3/T4,
// §8c step 4) — NOT a port: the twist by zeta^n is folded into the first radix-8 pass, the three passes run in
// registers, and the two exchanges go through (bank-conflict-free) shared memory; outputs stay in the
// digit-reversed order the pointwise MAC and the inverse transform consume, so no reordering pass exists.
//
// Math.  A real polynomial p (deg < N = 1024) is evaluated at zeta^(4k+1), zeta = exp(i*pi/N), k < 512:
//     F_k = sum_{n<512} z_n * w^(n k),   z_n = (p_n + i p_{n+512}) zeta^n,  w = zeta^4 = exp(2 pi i / 512)
// With n = 64a + m, m = 8b + c (thread id of pass 1), k = k0 + 8 k1 + 64 k2:
//     pass 1 (thread m):        Y[k0][m]     = zeta^(m(1+4k0)) * sum_a (x_a w32^a) w8^(a k0)
//     pass 2 (thread 8k0+c):    B[k0][k1][c] = w64^(c k1)      * sum_b Y[k0][8b+c] w8^(b k1)
//     pass 3 (thread 8k0+k1):   F[k0+8k1+64k2] =                 sum_c B[k0][k1][c] w8^(c k2)
// The inverse runs the transposed network with conjugated twiddles and is UNSCALED (the 1/512 is folded into
// the transformed bootstrapping key).
//
// Everything here is __host__ __device__ so that csrc/emu_team.cu can run the very same code with CPU threads.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define PEBA1_HD __host__ __device__ __forceinline__

namespace peba1 {

constexpr int kN = 1024;        // ring degree (both upstream parameter sets)
constexpr int kM = 512;         // complex points
constexpr int kTeam = 64;       // threads per transform

struct TeamTwiddles {
    double2 tw1[8];  // zeta^(tid (1+4j)), j = 0..7
    double2 tw2[8];  // w64^((tid&7) j),   j = 0..7   (tw2[0] = 1, unused)
};

// Shared memory owned by one team.
struct __align__(16) TeamSmem {
    double2  ex1[2][kM];    // exchange 1 (across the team), double buffered by transform parity
    double2  ex2[kM];       // exchange 2 (inside groups of 8 lanes)
    int32_t  acc[2][kN];    // TLWE accumulator: [0] mask, [1] body
    uint16_t bara[640];     // mod-switched input (n+1 <= 640 entries in [0, 2N))
};

PEBA1_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
PEBA1_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
PEBA1_HD double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
PEBA1_HD double2 cmul_conj(double2 a, double2 b) {  // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));
}
PEBA1_HD void cfma(double2& acc, double2 a, double2 b) {  // acc += a*b
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

// exp(i pi a / 16), a = 0..7  (= zeta^(64 a)).  On the device the values sit in the constant bank so that DMUL/DFMA
// take them as c[bank][offset] operands; as literals ptxas rebuilds each 64-bit immediate with two moves per use
// (4.6 % of all executed instructions in ncu r1d).
#ifdef __CUDACC__
static __constant__ double2 kW32c[8] = {
    {1.0, 0.0},
    {0.98078528040323044913, 0.19509032201612826785},
    {0.92387953251128675613, 0.38268343236508977173},
    {0.83146961230254523708, 0.55557023301960222474},
    {0.70710678118654752440, 0.70710678118654752440},
    {0.55557023301960222474, 0.83146961230254523708},
    {0.38268343236508977173, 0.92387953251128675613},
    {0.19509032201612826785, 0.98078528040323044913}};
#endif
PEBA1_HD double2 w32(int a) {
#ifdef __CUDA_ARCH__
    return kW32c[a];
#else
    switch (a) {
        case 1: return make_double2(0.98078528040323044913, 0.19509032201612826785);
        case 2: return make_double2(0.92387953251128675613, 0.38268343236508977173);
        case 3: return make_double2(0.83146961230254523708, 0.55557023301960222474);
        case 4: return make_double2(0.70710678118654752440, 0.70710678118654752440);
        case 5: return make_double2(0.55557023301960222474, 0.83146961230254523708);
        case 6: return make_double2(0.38268343236508977173, 0.92387953251128675613);
        case 7: return make_double2(0.19509032201612826785, 0.98078528040323044913);
        default: return make_double2(1.0, 0.0);
    }
#endif
}

// X_k = sum_a x_a w^(a k), w = exp(+2 pi i/8) (INV: exp(-2 pi i/8)); in place, natural order in and out.
// The odd half needs v1 * w and v3 * w^3 (w = (1 +- i)/sqrt2).  The factor 1/sqrt2 is not applied to v1 and v3 but
// folded into the last butterfly as an FMA: x1,5 = p +- h s, x3,7 = q +- i h d with s, d built from v1 +- v3 by
// additions only.  16 FP64 instructions for that half instead of 20 (52 per radix-8 instead of 56), one rounding less.
template <bool INV>
PEBA1_HD void radix8(double2 (&x)[8]) {
#ifdef __CUDA_ARCH__
    const double h = kW32c[4].x;
#else
    constexpr double h = 0.70710678118654752440;
#endif
    double2 u0 = cadd(x[0], x[4]), u1 = cadd(x[1], x[5]), u2 = cadd(x[2], x[6]), u3 = cadd(x[3], x[7]);
    double2 v0 = csub(x[0], x[4]), v1 = csub(x[1], x[5]), v2 = csub(x[2], x[6]), v3 = csub(x[3], x[7]);
    // even half: a 4-point DFT of u with w4 = +-i
    {
        double2 p = cadd(u0, u2), q = csub(u0, u2), r = cadd(u1, u3), t = csub(u1, u3);
        double2 it = INV ? make_double2(t.y, -t.x) : make_double2(-t.y, t.x);
        x[0] = cadd(p, r); x[4] = csub(p, r); x[2] = cadd(q, it); x[6] = csub(q, it);
    }
    // odd half
    {
        v2 = INV ? make_double2(v2.y, -v2.x) : make_double2(-v2.y, v2.x);  // * +-i
        const double2 p = cadd(v0, v2), q = csub(v0, v2);
        const double2 m = csub(v1, v3), n = cadd(v1, v3);
        if (!INV) {
            // v1 (1+i) + v3 (-1+i) = (m.x - n.y, n.x + m.y) =: s;  v1 (1+i) - v3 (-1+i) = (n.x - m.y, m.x + n.y) =: d
            const double2 sd = make_double2(m.x - n.y, n.x + m.y), dd = make_double2(n.x - m.y, m.x + n.y);
            x[1] = make_double2(fma(h, sd.x, p.x), fma(h, sd.y, p.y));
            x[5] = make_double2(fma(-h, sd.x, p.x), fma(-h, sd.y, p.y));
            x[3] = make_double2(fma(-h, dd.y, q.x), fma(h, dd.x, q.y));   // q + i h d
            x[7] = make_double2(fma(h, dd.y, q.x), fma(-h, dd.x, q.y));
        } else {
            // v1 (1-i) + v3 (-1-i) = (m.x + n.y, m.y - n.x) =: s;  v1 (1-i) - v3 (-1-i) = (n.x + m.y, n.y - m.x) =: d
            const double2 sd = make_double2(m.x + n.y, m.y - n.x), dd = make_double2(n.x + m.y, n.y - m.x);
            x[1] = make_double2(fma(h, sd.x, p.x), fma(h, sd.y, p.y));
            x[5] = make_double2(fma(-h, sd.x, p.x), fma(-h, sd.y, p.y));
            x[3] = make_double2(fma(h, dd.y, q.x), fma(-h, dd.x, q.y));   // q - i h d
            x[7] = make_double2(fma(-h, dd.y, q.x), fma(h, dd.x, q.y));
        }
    }
}

// Forward transform.  In: x[a] = p[64a+tid] + i p[64a+tid+512] (thread tid of the team).
// Out: x[k2] = F[k0 + 8 k1 + 64 k2] with k0 = tid>>3, k1 = tid&7.
template <class Env, class SM>
PEBA1_HD void team_fft_forward(Env& env, double2 (&x)[8], const TeamTwiddles& tw, SM& sm, int parity) {
    const int tid = env.tid;
#pragma unroll
    for (int a = 1; a < 8; ++a) x[a] = cmul(x[a], w32(a));
    radix8<false>(x);
    double2* e1 = sm.ex1[parity];
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0) e1[64 * k0 + tid] = cmul(x[k0], tw.tw1[k0]);
    env.team_sync();
    const int hi = tid >> 3, lo = tid & 7;
#pragma unroll
    for (int b = 0; b < 8; ++b) x[b] = e1[64 * hi + 8 * b + lo];
    radix8<false>(x);
    sm.ex2[64 * hi + lo] = x[0];
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) sm.ex2[64 * hi + 8 * k1 + (lo ^ k1)] = cmul(x[k1], tw.tw2[k1]);
    env.warp_sync();
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = sm.ex2[64 * hi + 8 * lo + (c ^ lo)];
    radix8<false>(x);
}

// Inverse transform (unscaled).  In: x[k2] = G[k0 + 8 k1 + 64 k2], k0 = tid>>3, k1 = tid&7.
// Out: x[a] = g[64a+tid] * zeta^-(64a+tid), i.e. real part -> coefficient 64a+tid, imaginary -> +512.
template <class Env, class SM>
PEBA1_HD void team_fft_inverse(Env& env, double2 (&x)[8], const TeamTwiddles& tw, SM& sm, int parity) {
    const int tid = env.tid;
    const int hi = tid >> 3, lo = tid & 7;
    radix8<true>(x);
    env.warp_sync();  // previous readers of ex2 (last forward transform) are done
    sm.ex2[64 * hi + 8 * lo + lo] = x[0];
#pragma unroll
    for (int c = 1; c < 8; ++c) sm.ex2[64 * hi + 8 * lo + (c ^ lo)] = cmul_conj(x[c], tw.tw2[c]);
    env.warp_sync();
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) x[k1] = sm.ex2[64 * hi + 8 * k1 + (lo ^ k1)];
    radix8<true>(x);
    double2* e1 = sm.ex1[parity];
#pragma unroll
    for (int b = 0; b < 8; ++b) e1[64 * hi + 8 * b + lo] = x[b];
    env.team_sync();
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0) x[k0] = cmul_conj(e1[64 * k0 + tid], tw.tw1[k0]);
    radix8<true>(x);
#pragma unroll
    for (int a = 1; a < 8; ++a) x[a] = cmul_conj(x[a], w32(a));
}

// Index of transform bin (tid, k2) of polynomial (i, r, c) in the transformed bootstrapping key.
// Layout [i][r][k2][c][tid] of double2: a warp reads 512 contiguous bytes per (k2, c).
PEBA1_HD size_t bkfft_index(int rows, int i, int r, int k2, int c, int tid) {
    return ((((size_t)i * rows + r) * 8 + k2) * 2 + c) * 64 + tid;
}

// exact int32 -> double without the (quarter-rate) I2F.F64 conversion
PEBA1_HD double i2d(int32_t v) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)) - 4503601774854144.0;
#else
    return (double)v;
#endif
}
// gadget digit as a double: field in [0, 2*half) -> (double)(field - half), straight from the bit field: 2^52 + field
// is the double with high word 0x43300000 and low word `field`, so one subtraction of (2^52 + half) replaces the
// integer subtraction, the sign flip and the conversion (2 integer instructions per point and digit level less)
PEBA1_HD double digit2d(uint32_t field, uint32_t half) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(0x43300000, (int)field) - (4503599627370496.0 + (double)half);
#else
    return (double)((int32_t)field - (int32_t)half);
#endif
}
// round to nearest integer, returned mod 2^32 (|v| < 2^51)
PEBA1_HD uint32_t d2u_rn(double v) {
    const double m = v + 6755399441055744.0;  // 1.5 * 2^52
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(m);
#else
    uint64_t bits;
    memcpy(&bits, &m, sizeof bits);
    return (uint32_t)bits;
#endif
}

}  // namespace peba1
