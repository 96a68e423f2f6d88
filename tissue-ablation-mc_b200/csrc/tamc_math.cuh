// tamc_math.cuh -- fp64 elementary functions of the transport kernels, specialised to the arguments the
// photon loop produces.
//
// The production kernels are instruction-issue bound, and the launch / scattering arithmetic is mostly
// log, sin/cos and sqrt (sourceph.f90:28-35, inttau2.f90:36, stokes.f90:48-68).  The CUDA library versions
// handle every IEEE input (denormals, infinities, NaN, huge arguments) and carry each 64-bit polynomial
// coefficient as an immediate, which costs two extra issue slots per coefficient (UMOV lo / UMOV hi).  The
// functions below
//   * accept only the ranges the callers guarantee (stated per function), so they have no slow paths;
//   * keep their coefficients in __constant__ memory, which fp64 instructions read as a direct operand;
//   * are accurate to about 1 ulp (tests/test_math_accuracy.py measures them on the host against libm).
// Polynomial coefficients: the fdlibm kernels (__kernel_sin, __kernel_cos, __ieee754_log; Sun Microsystems,
// freely distributable) -- minimax fits valid on the reduced ranges used here.
//
// Compiled by plain g++ (no nvcc) the same source builds for the host with std::fma, so the accuracy test
// (tests/test_math_accuracy.py via tools/math_check.cpp) runs without a GPU; under nvcc the functions are device-only.
#pragma once

#include <cstdint>

#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define TAMC_HD __device__ __forceinline__
#define TAMC_TABLE static __constant__
#else
#define TAMC_HD inline
#define TAMC_TABLE static const
#endif

#if defined(__CUDA_ARCH__)
#define TAMC_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define TAMC_FMA(a, b, c) std::fma((a), (b), (c))
#endif

namespace tamc {
namespace fm {

// fdlibm __kernel_sin S1..S6, __kernel_cos C1..C6 (|x| <= pi/4)
TAMC_TABLE double kSinC[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                              2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
TAMC_TABLE double kCosC[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                              -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
// fdlibm __ieee754_log Lg1..Lg7, then ln2_hi, ln2_lo
TAMC_TABLE double kLogC[9] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                              2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                              1.479819860511658591e-01, 6.93147180369123816490e-01, 1.90821492927058770002e-10};
// pi = hi + lo
TAMC_TABLE double kPiC[2] = {3.14159265358979311600e+00, 1.22464679914735317723e-16};
#define TAMC_SIN(i) kSinC[i]
#define TAMC_COS(i) kCosC[i]
#define TAMC_LOG(i) kLogC[i]
#define TAMC_PI(i) kPiC[i]

TAMC_HD int hi_word(double v)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(v);
#else
    uint64_t b;
    memcpy(&b, &v, 8);
    return (int)(b >> 32);
#endif
}
TAMC_HD int lo_word(double v)
{
#if defined(__CUDA_ARCH__)
    return __double2loint(v);
#else
    uint64_t b;
    memcpy(&b, &v, 8);
    return (int)(uint32_t)b;
#endif
}
TAMC_HD double from_words(int hi, int lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}

// 1/d for a normal d well inside the exponent range (no zero, denormal, infinity): hardware seed (2^-23) and
// two Newton steps.
TAMC_HD double rcp_normal(double d)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
#else
    double r = (double)(1.0f / (float)d);
#endif
    double e = TAMC_FMA(-d, r, 1.0);
    r = TAMC_FMA(r, e, r);
    e = TAMC_FMA(-d, r, 1.0);
    return TAMC_FMA(r, e, r);
}

// sqrt(a) for a normal a > 0 far from the ends of the exponent range.  Seed 2^-23, two coupled Newton steps on
// (g ~ sqrt a, h ~ 1/(2 sqrt a)) and a final residual correction: correctly rounded except in rare half-way cases.
TAMC_HD double sqrt_normal(double a)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
#else
    double y = (double)(1.0f / std::sqrt((float)a));
#endif
    double g = a * y, h = 0.5 * y;
    double e = TAMC_FMA(-h, g, 0.5);
    g = TAMC_FMA(g, e, g);
    h = TAMC_FMA(h, e, h);
    e = TAMC_FMA(-h, g, 0.5);
    g = TAMC_FMA(g, e, g);
    h = TAMC_FMA(h, e, h);
    const double d = TAMC_FMA(-g, g, a);
    return TAMC_FMA(d, h, g);
}

// 1/sqrt(a) for a normal a > 0 far from the ends of the exponent range: hardware seed (2^-23) and two Newton steps
// (relative error ~1e-16; not correctly rounded).
TAMC_HD double rsqrt_normal(double a)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
#else
    double y = (double)(1.0f / std::sqrt((float)a));
#endif
    const double ha = 0.5 * a;
    double e = TAMC_FMA(-ha * y, y, 0.5);
    y = TAMC_FMA(y, e, y);
    e = TAMC_FMA(-ha * y, y, 0.5);
    return TAMC_FMA(y, e, y);
}

// sin(pi a), cos(pi a) for 0 <= a <= 2.5 (callers: a = angle/pi with the angle in [0, 2 pi]).
// q = nearest integer to 2a, t = a - q/2 exactly, |t| <= 1/4, x = pi t, fdlibm kernels on |x| <= pi/4, then the
// quadrant symmetries.
TAMC_HD void sincospi_0_2(double a, double *sn, double *cs)
{
    const double magic = 6755399441055744.0;               // 1.5 * 2^52: adding it rounds to an integer in the low word
    const double qm = TAMC_FMA(a, 2.0, magic);
    const int q = lo_word(qm);
    const double qd = qm - magic;
    const double t = TAMC_FMA(qd, -0.5, a);
    const double x = TAMC_FMA(t, TAMC_PI(0), t * TAMC_PI(1));
    const double z = x * x;
    double ps = TAMC_SIN(5);
    ps = TAMC_FMA(ps, z, TAMC_SIN(4));
    ps = TAMC_FMA(ps, z, TAMC_SIN(3));
    ps = TAMC_FMA(ps, z, TAMC_SIN(2));
    ps = TAMC_FMA(ps, z, TAMC_SIN(1));
    ps = TAMC_FMA(ps, z, TAMC_SIN(0));
    const double s = TAMC_FMA(x * z, ps, x);
    double pc = TAMC_COS(5);
    pc = TAMC_FMA(pc, z, TAMC_COS(4));
    pc = TAMC_FMA(pc, z, TAMC_COS(3));
    pc = TAMC_FMA(pc, z, TAMC_COS(2));
    pc = TAMC_FMA(pc, z, TAMC_COS(1));
    pc = TAMC_FMA(pc, z, TAMC_COS(0));
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;                               // fdlibm: cos = w + (((1-w)-hz) + z*z*pc)
    const double c = w + TAMC_FMA(z * z, pc, (1.0 - w) - hz);
    // sin(x + q pi/2), cos(x + q pi/2)
    const bool odd = (q & 1) != 0;
    double rs = odd ? c : s, rc = odd ? s : c;
    if (q & 2) rs = -rs;
    if ((q + 1) & 2) rc = -rc;
    *sn = rs;
    *cs = rc;
}

// -log((x + 0.5) * 2^-32) for a 32-bit x: the optical depth drawn from one Philox word (inttau2.f90:36 with the
// uniform of tamc_transport.cuh's u32_to_unit).  The argument v = x + 0.5 lies in [0.5, 2^32): normal, positive.
// fdlibm's __ieee754_log: v = 2^k m, m in [sqrt(1/2), sqrt 2), f = m - 1, s = f/(2+f),
// log m = f - hfsq + s (hfsq + R(s^2)); the result is -( (k-32) ln2 + log m ).
TAMC_HD double neglog_u32(uint32_t x)
{
    const double v = (double)x + 0.5;
    int hi = hi_word(v);
    int k = (hi >> 20) - 1023 - 32;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi >= 0x3ff6a09f) { hi -= 0x00100000; k += 1; }
    const double m = from_words(hi, lo_word(v));
    const double f = m - 1.0;
    const double d = 2.0 + f;
    const double r = rcp_normal(d);
    double s = f * r;
    s = TAMC_FMA(TAMC_FMA(-d, s, f), r, s);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * TAMC_FMA(w, TAMC_FMA(w, TAMC_LOG(5), TAMC_LOG(3)), TAMC_LOG(1));
    const double t2 = z * TAMC_FMA(w, TAMC_FMA(w, TAMC_FMA(w, TAMC_LOG(6), TAMC_LOG(4)), TAMC_LOG(2)), TAMC_LOG(0));
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    // log v' = dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f); return its negative
    return ((hfsq - TAMC_FMA(s, hfsq + R, dk * TAMC_LOG(8))) - f) - dk * TAMC_LOG(7);
}

}  // namespace fm
}  // namespace tamc
