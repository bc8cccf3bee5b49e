// glibc_trig.cuh -- sinf / cosf / atanf / atan2f with the exact results of the host libm that OpenCV's projector
// code calls (glibc 2.39, x86-64, the variants its ifunc resolvers pick on FMA + AVX2 CPUs), usable from device
// code and -- for the exhaustive CPU cross-check against libm in tests/ -- from host code.
//
// Why: cv::detail::{Spherical,Cylindrical,Stereographic}Projector::mapBackward evaluate sinf / cosf / atan2f / atanf
// through libm (reference call sites src/math/_projection.cpp:51,81,321).  CUDA's own sinf/cosf/atan2f are faithful
// but not identical: a last-ulp difference moves a sample across a 1/32-px bin of cv::remap's fixed-point
// sampler.  With the same arithmetic as libm the maps -- and therefore the warped tiles -- are bit-identical.
//
// The algorithms are the published ones glibc ships:
//   sinf / cosf : double-precision polynomial evaluation after a pi/2 range reduction (S. Nagy's "optimized routines",
//                 glibc sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf.h); the x86-64 build selected at run time
//                 on FMA CPUs contracts every  a + b*c  of that code into one fused multiply-add -- the contraction
//                 pattern below was read off the disassembly of libm.so.6 and is pinned by the exhaustive test.
//   atanf       : fdlibm's single-precision kernel (argument reduction to 4 intervals + odd/even polynomial), no FMA.
//   atan2f      : fdlibm's quadrant logic around atanf(|y/x|).
// Every operation is an individually rounded IEEE operation; device code uses the __f*_rn / __d*_rn intrinsics so that
// nvcc cannot contract or reorder anything, host code must be compiled with -ffp-contract=off.
#pragma once
#include <stdint.h>
#include <string.h>
#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

#if defined(__CUDACC__)
#define GT_HD __host__ __device__ __forceinline__
#else
#define GT_HD inline
#endif

namespace gtrig {

#if defined(__CUDA_ARCH__)
GT_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
GT_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
GT_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
GT_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
GT_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
GT_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
GT_HD float d2f(double a) { return __double2float_rn(a); }
GT_HD uint32_t f2u(float f) { return __float_as_uint(f); }
GT_HD float u2f(uint32_t u) { return __uint_as_float(u); }
GT_HD int32_t d2i_trunc(double a) { return __double2int_rz(a); }
#else
GT_HD float fmul(float a, float b) { return a * b; }
GT_HD float fadd(float a, float b) { return a + b; }
GT_HD float fsub(float a, float b) { return a - b; }
GT_HD float fdiv(float a, float b) { return a / b; }
GT_HD double dmul(double a, double b) { return a * b; }
GT_HD double dfma(double a, double b, double c) { return fma(a, b, c); }
GT_HD float d2f(double a) { return (float)a; }
GT_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
GT_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
GT_HD int32_t d2i_trunc(double a) { return (int32_t)a; }
#endif

// ---- sinf / cosf -------------------------------------------------------------------------------------------------
// polynomial coefficients of sin and cos on [-pi/4, pi/4]; set 1 is the negated cosine (quadrants 2, 3)
struct SinCosPoly { double c0, c1, c2, c3, c4, s1, s2, s3; };

GT_HD SinCosPoly sincos_poly(bool neg_cos)
{
    SinCosPoly p;
    const double sg = neg_cos ? -1.0 : 1.0;
    p.c0 = sg * 0x1p0;
    p.c1 = sg * -0x1.ffffffd0c621cp-2;
    p.c2 = sg * 0x1.55553e1068f19p-5;
    p.c3 = sg * -0x1.6c087e89a359dp-10;
    p.c4 = sg * 0x1.99343027bf8c3p-16;
    p.s1 = -0x1.555545995a603p-3;
    p.s2 = 0x1.1107605230bc4p-7;
    p.s3 = -0x1.994eb3774cf24p-13;
    return p;
}

// sin polynomial when n is even, cos polynomial when n is odd (x already reduced, x2 = x*x)
GT_HD float sinf_poly(double x, double x2, const SinCosPoly &p, int n)
{
    if ((n & 1) == 0) {
        const double x3 = dmul(x, x2);
        const double s1 = dfma(x2, p.s3, p.s2);
        const double x7 = dmul(x3, x2);
        const double s = dfma(x3, p.s1, x);
        return d2f(dfma(x7, s1, s));
    }
    const double x4 = dmul(x2, x2);
    const double c2 = dfma(x2, p.c4, p.c3);
    const double c1 = dfma(x2, p.c1, p.c0);
    const double x6 = dmul(x4, x2);
    const double c = dfma(x4, p.c2, c1);
    return d2f(dfma(x6, c2, c));
}

// |x| < 120: n = round(x * 2/pi) through a 2^24-scaled truncation, x - n * pi/2 as one fused operation
GT_HD double reduce_fast(double x, int *np)
{
    const double r = dmul(x, 0x1.45F306DC9C883p+23);
    const int n = (d2i_trunc(r) + 0x800000) >> 24;
    *np = n;
    return dfma(-(double)n, 0x1.921FB54442D18p0, x);
}

#define GT_INV_PIO4 {0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, \
                     0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295, \
                     0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041}
#if defined(__CUDACC__)
static __constant__ uint32_t inv_pio4_dev[24] = GT_INV_PIO4;
#endif
static const uint32_t inv_pio4_host[24] = GT_INV_PIO4;

// |x| >= 120: 96 bits of 2/pi selected by the exponent, fixed-point product, quadrant from the top two bits
GT_HD double reduce_large(uint32_t xi, int *np)
{
#if defined(__CUDA_ARCH__)
    const uint32_t *arr = &inv_pio4_dev[(xi >> 26) & 15];
#else
    const uint32_t *arr = &inv_pio4_host[(xi >> 26) & 15];
#endif
    const int shift = (xi >> 23) & 7;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    uint64_t res0 = (uint32_t)(xi * arr[0]);
    const uint64_t res1 = (uint64_t)xi * arr[4];
    const uint64_t res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const uint64_t n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    const double x = (double)(int64_t)res0;
    *np = (int)n;
    return dmul(x, 0x1.921FB54442D18p-62);
}

GT_HD uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ff; }

template <bool COS>
GT_HD float sincosf_impl(float y)
{
    double x = (double)y;
    int n;
    const uint32_t top = abstop12(y);
    if (top < 0x3f4) {                       // |y| < pi/4
        const double x2 = dmul(x, x);
        if (top < 0x398) return COS ? 1.0f : y;   // |y| < 2^-12
        return sinf_poly(x, x2, sincos_poly(false), COS ? 1 : 0);
    }
    if (top < 0x42f) {                       // |y| < 120
        x = reduce_fast(x, &n);
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign of sine in quadrants 0..3
        const SinCosPoly p = sincos_poly((n & 2) != 0);
        return sinf_poly(dmul(x, s), dmul(x, x), p, COS ? (n ^ 1) : n);
    }
    if (top < 0x7f8) {                       // finite
        const uint32_t xi = f2u(y);
        const int sign = (int)(xi >> 31);
        x = reduce_large(xi, &n);
        const int q = n + sign;
        const double s = ((q & 3) == 1 || (q & 3) == 2) ? -1.0 : 1.0;
        const SinCosPoly p = sincos_poly((q & 2) != 0);
        return sinf_poly(dmul(x, s), dmul(x, x), p, COS ? (n ^ 1) : n);
    }
    return fdiv(fsub(y, y), fsub(y, y));     // inf / NaN -> NaN
}

GT_HD float sinf_glibc(float y) { return sincosf_impl<false>(y); }
GT_HD float cosf_glibc(float y) { return sincosf_impl<true>(y); }

// ---- atanf (fdlibm, single precision) -----------------------------------------------------------------------------
GT_HD float atanf_glibc(float x)
{
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
                aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
                aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
    int id;
    if (ix >= 0x4c000000u) {                 // |x| >= 2^25
        if (ix > 0x7f800000u) return fadd(x, x);   // NaN
        const float r = fadd(atanhi[3], atanlo[3]);
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) {                  // |x| < 0.4375
        if (ix < 0x31000000u) return x;      // |x| < 2^-29
        id = -1;
    } else {
        x = u2f(ix);                         // fabsf
        if (ix < 0x3f980000u) {              // |x| < 1.1875
            if (ix < 0x3f300000u) { id = 0; x = fdiv(fsub(fmul(2.0f, x), 1.0f), fadd(2.0f, x)); }
            else { id = 1; x = fdiv(fsub(x, 1.0f), fadd(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000u) { id = 2; x = fdiv(fsub(x, 1.5f), fadd(1.0f, fmul(1.5f, x))); }
            else { id = 3; x = fdiv(-1.0f, x); }
        }
    }
    const float z = fmul(x, x);
    const float w = fmul(z, z);
    const float s1 = fmul(z, fadd(aT0, fmul(w, fadd(aT2, fmul(w, fadd(aT4, fmul(w, fadd(aT6, fmul(w, fadd(aT8, fmul(w, aT10)))))))))));
    const float s2 = fmul(w, fadd(aT1, fmul(w, fadd(aT3, fmul(w, fadd(aT5, fmul(w, fadd(aT7, fmul(w, aT9)))))))));
    if (id < 0) return fsub(x, fmul(x, fadd(s1, s2)));
    const float r = fsub(atanhi[id], fsub(fsub(fmul(x, fadd(s1, s2)), atanlo[id]), x));
    return (hx >> 31) ? -r : r;
}

// ---- atan2f (fdlibm) ----------------------------------------------------------------------------------------------
GT_HD float atan2f_glibc(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const uint32_t hx = f2u(x), hy = f2u(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return fadd(x, y);   // NaN
    if (hx == 0x3f800000u) return atanf_glibc(y);                   // x == 1
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);  // 2 sign(x) + sign(y)
    if (iy == 0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return fadd(pi, tiny);
        default: return fsub(-pi, tiny);
        }
    }
    if (ix == 0) return (hy >> 31) ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
            case 0: return fadd(pi_o_4, tiny);
            case 1: return fsub(-pi_o_4, tiny);
            case 2: return fadd(fmul(3.0f, pi_o_4), tiny);
            default: return fsub(fmul(-3.0f, pi_o_4), tiny);
            }
        }
        switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return fadd(pi, tiny);
        default: return fsub(-pi, tiny);
        }
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = fadd(pi_o_2, fmul(0.5f, pi_lo));
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = atanf_glibc(u2f(f2u(fdiv(y, x)) & 0x7fffffffu));
    switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return fsub(pi, fsub(z, pi_lo));
    default: return fsub(fsub(z, pi_lo), pi);
    }
}

} // namespace gtrig
