// Portable, bit-reproducible fp32 exp / log for the CRF decode arithmetic.
//
// The CRF posteriors feed an arg-max (max-marginal Viterbi, bonito/crf/model.py:92-95,215-218),
// so "bit-exact decodes from identical fp32 scores" needs transcendental functions that return the
// same bits in the CUDA kernels and in the CPU checker (oracle/c/crf_exact.c).  libm's expf/logf and
// CUDA's expf/logf differ in the last ulp, so both sides use the functions below instead: only IEEE
// round-to-nearest add / mul / fma and integer bit operations, written through XB_ADD / XB_MUL /
// XB_FMA so that no compiler contracts or reassociates them (device: __fadd_rn / __fmul_rn /
// __fmaf_rn are never fused; host: compile with -ffp-contract=off).
//
// Accuracy (tests/test_exact_math.py): <= 2 ulp against double-precision exp / log on the ranges used.
#ifndef XB_EXACT_MATH_H
#define XB_EXACT_MATH_H

#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define XB_HD __device__ __forceinline__
#define XB_ADD(a, b) __fadd_rn((a), (b))
#define XB_SUB(a, b) __fsub_rn((a), (b))
#define XB_MUL(a, b) __fmul_rn((a), (b))
#define XB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define XB_RCP(a) __frcp_rn((a))
#define XB_F2U(x) __float_as_uint((x))
#define XB_U2F(x) __uint_as_float((x))
#else
#include <string.h>
#if defined(__CUDACC__)
#define XB_HD __host__ __device__ inline
#else
#define XB_HD static inline
#endif
#define XB_ADD(a, b) ((a) + (b))
#define XB_SUB(a, b) ((a) - (b))
#define XB_MUL(a, b) ((a) * (b))
#define XB_FMA(a, b, c) __builtin_fmaf((a), (b), (c))
#define XB_RCP(a) (1.0f / (a))
static inline uint32_t xb_f2u_(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float xb_u2f_(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
#define XB_F2U(x) xb_f2u_((x))
#define XB_U2F(x) xb_u2f_((x))
#endif

#define XB_NEG_BIG (-1e38f)      /* seqdist semiring zero */
#define XB_POST_EPS (1e-8f)      /* bonito/crf/model.py:216 */

// exp(x).  Returns 0 for x < -86 (covers -1e38 and -inf; avoids subnormal results) and +inf for
// x > 88.  Range reduction x = n ln2 + r with the round-to-nearest magic constant, degree-6
// polynomial (Cephes expf coefficients) on |r| <= ln2/2, exponent insertion by integer add.
XB_HD float xb_expf(float x) {
    if (!(x >= -86.0f)) return 0.0f;
    if (x > 88.0f) return XB_U2F(0x7f800000u);
    const float magic = 12582912.0f;                       // 1.5 * 2^23
    float t = XB_FMA(x, 1.44269504088896341f, magic);
    float n = XB_SUB(t, magic);
    float r = XB_FMA(n, -0.693359375f, x);
    r = XB_FMA(n, 2.12194440e-4f, r);
    float z = XB_MUL(r, r);
    float p = 1.9875691500e-4f;
    p = XB_FMA(p, r, 1.3981999507e-3f);
    p = XB_FMA(p, r, 8.3334519073e-3f);
    p = XB_FMA(p, r, 4.1665795894e-2f);
    p = XB_FMA(p, r, 1.6666665459e-1f);
    p = XB_FMA(p, r, 5.0000001201e-1f);
    p = XB_FMA(p, z, r);
    p = XB_ADD(p, 1.0f);
    int32_t ni = (int32_t)n;
    return XB_U2F(XB_F2U(p) + ((uint32_t)ni << 23));
}

// log(x) for finite x > 0 (subnormals handled); x == 0 -> -inf, x < 0 or NaN -> NaN, +inf -> +inf.
// Mantissa folded into [sqrt(1/2), sqrt(2)), degree-9 polynomial (Cephes logf coefficients).
XB_HD float xb_logf(float x) {
    uint32_t ix = XB_F2U(x);
    int32_t eadj = 0;
    if (ix >= 0x7f800000u || ix < 0x00800000u) {
        if ((ix << 1) == 0) return XB_U2F(0xff800000u);            // +-0 -> -inf
        if (ix == 0x7f800000u) return x;                           // +inf
        if (ix > 0x7f800000u) return XB_U2F(0x7fc00000u);          // negative or NaN
        x = XB_MUL(x, 8388608.0f);                                 // subnormal: scale by 2^23
        ix = XB_F2U(x);
        eadj = -23;
    }
    uint32_t iy = ix - 0x3f3504f3u;
    int32_t e = ((int32_t)iy >> 23) + eadj;
    float m = XB_U2F((iy & 0x007fffffu) + 0x3f3504f3u);
    float f = XB_SUB(m, 1.0f);
    float z = XB_MUL(f, f);
    float p = 7.0376836292e-2f;
    p = XB_FMA(p, f, -1.1514610310e-1f);
    p = XB_FMA(p, f, 1.1676998740e-1f);
    p = XB_FMA(p, f, -1.2420140846e-1f);
    p = XB_FMA(p, f, 1.4249322787e-1f);
    p = XB_FMA(p, f, -1.6668057665e-1f);
    p = XB_FMA(p, f, 2.0000714765e-1f);
    p = XB_FMA(p, f, -2.4999993993e-1f);
    p = XB_FMA(p, f, 3.3333331174e-1f);
    p = XB_MUL(XB_MUL(p, f), z);
    float fe = (float)e;
    p = XB_FMA(fe, -2.12194440e-4f, p);
    p = XB_FMA(z, -0.5f, p);
    float r = XB_ADD(f, p);
    return XB_FMA(fe, 0.693359375f, r);
}

// Specialisations for the arguments the decode sweeps actually produce; same bits as the general functions on
// their domain (tests/test_cpu_oracle.py::test_exact_math_fast_variants), fewer instructions:
//   xb_expf_le0: x <= 0 (a score minus a running maximum), so no overflow branch;
//   xb_logf_norm: x a positive normal number (a sum of exponentials >= 1, or a posterior + 1e-8).
XB_HD float xb_expf_le0(float x) {
    const float magic = 12582912.0f;
    float t = XB_FMA(x, 1.44269504088896341f, magic);
    float n = XB_SUB(t, magic);
    float r = XB_FMA(n, -0.693359375f, x);
    r = XB_FMA(n, 2.12194440e-4f, r);
    float z = XB_MUL(r, r);
    float p = 1.9875691500e-4f;
    p = XB_FMA(p, r, 1.3981999507e-3f);
    p = XB_FMA(p, r, 8.3334519073e-3f);
    p = XB_FMA(p, r, 4.1665795894e-2f);
    p = XB_FMA(p, r, 1.6666665459e-1f);
    p = XB_FMA(p, r, 5.0000001201e-1f);
    p = XB_FMA(p, z, r);
    p = XB_ADD(p, 1.0f);
    // the exponent comes from the integer part of t (= magic + n: its low mantissa bits hold n in two's complement),
    // and the underflow case is a select at the end instead of an early return: no divergent branch per call
    uint32_t ni = XB_F2U(t) - 0x4b400000u;
    float v = XB_U2F(XB_F2U(p) + (ni << 23));
    return (x >= -86.0f) ? v : 0.0f;
}

// xb_expf on the clamped score domain [-80, 80] of the linear-domain decode: no overflow / underflow guards at all.
XB_HD float xb_expf_mid(float x) {
    const float magic = 12582912.0f;
    float t = XB_FMA(x, 1.44269504088896341f, magic);
    float n = XB_SUB(t, magic);
    float r = XB_FMA(n, -0.693359375f, x);
    r = XB_FMA(n, 2.12194440e-4f, r);
    float z = XB_MUL(r, r);
    float p = 1.9875691500e-4f;
    p = XB_FMA(p, r, 1.3981999507e-3f);
    p = XB_FMA(p, r, 8.3334519073e-3f);
    p = XB_FMA(p, r, 4.1665795894e-2f);
    p = XB_FMA(p, r, 1.6666665459e-1f);
    p = XB_FMA(p, r, 5.0000001201e-1f);
    p = XB_FMA(p, z, r);
    p = XB_ADD(p, 1.0f);
    uint32_t ni = XB_F2U(t) - 0x4b400000u;
    return XB_U2F(XB_F2U(p) + (ni << 23));
}
// exp of a CRF score as the linear-domain decode defines it: clamp to [-80, 80], then xb_expf_mid
XB_HD float xb_score_exp(float m) {
    m = m < -80.0f ? -80.0f : (m > 80.0f ? 80.0f : m);
    return xb_expf_mid(m);
}
// power-of-two rescaling factor of a non-negative state vector with maximum mx: 2^(127 - biased exponent), so that
// mx * scale lies in [1, 2); 1 for zero / subnormal / inf / nan.  Exact (no rounding) on both sides.
XB_HD float xb_pow2_scale(float mx) {
    uint32_t b = XB_F2U(mx);
    uint32_t e = (b >> 23) & 0xffu;
    if (e == 0u || e == 255u || (b >> 31)) return 1.0f;
    return XB_U2F((254u - e) << 23);
}

XB_HD float xb_logf_norm(float x) {
    uint32_t iy = XB_F2U(x) - 0x3f3504f3u;
    int32_t e = (int32_t)iy >> 23;
    float m = XB_U2F((iy & 0x007fffffu) + 0x3f3504f3u);
    float f = XB_SUB(m, 1.0f);
    float z = XB_MUL(f, f);
    float p = 7.0376836292e-2f;
    p = XB_FMA(p, f, -1.1514610310e-1f);
    p = XB_FMA(p, f, 1.1676998740e-1f);
    p = XB_FMA(p, f, -1.2420140846e-1f);
    p = XB_FMA(p, f, 1.4249322787e-1f);
    p = XB_FMA(p, f, -1.6668057665e-1f);
    p = XB_FMA(p, f, 2.0000714765e-1f);
    p = XB_FMA(p, f, -2.4999993993e-1f);
    p = XB_FMA(p, f, 3.3333331174e-1f);
    p = XB_MUL(XB_MUL(p, f), z);
    float fe = (float)e;
    p = XB_FMA(fe, -2.12194440e-4f, p);
    p = XB_FMA(z, -0.5f, p);
    float r = XB_ADD(f, p);
    return XB_FMA(fe, 0.693359375f, r);
}

#endif  // XB_EXACT_MATH_H
