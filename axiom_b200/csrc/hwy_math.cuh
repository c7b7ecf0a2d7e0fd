// hwy_math.cuh — fp32 transcendental functions restated from the PUBLISHED Google Highway algorithms
// (hwy/contrib/math/math-inl.h, the version pinned by the reference: third_party/highway @ da7098a)
// and from the reference's own erf/cbrt (src/backends/cpu/simd/hwy_math_ext-inl.h:1-89).
//
// Why: the reference CPU backend evaluates exp/log/sin/cos/tan/tanh/erf/gelu/sigmoid/silu/pow through
// these polynomials (src/backends/cpu/simd/hwy_kernels-inl.h:420-530,786-880), NOT through libm, for
// every full SIMD vector of a contiguous fp32 tensor. Re-stating the same sequence of IEEE fp32
// mul/add/fma operations on the GPU makes those results BIT-IDENTICAL to the reference's vector body
// (its <16-element scalar tails use libm and may differ by <=2 ulp).
//
// Every arithmetic step uses explicit round-to-nearest intrinsics so nvcc cannot contract or reorder:
//   hn::Mul -> __fmul_rn, hn::Add/Sub -> __fadd_rn/__fsub_rn, hn::MulAdd(a,b,c)=a*b+c -> __fmaf_rn,
//   hn::NegMulAdd(a,b,c)=c-a*b -> __fmaf_rn(-a,b,c), hn::MulSub(a,b,c)=a*b-c -> __fmaf_rn(a,b,-c),
//   hn::Div -> __fdiv_rn, hn::Sqrt -> __fsqrt_rn, ConvertTo<int32>(float) -> __float2int_rz (saturating).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace hwym {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float f_from_bits(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t bits(float f) { return __float_as_uint(f); }

// math-inl.h: Estrin(x, c0..c3) / (x, c0..c5)
__device__ __forceinline__ float estrin4(float x, float c0, float c1, float c2, float c3) {
    float x2 = mul(x, x);
    return fma_(x2, fma_(c3, x, c2), fma_(c1, x, c0));
}
__device__ __forceinline__ float estrin6(float x, float c0, float c1, float c2, float c3, float c4, float c5) {
    float x2 = mul(x, x);
    float x4 = mul(x2, x2);
    return fma_(x4, fma_(c5, x, c4), fma_(x2, fma_(c3, x, c2), fma_(c1, x, c0)));
}

// ---- ExpImpl<float> (math-inl.h:818-877) ------------------------------------------------------------
__device__ __forceinline__ float exp_poly(float x) {
    return fma_(estrin6(x, +0.5f, +0.166666671633720397949219f, +0.0416664853692054748535156f,
                        +0.00833336077630519866943359f, +0.00139304355252534151077271f,
                        +0.000198527617612853646278381f),
                mul(x, x), x);
}
__device__ __forceinline__ float pow2i(int32_t x) { return f_from_bits((uint32_t)(x + 0x7F) << 23); }
__device__ __forceinline__ float load_exp_short_range(float x, int32_t e) {
    int32_t y = e >> 1;
    return mul(mul(x, pow2i(y)), pow2i(e - y));
}
__device__ __forceinline__ float exp_reduce(float x, int32_t q) {
    float qf = __int2float_rn(q);
    x = fma_(qf, -0.693145751953125f, x);
    x = fma_(qf, -1.428606765330187045e-6f, x);
    return x;
}
// Exp (math-inl.h:1465-1485)
__device__ __forceinline__ float Exp(float x) {
    const float kOneOverLog2 = 1.442695040888963407359924681f;
    float half_signed = f_from_bits(bits(0.5f) | (bits(x) & 0x80000000u));
    int32_t q = __float2int_rz(fma_(x, kOneOverLog2, half_signed));
    float y = load_exp_short_range(add(exp_poly(exp_reduce(x, q)), 1.0f), q);
    return (x >= -104.0f) ? y : 0.0f;
}
// Expm1 (math-inl.h:1507-1531)
__device__ __forceinline__ float Expm1(float x) {
    const float kOneOverLog2 = 1.442695040888963407359924681f;
    float half_signed = f_from_bits(bits(0.5f) | (bits(x) & 0x80000000u));
    int32_t q = __float2int_rz(fma_(x, kOneOverLog2, half_signed));
    float y = exp_poly(exp_reduce(x, q));
    float z = (fabsf(x) < 0.346573590279972654708616f) ? y : sub(load_exp_short_range(add(y, 1.0f), q), 1.0f);
    return (x < -104.0f) ? -1.0f : z;
}

// ---- Log (math-inl.h:1002-1063, LogImpl<float> :880-902) -------------------------------------------
__device__ __forceinline__ float Log(float x) {
    const float kLn2Hi = 0.69313812256f, kLn2Lo = 9.0580006145e-6f;
    const float kMinNormal = 1.175494351e-38f, kScale = 3.355443200e+7f;
    const int32_t kMagic = 0x3F3504F3, kExpMask = 0x3F800000, kManMask = 0x7FFFFF;
    bool is_denormal = x < kMinNormal;
    x = is_denormal ? mul(x, kScale) : x;
    int32_t exp_bits = (int32_t)bits(x) + (kExpMask - kMagic);
    int32_t exp_scale = is_denormal ? -25 : 0;
    int32_t l2 = (int32_t)((uint32_t)exp_bits >> 23) - 0x7F;
    float e = __int2float_rn(exp_scale + l2);
    float y = f_from_bits((uint32_t)((exp_bits & kManMask) + kMagic));
    float ym1 = sub(y, 1.0f);
    float z = __fdiv_rn(ym1, add(y, 1.0f));
    // LogPoly
    float z2 = mul(z, z), z4 = mul(z2, z2);
    float poly = fma_(fma_(0.28498786688f, z4, 0.66666662693f), z2, mul(fma_(0.24279078841f, z4, 0.40000972152f), z4));
    // MulSub(exp, kLn2Hi, Sub(MulSub(z, Sub(ym1, poly), Mul(exp, kLn2Lo)), ym1))
    float inner = fma_(z, sub(ym1, poly), -mul(e, kLn2Lo));
    return fma_(e, kLn2Hi, -sub(inner, ym1));
}

// ---- Sin / Cos (math-inl.h:1442-1463,1563-1585; CosSinImpl<float> :671-737) -------------------------
__device__ __forceinline__ float sincos_poly(float x) {
    float y = mul(x, x);
    return fma_(estrin4(y, -1.66666597127914428710938e-1f, +8.33307858556509017944336e-3f,
                        -1.981069071916863322258e-4f, +2.6083159809786593541503e-6f),
                mul(y, x), x);
}
__device__ __forceinline__ float Cos(float x) {
    const float kOneOverPi = 0.31830988618379067153f;
    float y = fabsf(x);
    int32_t q = (__float2int_rz(mul(y, kOneOverPi)) << 1) + 1;
    float qf = __int2float_rn(q);
    y = fma_(qf, -0.5f * 3.140625f, y);
    y = fma_(qf, -0.5f * 0.0009670257568359375f, y);
    y = fma_(qf, -0.5f * 6.2771141529083251953e-7f, y);
    y = fma_(qf, -0.5f * 1.2154201256553420762e-10f, y);
    uint32_t sign = ((uint32_t)(~q & 2)) << 30; // (q & 2) == 0 ? -0.0 : +0.0
    return sincos_poly(f_from_bits(bits(y) ^ sign));
}
__device__ __forceinline__ float Sin(float x) {
    const float kOneOverPi = 0.31830988618379067153f;
    float ax = fabsf(x);
    uint32_t sign_x = bits(ax) ^ bits(x);
    int32_t q = __float2int_rz(fma_(ax, kOneOverPi, 0.5f));
    float qf = __int2float_rn(q);
    float y = ax;
    y = fma_(qf, -3.140625f, y);
    y = fma_(qf, -0.0009670257568359375f, y);
    y = fma_(qf, -6.2771141529083251953e-7f, y);
    y = fma_(qf, -1.2154201256553420762e-10f, y);
    uint32_t sign = (((uint32_t)q & 1u) << 31) ^ sign_x;
    return sincos_poly(f_from_bits(bits(y) ^ sign));
}
// reference UnaryTanImplT: sin/cos (hwy_kernels-inl.h:500-517)
__device__ __forceinline__ float Tan(float x) { return __fdiv_rn(Sin(x), Cos(x)); }

// ---- Tanh (math-inl.h:1601-1613) ---------------------------------------------------------------------
__device__ __forceinline__ float Tanh(float x) {
    uint32_t sign = bits(x) & 0x80000000u;
    float ax = f_from_bits(bits(x) ^ sign);
    float y = Expm1(mul(ax, 2.0f));
    float z = (ax > 18.714973875f) ? 1.0f : __fdiv_rn(y, add(y, 2.0f));
    return f_from_bits(bits(z) ^ sign);
}

// Exp(x) as Erf below needs it: x = -(s * s) is never positive and the result e only enters 1 - poly * e with
// 0 <= poly <= 1.1. Same value as Exp(x) wherever it can matter, in 10 instructions fewer:
//  * x <= -0 always, so Exp's half_signed is the constant -0.5;
//  * x is clamped at -86: there e <= 4.5e-38, poly * e < 2^-25 and 1 - poly * e rounds to exactly 1 for the clamped and
//    for the true argument alike (Exp's own cut-off to 0 below -104 included); NaN and -inf still end as they did — NaN
//    through poly, -inf through poly = 0;
//  * inside [-86, 0] q >= -124, so both power-of-two factors of load_exp_short_range and both products are normal and
//    exact: y * 2^y1 * 2^y2 == y * 2^q, which is one integer add on y's exponent field.
// tools/erf_equiv.cu compares Erf built on this with Erf built on Exp for all 2^32 inputs.
__device__ __forceinline__ float ExpNegForErf(float x) {
    const float kOneOverLog2 = 1.442695040888963407359924681f;
    x = fmaxf(x, -86.0f);
    int32_t q = __float2int_rz(fma_(x, kOneOverLog2, -0.5f));
    float y = add(exp_poly(exp_reduce(x, q)), 1.0f);
    return f_from_bits(bits(y) + ((uint32_t)q << 23));
}

// ---- reference Erf: Abramowitz-Stegun 7.1.26 (hwy_math_ext-inl.h:1-37) --------------------------------
__device__ __forceinline__ float Erf(float x) {
    const float p = 0.3275911f, a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f,
                a5 = 1.061405429f;
    float sign = copysignf(1.0f, x);
    float ax = fabsf(x);
    float t = __frcp_rn(fma_(p, ax, 1.0f)); // == Div(1, d): both are the correctly rounded 1 / d; rcp.rn is the shorter sequence
    float poly = fma_(t, a5, a4);
    poly = fma_(t, poly, a3);
    poly = fma_(t, poly, a2);
    poly = fma_(t, poly, a1);
    poly = mul(t, poly);
    float e = ExpNegForErf(-mul(x, x)); // == Exp(-x*x) wherever r depends on it
    float r = fma_(-poly, e, 1.0f); // NegMulAdd(poly, e, one) = one - poly*e
    return mul(sign, r);
}
// reference GELU (hwy_kernels-inl.h:821-843): 0.5 * (x * (1 + erf(x * 1/sqrt2)))
__device__ __forceinline__ float Gelu(float x) {
    float s = mul(x, 0.7071067811865476f);
    float inner = add(1.0f, Erf(s));
    return mul(0.5f, mul(x, inner));
}
// reference Sigmoid / SiLU (hwy_kernels-inl.h:786-816, 845-880)
__device__ __forceinline__ float Sigmoid(float x) {
    float e = Exp(-fabsf(x));
    float denom = add(1.0f, e);
    return __fdiv_rn((x >= 0.0f) ? 1.0f : e, denom); // IfThenElse(x >= 0, 1 / d, e / d): ONE division of the selected numerator
}
__device__ __forceinline__ float Silu(float x) {
    float e = Exp(-fabsf(x));
    float denom = add(1.0f, e);
    return __fdiv_rn((x >= 0.0f) ? x : mul(x, e), denom); // one division of the selected numerator (same values as two)
}
// reference pow (hwy_kernels-inl.h:142-161): exp(b * log(a))
__device__ __forceinline__ float Pow(float a, float b) { return Exp(mul(b, Log(a))); }
// reference hypot (hwy_kernels-inl.h:180-199): sqrt(a*a + b*b)
__device__ __forceinline__ float Hypot(float a, float b) { return __fsqrt_rn(add(mul(a, a), mul(b, b))); }
// reference fmod vector body (hwy_kernels-inl.h:201-222): a - trunc(a/b)*b   [NegMulAdd]
__device__ __forceinline__ float Fmod(float a, float b) {
    float q = truncf(__fdiv_rn(a, b));
    return fma_(-q, b, a);
}
// reference Cbrt (hwy_math_ext-inl.h:48-89)
__device__ __forceinline__ float Cbrt(float x) {
    float sign = copysignf(1.0f, x);
    float ax = fabsf(x);
    float y = Exp(mul(Log(ax), (float)(1.0 / 3.0)));
    float ysq = mul(y, y);
    y = __fdiv_rn(fma_(2.0f, y, __fdiv_rn(ax, ysq)), 3.0f);
    ysq = mul(y, y);
    y = __fdiv_rn(fma_(2.0f, y, __fdiv_rn(ax, ysq)), 3.0f);
    return (ax == 0.0f) ? 0.0f : mul(sign, y);
}

} // namespace hwym
