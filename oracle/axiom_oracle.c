/*
 * oracle/axiom_oracle.c — TEST INFRASTRUCTURE ONLY (never linked or loaded by axiom_b200/).
 *
 * Plain-C restatement of the fp32 math the reference CPU backend runs for contiguous Float32 tensors.
 * The reference evaluates these through Google Highway's polynomial kernels for every full SIMD vector
 * (src/backends/cpu/simd/hwy_kernels-inl.h:420-530,786-880 -> third_party/highway/hwy/contrib/math/
 * math-inl.h, pinned submodule highway@da7098a) and through its own erf/cbrt
 * (src/backends/cpu/simd/hwy_math_ext-inl.h:1-89). Each function below cites the lines it follows and
 * performs the SAME sequence of IEEE-754 binary32 operations (mul/add/fma/div/sqrt, truncating
 * float->int32), so results are bit-identical to the reference's vector body on any FMA-capable host.
 *
 * Pinned by tests/test_oracle_pin.py against oracle/_ref/libaxiom_ref.so (the reference itself).
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle/axiom_oracle.c -o oracle/liboracle.so -lm
 * (-ffp-contract=off: gcc must not fuse the un-fused Mul/Add steps; the fused ones call fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float f_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t bits_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
/* hn::ConvertTo(int32, float) on x86: truncation, saturating on overflow (x86_128-inl.h fix-ups) */
static inline int32_t cvt_trunc(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

/* math-inl.h:385-403 Estrin(x, c0..c3) and (x, c0..c5) */
static inline float estrin4(float x, float c0, float c1, float c2, float c3) {
    float x2 = x * x;
    return fmaf(x2, fmaf(c3, x, c2), fmaf(c1, x, c0));
}
static inline float estrin6(float x, float c0, float c1, float c2, float c3, float c4, float c5) {
    float x2 = x * x;
    float x4 = x2 * x2;
    return fmaf(x4, fmaf(c5, x, c4), fmaf(x2, fmaf(c3, x, c2), fmaf(c1, x, c0)));
}
/* ExpImpl<float>: math-inl.h:818-877 */
static inline float exp_poly(float x) {
    return fmaf(estrin6(x, +0.5f, +0.166666671633720397949219f, +0.0416664853692054748535156f,
                        +0.00833336077630519866943359f, +0.00139304355252534151077271f,
                        +0.000198527617612853646278381f),
                x * x, x);
}
static inline float pow2i(int32_t x) { return f_from_bits((uint32_t)(x + 0x7F) << 23); }
static inline float load_exp_short_range(float x, int32_t e) {
    int32_t y = e >> 1;
    return (x * pow2i(y)) * pow2i(e - y);
}
static inline float exp_reduce(float x, int32_t q) {
    float qf = (float)q;
    x = fmaf(qf, -0.693145751953125f, x);
    x = fmaf(qf, -1.428606765330187045e-6f, x);
    return x;
}
/* Exp: math-inl.h:1465-1485 */
float axo_exp(float x) {
    const float kOneOverLog2 = 1.442695040888963407359924681f;
    float half_signed = f_from_bits(bits_of(0.5f) | (bits_of(x) & 0x80000000u));
    int32_t q = cvt_trunc(fmaf(x, kOneOverLog2, half_signed));
    float y = load_exp_short_range(exp_poly(exp_reduce(x, q)) + 1.0f, q);
    return (x >= -104.0f) ? y : 0.0f;
}
/* Expm1: math-inl.h:1507-1531 */
float axo_expm1(float x) {
    const float kOneOverLog2 = 1.442695040888963407359924681f;
    float half_signed = f_from_bits(bits_of(0.5f) | (bits_of(x) & 0x80000000u));
    int32_t q = cvt_trunc(fmaf(x, kOneOverLog2, half_signed));
    float y = exp_poly(exp_reduce(x, q));
    float z = (fabsf(x) < 0.346573590279972654708616f) ? y : (load_exp_short_range(y + 1.0f, q) - 1.0f);
    return (x < -104.0f) ? -1.0f : z;
}
/* Log: math-inl.h:1002-1063 with LogImpl<float> :880-902 */
float axo_log(float x) {
    const float kLn2Hi = 0.69313812256f, kLn2Lo = 9.0580006145e-6f;
    const float kMinNormal = 1.175494351e-38f, kScale = 3.355443200e+7f;
    const int32_t kMagic = 0x3F3504F3, kExpMask = 0x3F800000, kManMask = 0x7FFFFF;
    int is_denormal = x < kMinNormal;
    x = is_denormal ? x * kScale : x;
    int32_t exp_bits = (int32_t)(bits_of(x) + (uint32_t)(kExpMask - kMagic));
    int32_t exp_scale = is_denormal ? -25 : 0;
    int32_t l2 = (int32_t)((uint32_t)exp_bits >> 23) - 0x7F;
    float e = (float)(exp_scale + l2);
    float y = f_from_bits((uint32_t)((exp_bits & kManMask) + kMagic));
    float ym1 = y - 1.0f;
    float z = ym1 / (y + 1.0f);
    float z2 = z * z, z4 = z2 * z2;
    float poly = fmaf(fmaf(0.28498786688f, z4, 0.66666662693f), z2, fmaf(0.24279078841f, z4, 0.40000972152f) * z4);
    float inner = fmaf(z, ym1 - poly, -(e * kLn2Lo));
    return fmaf(e, kLn2Hi, -(inner - ym1));
}
/* CosSinImpl<float>::Poly math-inl.h:679-687 */
static inline float sincos_poly(float x) {
    float y = x * x;
    return fmaf(estrin4(y, -1.66666597127914428710938e-1f, +8.33307858556509017944336e-3f,
                        -1.981069071916863322258e-4f, +2.6083159809786593541503e-6f),
                y * x, x);
}
/* Cos: math-inl.h:1442-1463 */
float axo_cos(float x) {
    const float kOneOverPi = 0.31830988618379067153f;
    float y = fabsf(x);
    int32_t q = (int32_t)((uint32_t)cvt_trunc(y * kOneOverPi) << 1) + 1;
    float qf = (float)q;
    y = fmaf(qf, -0.5f * 3.140625f, y);
    y = fmaf(qf, -0.5f * 0.0009670257568359375f, y);
    y = fmaf(qf, -0.5f * 6.2771141529083251953e-7f, y);
    y = fmaf(qf, -0.5f * 1.2154201256553420762e-10f, y);
    uint32_t sign = ((uint32_t)(~q & 2)) << 30;
    return sincos_poly(f_from_bits(bits_of(y) ^ sign));
}
/* Sin: math-inl.h:1563-1585 */
float axo_sin(float x) {
    const float kOneOverPi = 0.31830988618379067153f;
    float ax = fabsf(x);
    uint32_t sign_x = bits_of(ax) ^ bits_of(x);
    int32_t q = cvt_trunc(fmaf(ax, kOneOverPi, 0.5f));
    float qf = (float)q;
    float y = ax;
    y = fmaf(qf, -3.140625f, y);
    y = fmaf(qf, -0.0009670257568359375f, y);
    y = fmaf(qf, -6.2771141529083251953e-7f, y);
    y = fmaf(qf, -1.2154201256553420762e-10f, y);
    uint32_t sign = (((uint32_t)q & 1u) << 31) ^ sign_x;
    return sincos_poly(f_from_bits(bits_of(y) ^ sign));
}
/* reference tan = sin/cos: hwy_kernels-inl.h:500-517 */
float axo_tan(float x) { return axo_sin(x) / axo_cos(x); }
/* Tanh: math-inl.h:1601-1613 */
float axo_tanh(float x) {
    uint32_t sign = bits_of(x) & 0x80000000u;
    float ax = f_from_bits(bits_of(x) ^ sign);
    float y = axo_expm1(ax * 2.0f);
    float z = (ax > 18.714973875f) ? 1.0f : y / (y + 2.0f);
    return f_from_bits(bits_of(z) ^ sign);
}
/* reference Erf (Abramowitz-Stegun): hwy_math_ext-inl.h:1-37 */
float axo_erf(float x) {
    const float p = 0.3275911f, a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f,
                a5 = 1.061405429f;
    float sign = copysignf(1.0f, x);
    float ax = fabsf(x);
    float t = 1.0f / fmaf(p, ax, 1.0f);
    float poly = fmaf(t, a5, a4);
    poly = fmaf(t, poly, a3);
    poly = fmaf(t, poly, a2);
    poly = fmaf(t, poly, a1);
    poly = t * poly;
    float e = axo_exp(-(x * x));
    float r = fmaf(-poly, e, 1.0f);
    return sign * r;
}
/* reference GELU: hwy_kernels-inl.h:821-843 */
float axo_gelu(float x) {
    float s = x * 0.7071067811865476f;
    float inner = 1.0f + axo_erf(s);
    return 0.5f * (x * inner);
}
/* reference sigmoid / silu: hwy_kernels-inl.h:786-816,845-880 */
float axo_sigmoid(float x) {
    float e = axo_exp(-fabsf(x));
    float d = 1.0f + e;
    return (x >= 0.0f) ? 1.0f / d : e / d;
}
float axo_silu(float x) {
    float e = axo_exp(-fabsf(x));
    float d = 1.0f + e;
    return (x >= 0.0f) ? x / d : (x * e) / d;
}
/* reference pow / hypot / cbrt: hwy_kernels-inl.h:142-199, hwy_math_ext-inl.h:48-89 */
float axo_pow(float a, float b) { return axo_exp(b * axo_log(a)); }
float axo_hypot(float a, float b) { return sqrtf(a * a + b * b); }
float axo_cbrt(float x) {
    float sign = copysignf(1.0f, x);
    float ax = fabsf(x);
    float y = axo_exp(axo_log(ax) * (float)(1.0 / 3.0));
    float ysq = y * y;
    y = fmaf(2.0f, y, ax / ysq) / 3.0f;
    ysq = y * y;
    y = fmaf(2.0f, y, ax / ysq) / 3.0f;
    return (ax == 0.0f) ? 0.0f : sign * y;
}

/* ---- array drivers: vector body = polynomial, scalar tail (n % lanes) = libm, exactly like the
 *      reference loops `for (; i + N <= n; i += N) ... for (; i < n; ++i) std::f(x)` ------------------ */
typedef float (*un_fn)(float);
static void drive1(un_fn body, un_fn tail, const float *in, float *out, int64_t n, int lanes) {
    int64_t nb = lanes > 0 ? (n / lanes) * lanes : n;
    for (int64_t i = 0; i < nb; ++i) out[i] = body(in[i]);
    for (int64_t i = nb; i < n; ++i) out[i] = tail(in[i]);
}
static float t_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }
static float t_sigmoid(float x) { float e = expf(-fabsf(x)); float d = 1.0f + e; return x >= 0.0f ? 1.0f / d : e / d; }
static float t_silu(float x) { float e = expf(-fabsf(x)); float d = 1.0f + e; return x >= 0.0f ? x / d : (x * e) / d; }

/* op codes follow axiom::ops::OpType */
int axo_unary_f32(int op, const float *in, float *out, int64_t n, int lanes) {
    switch (op) {
    case 28: drive1(axo_exp, expf, in, out, n, lanes); return 0;
    case 29: drive1(axo_log, logf, in, out, n, lanes); return 0;
    case 30: drive1(axo_sin, sinf, in, out, n, lanes); return 0;
    case 31: drive1(axo_cos, cosf, in, out, n, lanes); return 0;
    case 32: drive1(axo_tan, tanf, in, out, n, lanes); return 0;
    case 33: drive1(axo_erf, erff, in, out, n, lanes); return 0;
    case 41: drive1(axo_cbrt, cbrtf, in, out, n, lanes); return 0;
    case 50: drive1(axo_silu, t_silu, in, out, n, lanes); return 0;
    case 51: drive1(axo_sigmoid, t_sigmoid, in, out, n, lanes); return 0;
    case 52: drive1(axo_tanh, tanhf, in, out, n, lanes); return 0;
    case 53: drive1(axo_gelu, t_gelu, in, out, n, lanes); return 0;
    default: return -1;
    }
}
int axo_binary_f32(int op, const float *a, const float *b, float *out, int64_t n, int lanes) {
    int64_t nb = lanes > 0 ? (n / lanes) * lanes : n;
    switch (op) {
    case 4:
        for (int64_t i = 0; i < nb; ++i) out[i] = axo_pow(a[i], b[i]);
        for (int64_t i = nb; i < n; ++i) out[i] = powf(a[i], b[i]);
        return 0;
    case 24:
        for (int64_t i = 0; i < nb; ++i) out[i] = axo_hypot(a[i], b[i]);
        for (int64_t i = nb; i < n; ++i) out[i] = hypotf(a[i], b[i]);
        return 0;
    default: return -1;
    }
}

/* softmax row (cpu_operations.cpp:3123-3162): max, libm expf(x-max) stored, SERIAL fp32 sum, divide */
void axo_softmax_rows_f32(const float *in, float *out, int64_t rows, int64_t n, int is_log) {
    for (int64_t r = 0; r < rows; ++r) {
        const float *x = in + r * n;
        float *y = out + r * n;
        float m = -3.402823466e+38f;
        for (int64_t k = 0; k < n; ++k) m = (m < x[k]) ? x[k] : m; /* std::max(max_val, x) */
        float s = 0.0f;
        for (int64_t k = 0; k < n; ++k) { float e = expf(x[k] - m); y[k] = e; s += e; }
        if (is_log) { float ls = logf(s); for (int64_t k = 0; k < n; ++k) y[k] = x[k] - m - ls; }
        else { for (int64_t k = 0; k < n; ++k) y[k] = y[k] / s; }
    }
}

/* bf16 matmul fallback accumulates in bf16 (cpu_operations.cpp:1835-1882) — NOT the parity target, kept
 * only to document the reference's behaviour for small cases. */
