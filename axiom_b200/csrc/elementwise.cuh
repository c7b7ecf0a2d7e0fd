// elementwise.cuh — device side of kernel family 1: the fused elementwise / broadcast / where chain.
//
// One kernel runs a whole fusion group: every thread owns VEC consecutive output
// elements of one innermost run, keeps the chain's live values in registers and walks a tiny
// warp-uniform micro-program (decoded from kernel parameters = constant bank, no divergence).
// Memory side: inputs whose innermost run is contiguous and 16-byte aligned are read with 128-bit
// loads, broadcast inputs with one scalar load, anything else with per-lane strided loads; the output is
// written with 128-bit stores. HBM-bound by construction: every operand byte is read once.
//
// Register machine (per thread): acc, t, u and NSLOT save slots, each VEC words wide.
//   canonical value forms   F: fp32 (bf16/fp16 values are fp32 values already rounded to that format)
//                           I: 32-bit integer (narrow ints sign/zero-extended, bool = 0/1)
//                           D: fp64, L: 64-bit integer          (only in the 64-bit-word instantiation)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hwy_math.cuh"

namespace axb {
namespace ew {

constexpr int kMaxMicro = 96;
constexpr int kMaxImm = 16;
constexpr int kMaxIn = 8;
constexpr int kNSlot = 4;
constexpr int kMaxOuter = 7;

enum Sel : uint8_t { R_T = 0, R_ACC = 1, R_U = 2, R_S0 = 3 }; // R_S0 + k = slot k
enum Mode : uint8_t { MODE_VEC = 0, MODE_STRIDED = 1, MODE_BCAST = 2 };
enum Round : uint8_t { RND_NONE = 0, RND_BF16 = 1, RND_F16 = 2 };
enum Flag : uint8_t { FL_SWAP = 4 }; // low 2 bits of flags = Round

enum M : uint8_t {
    M_END = 0,
    M_LOAD,  // a = input, b = dst sel (t/acc/u)
    M_IMM,   // a = imm index, b = dst sel
    M_MOV,   // a = src sel, b = dst sel
    M_STORE, // store acc
    M_SEL,   // acc = (acc != 0) ? t : u
    // conversions, a = register sel (t/acc/u)
    M_CVT_RND_BF16, M_CVT_RND_F16, M_CVT_I2F, M_CVT_U2F, M_CVT_F2I, M_CVT_F2U, M_CVT_F2BOOL, M_CVT_I2BOOL,
    M_CVT_WRAP_I8, M_CVT_WRAP_U8, M_CVT_WRAP_I16, M_CVT_WRAP_U16,
    // 64-bit word conversions
    M_CVT_F2D, M_CVT_D2F, M_CVT_I2L, M_CVT_U2L, M_CVT_L2I, M_CVT_L2F, M_CVT_UL2F, M_CVT_L2D, M_CVT_UL2D,
    M_CVT_D2L, M_CVT_D2UL, M_CVT_F2L, M_CVT_F2UL, M_CVT_D2BOOL, M_CVT_L2BOOL, M_CVT_I2D, M_CVT_U2D, M_CVT_D2I,
    M_CVT_D2U,
    // F unary (acc)
    M_F_NEG, M_F_ABS, M_F_SQRT, M_F_EXP, M_F_LOG, M_F_SIN, M_F_COS, M_F_TAN, M_F_ERF, M_F_SIGN, M_F_FLOOR,
    M_F_CEIL, M_F_TRUNC, M_F_ROUND, M_F_RECIP, M_F_SQUARE, M_F_CBRT, M_F_RELU, M_F_LEAKY, M_F_SILU, M_F_SIGMOID,
    M_F_TANH, M_F_GELU, M_F_ISNAN, M_F_ISINF, M_F_ISFINITE, M_F_LNOT,
    // F unary, libm-accurate variants used for bf16/fp16 (reference evaluates those in double)
    M_FA_SQRT, M_FA_EXP, M_FA_LOG, M_FA_SIN, M_FA_COS, M_FA_TAN, M_FA_ERF, M_FA_CBRT, M_FA_SILU, M_FA_SIGMOID,
    M_FA_TANH, M_FA_GELU,
    // F binary: acc = f(acc, t)  (FL_SWAP: f(t, acc))
    M_F_ADD, M_F_SUB, M_F_MUL, M_F_DIV, M_F_POW, M_F_MOD, M_F_MAX, M_F_MIN, M_F_ATAN2, M_F_HYPOT, M_FA_POW,
    M_FA_HYPOT,
    M_F_EQ, M_F_NE, M_F_LT, M_F_LE, M_F_GT, M_F_GE, M_F_LAND, M_F_LOR, M_F_LXOR,
    // I unary / binary (32-bit ints; U = unsigned flavour where it matters)
    M_I_NEG, M_I_ABS, M_I_SIGN, M_I_SQUARE, M_I_RELU, M_I_RECIP, M_U_RECIP, M_I_LNOT,
    M_I_ADD, M_I_SUB, M_I_MUL, M_I_DIV, M_U_DIV, M_I_MOD, M_U_MOD, M_I_MAX, M_U_MAX, M_I_MIN, M_U_MIN, M_I_AND,
    M_I_OR, M_I_XOR, M_I_SHL, M_I_SHR, M_U_SHR,
    // D (fp64) and L (int64)
    M_D_NEG, M_D_ABS, M_D_SQRT, M_D_EXP, M_D_LOG, M_D_SIN, M_D_COS, M_D_TAN, M_D_TANH, M_D_ERF, M_D_SIGN,
    M_D_FLOOR, M_D_CEIL, M_D_TRUNC, M_D_ROUND, M_D_RECIP, M_D_SQUARE, M_D_RELU, M_D_SIGMOID, M_D_ISNAN,
    M_D_ISINF, M_D_ISFINITE, M_D_LNOT,
    M_D_ADD, M_D_SUB, M_D_MUL, M_D_DIV, M_D_POW, M_D_MOD, M_D_MAX, M_D_MIN, M_D_HYPOT, M_D_ATAN2,
    M_L_NEG, M_L_ABS, M_L_SIGN, M_L_SQUARE, M_L_RELU, M_L_LNOT,
    M_L_ADD, M_L_SUB, M_L_MUL, M_L_DIV, M_UL_DIV, M_L_MOD, M_UL_MOD, M_L_MAX, M_UL_MAX, M_L_MIN, M_UL_MIN,
    M_L_AND, M_L_OR, M_L_XOR, M_L_SHL, M_L_SHR, M_UL_SHR,
    M_COUNT
};

struct Operand {
    const char *ptr;            // base + byte_offset
    int64_t stride[kMaxOuter];  // byte strides of the collapsed OUTER dims (index 0 = dim next to inner)
    int64_t inner_stride;       // byte stride along the innermost run
    uint8_t dtype;
    uint8_t mode;
    uint8_t pad[6];
};

struct Params {
    Operand in[kMaxIn];
    Operand out;
    uint64_t imm[kMaxImm];
    uint32_t micro[kMaxMicro]; // op | a<<8 | b<<16 | flags<<24
    // iteration space: innermost run of `inner_len` elements, n_outer collapsed outer dims
    uint32_t inner_len;
    uint32_t nvec_inner; // ceil(inner_len / VEC)
    uint32_t total_vec;  // nvec_inner * prod(outer)  (< 2^31)
    uint32_t n_outer;
    uint32_t outer_shape[kMaxOuter];
    uint32_t div_m[kMaxOuter + 1]; // fast-div magics: [0] = nvec_inner, [1+d] = outer_shape[d]
    uint32_t div_s[kMaxOuter + 1];
    int32_t n_micro;
    int32_t reduce_op; // fused full reduction (axb_elementwise_reduce): 0 = none
    void *reduce_ws;   // per-block partials
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t m, uint32_t s) {
    return (__umulhi(n, m) + n) >> s;
}

// ---- canonicalisation of storage bits -------------------------------------------------------------
template <typename W> __device__ __forceinline__ W canon(uint64_t raw, int dt) {
    switch (dt) {
    case AXB_F32: case AXB_I32: case AXB_U32: return (W)(uint32_t)raw;
    case AXB_BF16: return (W)((uint32_t)raw << 16);
    case AXB_F16: return (W)__float_as_uint(__half2float(__ushort_as_half((unsigned short)raw)));
    case AXB_I16: return (W)(uint32_t)(int32_t)(int16_t)raw;
    case AXB_U16: return (W)(uint32_t)(uint16_t)raw;
    case AXB_I8: return (W)(uint32_t)(int32_t)(int8_t)raw;
    case AXB_U8: case AXB_BOOL: return (W)(uint32_t)(uint8_t)raw;
    default: return (W)raw; // 64-bit types
    }
}
__device__ __forceinline__ uint32_t f32_to_bf16_bits(float f) {
    // reference bfloat16_t(float): NaN -> 0x7FC0, else round-to-nearest-even (include/axiom/bfloat16.hpp:16-27)
    if (f != f)
        return 0x7FC0u;
    uint32_t b = __float_as_uint(f);
    b += 0x7FFFu + ((b >> 16) & 1u);
    return b >> 16;
}
__device__ __forceinline__ float round_bf16(float f) { return __uint_as_float(f32_to_bf16_bits(f) << 16); }
__device__ __forceinline__ float round_f16(float f) { return __half2float(__float2half_rn(f)); }

template <typename W> __device__ __forceinline__ uint64_t decanon(W v, int dt) {
    switch (dt) {
    case AXB_BF16: return f32_to_bf16_bits(__uint_as_float((uint32_t)v));
    case AXB_F16: return __half_as_ushort(__float2half_rn(__uint_as_float((uint32_t)v)));
    case AXB_BOOL: return ((uint32_t)v != 0) ? 1u : 0u;
    default: return (uint64_t)v;
    }
}

__device__ __forceinline__ uint64_t ld_raw(const char *p, int size) {
    switch (size) {
    case 1: return *reinterpret_cast<const uint8_t *>(p);
    case 2: return *reinterpret_cast<const uint16_t *>(p);
    case 4: return *reinterpret_cast<const uint32_t *>(p);
    default: return *reinterpret_cast<const uint64_t *>(p);
    }
}
__device__ __forceinline__ void st_raw(char *p, int size, uint64_t v) {
    switch (size) {
    case 1: *reinterpret_cast<uint8_t *>(p) = (uint8_t)v; break;
    case 2: *reinterpret_cast<uint16_t *>(p) = (uint16_t)v; break;
    case 4: *reinterpret_cast<uint32_t *>(p) = (uint32_t)v; break;
    default: *reinterpret_cast<uint64_t *>(p) = v; break;
    }
}
__device__ __forceinline__ int dsize(int dt) {
    return (dt == AXB_BOOL || dt == AXB_I8 || dt == AXB_U8)                      ? 1
           : (dt == AXB_I16 || dt == AXB_U16 || dt == AXB_F16 || dt == AXB_BF16) ? 2
           : (dt == AXB_I32 || dt == AXB_U32 || dt == AXB_F32)                   ? 4
                                                                                 : 8;
}

// streaming 128-bit accesses (read-once data: do not allocate in L1)
__device__ __forceinline__ uint4 ldg128(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg64(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg128(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg64(void *p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// Load VEC elements of operand `o` starting at element column `col` of the current run.
template <int VEC, typename W>
__device__ __forceinline__ void load_vec(const Operand &o, int64_t outer_off, uint32_t col, int valid,
                                         W (&r)[VEC]) {
    const int dt = o.dtype;
    const int sz = dsize(dt);
    const char *base = o.ptr + outer_off;
    if (o.mode == MODE_BCAST) {
        W v = canon<W>(ld_raw(base, sz), dt);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            r[i] = v;
        return;
    }
    if (VEC == 8 && o.mode == MODE_VEC && valid == VEC) {
        const char *p = base + (int64_t)col * sz;
        if (sz == 4) {
            uint4 a = ldg128(p), b = ldg128(p + 16);
            r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
            r[4 % VEC] = b.x; r[5 % VEC] = b.y; r[6 % VEC] = b.z; r[7 % VEC] = b.w;
            return;
        }
        if (sz == 2) {
            uint4 a = ldg128(p);
            uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[(2 * i) % VEC] = canon<W>(w[i] & 0xFFFFu, dt);
                r[(2 * i + 1) % VEC] = canon<W>(w[i] >> 16, dt);
            }
            return;
        }
        if (sz == 1) {
            uint2 a = ldg64(p);
            uint32_t w[2] = {a.x, a.y};
#pragma unroll
            for (int i = 0; i < 8; ++i)
                r[i % VEC] = canon<W>((w[i >> 2] >> ((i & 3) * 8)) & 0xFFu, dt);
            return;
        }
    }
    if (VEC == 4 && sizeof(W) == 8 && o.mode == MODE_VEC && valid == VEC && sz == 8) {
        const char *p = base + (int64_t)col * 8;
        uint4 a = ldg128(p), b = ldg128(p + 16);
        r[0] = (W)(((uint64_t)a.y << 32) | a.x);
        r[1 % VEC] = (W)(((uint64_t)a.w << 32) | a.z);
        r[2 % VEC] = (W)(((uint64_t)b.y << 32) | b.x);
        r[3 % VEC] = (W)(((uint64_t)b.w << 32) | b.z);
        return;
    }
    const int64_t is = o.inner_stride;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        r[i] = 0;
        if (i < valid)
            r[i] = canon<W>(ld_raw(base + (int64_t)(col + i) * is, sz), dt);
    }
}

template <int VEC, typename W>
__device__ __forceinline__ void store_vec(const Operand &o, int64_t outer_off, uint32_t col, int valid,
                                          const W (&r)[VEC]) {
    const int dt = o.dtype;
    const int sz = dsize(dt);
    char *base = const_cast<char *>(o.ptr) + outer_off;
    if (VEC == 8 && o.mode == MODE_VEC && valid == VEC) {
        char *p = base + (int64_t)col * sz;
        if (sz == 4) {
            stg128(p, make_uint4((uint32_t)r[0], (uint32_t)r[1 % VEC], (uint32_t)r[2 % VEC], (uint32_t)r[3 % VEC]));
            stg128(p + 16,
                   make_uint4((uint32_t)r[4 % VEC], (uint32_t)r[5 % VEC], (uint32_t)r[6 % VEC], (uint32_t)r[7 % VEC]));
            return;
        }
        if (sz == 2) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                w[i] = (uint32_t)(decanon<W>(r[(2 * i) % VEC], dt) & 0xFFFFu) |
                       ((uint32_t)(decanon<W>(r[(2 * i + 1) % VEC], dt) & 0xFFFFu) << 16);
            stg128(p, make_uint4(w[0], w[1], w[2], w[3]));
            return;
        }
        if (sz == 1) {
            uint32_t w[2] = {0, 0};
#pragma unroll
            for (int i = 0; i < 8; ++i)
                w[i >> 2] |= (uint32_t)(decanon<W>(r[i % VEC], dt) & 0xFFu) << ((i & 3) * 8);
            stg64(p, make_uint2(w[0], w[1]));
            return;
        }
    }
    if (VEC == 4 && sizeof(W) == 8 && o.mode == MODE_VEC && valid == VEC && sz == 8) {
        char *p = base + (int64_t)col * 8;
        stg128(p, make_uint4((uint32_t)r[0], (uint32_t)((uint64_t)r[0] >> 32), (uint32_t)r[1 % VEC],
                             (uint32_t)((uint64_t)r[1 % VEC] >> 32)));
        stg128(p + 16, make_uint4((uint32_t)r[2 % VEC], (uint32_t)((uint64_t)r[2 % VEC] >> 32), (uint32_t)r[3 % VEC],
                                  (uint32_t)((uint64_t)r[3 % VEC] >> 32)));
        return;
    }
    const int64_t is = o.inner_stride;
#pragma unroll
    for (int i = 0; i < VEC; ++i)
        if (i < valid)
            st_raw(base + (int64_t)(col + i) * is, sz, decanon<W>(r[i], dt));
}

// ---- scalar op helpers ---------------------------------------------------------------------------
template <typename W> __device__ __forceinline__ float asF(W v) { return __uint_as_float((uint32_t)v); }
template <typename W> __device__ __forceinline__ W fromF(float f) { return (W)__float_as_uint(f); }
template <typename W> __device__ __forceinline__ int32_t asI(W v) { return (int32_t)(uint32_t)v; }
template <typename W> __device__ __forceinline__ W fromI(int32_t i) { return (W)(uint32_t)i; }
__device__ __forceinline__ double asD(uint64_t v) { return __longlong_as_double((long long)v); }
__device__ __forceinline__ uint64_t fromD(double d) { return (uint64_t)__double_as_longlong(d); }

__device__ __forceinline__ float f_sign(float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); }
// Cold paths (double-precision libm, used for half dtypes / fp64 / rarely used ops). Kept out of line so
// their register needs do not inflate the hot interpreter loop.
namespace slow {
#define AXB_SLOW1(NAME, EXPR) __device__ __noinline__ float NAME(float a) { return (float)(EXPR); }
AXB_SLOW1(f_exp, exp((double)a))
AXB_SLOW1(f_log, log((double)a))
AXB_SLOW1(f_sin, sin((double)a))
AXB_SLOW1(f_cos, cos((double)a))
AXB_SLOW1(f_tan, tan((double)a))
AXB_SLOW1(f_erf, erf((double)a))
AXB_SLOW1(f_cbrt, cbrt((double)a))
AXB_SLOW1(f_tanh, tanh((double)a))
#undef AXB_SLOW1
__device__ __noinline__ float f_pow(float a, float b) { return (float)pow((double)a, (double)b); }
__device__ __noinline__ float f_hypot(float a, float b) { return (float)hypot((double)a, (double)b); }
__device__ __noinline__ float f_atan2(float a, float b) { return atan2f(a, b); }
__device__ __noinline__ float f_fmod(float a, float b) { return fmodf(a, b); }
#define AXB_SLOWD1(NAME, EXPR) __device__ __noinline__ double NAME(double a) { return (EXPR); }
AXB_SLOWD1(d_exp, exp(a))
AXB_SLOWD1(d_log, log(a))
AXB_SLOWD1(d_sin, sin(a))
AXB_SLOWD1(d_cos, cos(a))
AXB_SLOWD1(d_tan, tan(a))
AXB_SLOWD1(d_tanh, tanh(a))
AXB_SLOWD1(d_erf, erf(a))
AXB_SLOWD1(d_sigmoid, a >= 0.0 ? 1.0 / (1.0 + exp(-fabs(a))) : exp(-fabs(a)) / (1.0 + exp(-fabs(a))))
#undef AXB_SLOWD1
__device__ __noinline__ double d_pow(double a, double b) { return pow(a, b); }
__device__ __noinline__ double d_fmod(double a, double b) { return fmod(a, b); }
__device__ __noinline__ double d_hypot(double a, double b) { return hypot(a, b); }
__device__ __noinline__ double d_atan2(double a, double b) { return atan2(a, b); }
// reference scalar functors for half types evaluate in double (cpu_operations.hpp:635-700)
__device__ __noinline__ float f_gelu(float x) {
    double d = (double)x;
    return (float)(0.5 * d * (1.0 + erf(d / 1.4142135623730951)));
}
__device__ __noinline__ float f_sigmoid(float x) {
    double d = (double)x, e = exp(-fabs(d)), den = 1.0 + e;
    return (float)(d >= 0.0 ? 1.0 / den : e / den);
}
__device__ __noinline__ float f_silu(float x) {
    double d = (double)x, e = exp(-fabs(d)), den = 1.0 + e;
    return (float)(d >= 0.0 ? d / den : (d * e) / den);
}
} // namespace slow
__device__ __forceinline__ int32_t i_div(int32_t a, int32_t b) {
    return b == 0 ? 0 : ((a == INT32_MIN && b == -1) ? INT32_MIN : a / b);
}
__device__ __forceinline__ int32_t i_mod(int32_t a, int32_t b) {
    return b == 0 ? 0 : ((b == -1) ? 0 : a % b);
}
__device__ __forceinline__ int64_t l_div(int64_t a, int64_t b) {
    return b == 0 ? 0 : ((a == INT64_MIN && b == -1) ? INT64_MIN : a / b);
}
__device__ __forceinline__ int64_t l_mod(int64_t a, int64_t b) { return b == 0 ? 0 : ((b == -1) ? 0 : a % b); }

#define AXB_UN_F(EXPR)                                                                                  \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            float a = asF<W>(acc[i]);                                                                   \
            acc[i] = fromF<W>(EXPR);                                                                    \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_UN_F2I(EXPR)                                                                                \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            float a = asF<W>(acc[i]);                                                                   \
            acc[i] = fromI<W>((EXPR) ? 1 : 0);                                                          \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_BIN_F(EXPR)                                                                                 \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            float a = asF<W>(swap ? t[i] : acc[i]);                                                     \
            float b = asF<W>(swap ? acc[i] : t[i]);                                                     \
            acc[i] = fromF<W>(EXPR);                                                                    \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_CMP_F(EXPR)                                                                                 \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            float a = asF<W>(swap ? t[i] : acc[i]);                                                     \
            float b = asF<W>(swap ? acc[i] : t[i]);                                                     \
            acc[i] = fromI<W>((EXPR) ? 1 : 0);                                                          \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_UN_I(EXPR)                                                                                  \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            int32_t a = asI<W>(acc[i]);                                                                 \
            uint32_t ua = (uint32_t)a;                                                                  \
            (void)ua;                                                                                   \
            acc[i] = fromI<W>((int32_t)(EXPR));                                                         \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_BIN_I(EXPR)                                                                                 \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            int32_t a = asI<W>(swap ? t[i] : acc[i]);                                                   \
            int32_t b = asI<W>(swap ? acc[i] : t[i]);                                                   \
            uint32_t ua = (uint32_t)a, ub = (uint32_t)b;                                                \
            (void)ua;                                                                                   \
            (void)ub;                                                                                   \
            acc[i] = fromI<W>((int32_t)(EXPR));                                                         \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_UN_D(EXPR)                                                                                  \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                double a = asD(acc[i]);                                                                 \
                acc[i] = fromD(EXPR);                                                                   \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_UN_D2I(EXPR)                                                                                \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                double a = asD(acc[i]);                                                                 \
                acc[i] = (W)((EXPR) ? 1 : 0);                                                           \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_BIN_D(EXPR)                                                                                 \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                double a = asD(swap ? t[i] : acc[i]);                                                   \
                double b = asD(swap ? acc[i] : t[i]);                                                   \
                acc[i] = fromD(EXPR);                                                                   \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_UN_L(EXPR)                                                                                  \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                int64_t a = (int64_t)acc[i];                                                            \
                acc[i] = (W)(int64_t)(EXPR);                                                            \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_BIN_L(EXPR)                                                                                 \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                int64_t a = (int64_t)(swap ? t[i] : acc[i]);                                            \
                int64_t b = (int64_t)(swap ? acc[i] : t[i]);                                            \
                uint64_t ua = (uint64_t)a, ub = (uint64_t)b;                                            \
                (void)ua;                                                                               \
                (void)ub;                                                                               \
                acc[i] = (W)(int64_t)(EXPR);                                                            \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   \
    break;
// conversion applied to the register selected by `a` (t / acc / u)
#define AXB_CVT(EXPR)                                                                                   \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            W x = (ra == R_T) ? t[i] : (ra == R_ACC ? acc[i] : u[i]);                                   \
            W y = (W)(EXPR);                                                                            \
            if (ra == R_T)                                                                              \
                t[i] = y;                                                                               \
            else if (ra == R_ACC)                                                                       \
                acc[i] = y;                                                                             \
            else                                                                                        \
                u[i] = y;                                                                               \
        }                                                                                               \
    }                                                                                                   \
    break;
#define AXB_CVT64(EXPR)                                                                                 \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                W x = (ra == R_T) ? t[i] : (ra == R_ACC ? acc[i] : u[i]);                               \
                W y = (W)(EXPR);                                                                        \
                if (ra == R_T)                                                                          \
                    t[i] = y;                                                                           \
                else if (ra == R_ACC)                                                                   \
                    acc[i] = y;                                                                         \
                else                                                                                    \
                    u[i] = y;                                                                           \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   \
    break;

// Executes the micro-program for one vector. Returns with the final value in acc (after M_STORE if any).
template <int VEC, typename W>
__device__ __forceinline__ void run_program(const Params &p, const int64_t *__restrict__ coord_unused,
                                            const uint32_t (&c)[kMaxOuter], uint32_t col, int valid,
                                            W (&acc)[VEC], bool do_store) {
    (void)coord_unused;
    W t[VEC], u[VEC], s[kNSlot][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        t[i] = 0;
        u[i] = 0;
        acc[i] = 0;
#pragma unroll
        for (int k = 0; k < kNSlot; ++k)
            s[k][i] = 0;
    }
    const int n_outer = p.n_outer;
    auto outer_off = [&](const Operand &o) -> int64_t {
        int64_t off = 0;
#pragma unroll
        for (int d = 0; d < kMaxOuter; ++d)
            if (d < n_outer)
                off += (int64_t)c[d] * o.stride[d];
        return off;
    };

    for (int pc = 0; pc < p.n_micro; ++pc) {
        const uint32_t m = p.micro[pc];
        const uint32_t op = m & 0xFFu, ra = (m >> 8) & 0xFFu, rb = (m >> 16) & 0xFFu, fl = m >> 24;
        const bool swap = (fl & FL_SWAP) != 0;
        switch (op) {
        case M_LOAD: {
            const Operand &o = p.in[ra];
            int64_t off = outer_off(o);
            if (rb == R_T)
                load_vec<VEC, W>(o, off, col, valid, t);
            else if (rb == R_ACC)
                load_vec<VEC, W>(o, off, col, valid, acc);
            else
                load_vec<VEC, W>(o, off, col, valid, u);
            break;
        }
        case M_IMM: {
            W v = (W)p.imm[ra];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                if (rb == R_T)
                    t[i] = v;
                else if (rb == R_ACC)
                    acc[i] = v;
                else
                    u[i] = v;
            }
            break;
        }
        case M_MOV: {
            W tmp[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                switch (ra) {
                case R_T: tmp[i] = t[i]; break;
                case R_ACC: tmp[i] = acc[i]; break;
                case R_U: tmp[i] = u[i]; break;
                case R_S0: tmp[i] = s[0][i]; break;
                case R_S0 + 1: tmp[i] = s[1][i]; break;
                case R_S0 + 2: tmp[i] = s[2][i]; break;
                default: tmp[i] = s[3][i]; break;
                }
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                switch (rb) {
                case R_T: t[i] = tmp[i]; break;
                case R_ACC: acc[i] = tmp[i]; break;
                case R_U: u[i] = tmp[i]; break;
                case R_S0: s[0][i] = tmp[i]; break;
                case R_S0 + 1: s[1][i] = tmp[i]; break;
                case R_S0 + 2: s[2][i] = tmp[i]; break;
                default: s[3][i] = tmp[i]; break;
                }
            }
            break;
        }
        case M_STORE:
            if (do_store)
                store_vec<VEC, W>(p.out, outer_off(p.out), col, valid, acc);
            break;
        case M_SEL: {
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                acc[i] = (acc[i] != 0) ? t[i] : u[i];
            break;
        }
        // ---- conversions -------------------------------------------------------------------------
        case M_CVT_RND_BF16: AXB_CVT(fromF<W>(round_bf16(asF<W>(x))))
        case M_CVT_RND_F16: AXB_CVT(fromF<W>(round_f16(asF<W>(x))))
        case M_CVT_I2F: AXB_CVT(fromF<W>(__int2float_rn(asI<W>(x))))
        case M_CVT_U2F: AXB_CVT(fromF<W>(__uint2float_rn((uint32_t)x)))
        case M_CVT_F2I: AXB_CVT(fromI<W>(__float2int_rz(asF<W>(x))))
        case M_CVT_F2U: AXB_CVT((W)__float2uint_rz(asF<W>(x)))
        case M_CVT_F2BOOL: AXB_CVT((W)(asF<W>(x) != 0.0f ? 1u : 0u))
        case M_CVT_I2BOOL: AXB_CVT((W)((uint32_t)x != 0u ? 1u : 0u))
        case M_CVT_WRAP_I8: AXB_CVT(fromI<W>((int32_t)(int8_t)(uint32_t)x))
        case M_CVT_WRAP_U8: AXB_CVT((W)((uint32_t)x & 0xFFu))
        case M_CVT_WRAP_I16: AXB_CVT(fromI<W>((int32_t)(int16_t)(uint32_t)x))
        case M_CVT_WRAP_U16: AXB_CVT((W)((uint32_t)x & 0xFFFFu))
        case M_CVT_F2D: AXB_CVT64(fromD((double)asF<W>(x)))
        case M_CVT_D2F: AXB_CVT64(fromF<W>((float)asD(x)))
        case M_CVT_I2L: AXB_CVT64((W)(int64_t)asI<W>(x))
        case M_CVT_U2L: AXB_CVT64((W)(uint64_t)(uint32_t)x)
        case M_CVT_L2I: AXB_CVT64((W)(uint32_t)x)
        case M_CVT_L2F: AXB_CVT64(fromF<W>(__ll2float_rn((long long)x)))
        case M_CVT_UL2F: AXB_CVT64(fromF<W>(__ull2float_rn((unsigned long long)x)))
        case M_CVT_L2D: AXB_CVT64(fromD(__ll2double_rn((long long)x)))
        case M_CVT_UL2D: AXB_CVT64(fromD(__ull2double_rn((unsigned long long)x)))
        case M_CVT_D2L: AXB_CVT64((W)__double2ll_rz(asD(x)))
        case M_CVT_D2UL: AXB_CVT64((W)__double2ull_rz(asD(x)))
        case M_CVT_F2L: AXB_CVT64((W)__float2ll_rz(asF<W>(x)))
        case M_CVT_F2UL: AXB_CVT64((W)__float2ull_rz(asF<W>(x)))
        case M_CVT_D2BOOL: AXB_CVT64((W)(asD(x) != 0.0 ? 1u : 0u))
        case M_CVT_L2BOOL: AXB_CVT64((W)(x != 0 ? 1u : 0u))
        case M_CVT_I2D: AXB_CVT64(fromD((double)asI<W>(x)))
        case M_CVT_U2D: AXB_CVT64(fromD((double)(uint32_t)x))
        case M_CVT_D2I: AXB_CVT64(fromI<W>(__double2int_rz(asD(x))))
        case M_CVT_D2U: AXB_CVT64((W)__double2uint_rz(asD(x)))
        // ---- fp32 unary (Highway-exact forms) ------------------------------------------------------
        case M_F_NEG: AXB_UN_F(__uint_as_float(__float_as_uint(a) ^ 0x80000000u))
        case M_F_ABS: AXB_UN_F(fabsf(a))
        case M_F_SQRT: AXB_UN_F(__fsqrt_rn(a))
        case M_F_EXP: AXB_UN_F(hwym::Exp(a))
        case M_F_LOG: AXB_UN_F(hwym::Log(a))
        case M_F_SIN: AXB_UN_F(hwym::Sin(a))
        case M_F_COS: AXB_UN_F(hwym::Cos(a))
        case M_F_TAN: AXB_UN_F(hwym::Tan(a))
        case M_F_ERF: AXB_UN_F(hwym::Erf(a))
        case M_F_SIGN: AXB_UN_F(f_sign(a))
        case M_F_FLOOR: AXB_UN_F(floorf(a))
        case M_F_CEIL: AXB_UN_F(ceilf(a))
        case M_F_TRUNC: AXB_UN_F(truncf(a))
        case M_F_ROUND: AXB_UN_F(rintf(a)) // hn::Round = nearest-even (hwy_kernels-inl.h:640-655)
        case M_F_RECIP: AXB_UN_F(__fdiv_rn(1.0f, a))
        case M_F_SQUARE: AXB_UN_F(__fmul_rn(a, a))
        case M_F_CBRT: AXB_UN_F(hwym::Cbrt(a))
        case M_F_RELU: AXB_UN_F(a > 0.0f ? a : 0.0f)
        case M_F_LEAKY: {
            const float slope = __uint_as_float((uint32_t)p.imm[ra]);
            AXB_UN_F(a > 0.0f ? a : __fmul_rn(slope, a))
        }
        case M_F_SILU: AXB_UN_F(hwym::Silu(a))
        case M_F_SIGMOID: AXB_UN_F(hwym::Sigmoid(a))
        case M_F_TANH: AXB_UN_F(hwym::Tanh(a))
        case M_F_GELU: AXB_UN_F(hwym::Gelu(a))
        case M_F_ISNAN: AXB_UN_F2I(a != a)
        case M_F_ISINF: AXB_UN_F2I(isinf(a))
        case M_F_ISFINITE: AXB_UN_F2I(isfinite(a))
        case M_F_LNOT: AXB_UN_F2I(!(a != 0.0f))
        // libm-accurate variants (half dtypes; reference computes these in double then rounds)
        case M_FA_SQRT: AXB_UN_F((float)sqrt((double)a))
        case M_FA_EXP: AXB_UN_F(slow::f_exp(a))
        case M_FA_LOG: AXB_UN_F(slow::f_log(a))
        case M_FA_SIN: AXB_UN_F(slow::f_sin(a))
        case M_FA_COS: AXB_UN_F(slow::f_cos(a))
        case M_FA_TAN: AXB_UN_F(slow::f_tan(a))
        case M_FA_ERF: AXB_UN_F(slow::f_erf(a))
        case M_FA_CBRT: AXB_UN_F(slow::f_cbrt(a))
        case M_FA_SILU: AXB_UN_F(slow::f_silu(a))
        case M_FA_SIGMOID: AXB_UN_F(slow::f_sigmoid(a))
        case M_FA_TANH: AXB_UN_F(slow::f_tanh(a))
        case M_FA_GELU: AXB_UN_F(slow::f_gelu(a))
        // ---- fp32 binary -------------------------------------------------------------------------
        case M_F_ADD: AXB_BIN_F(__fadd_rn(a, b))
        case M_F_SUB: AXB_BIN_F(__fsub_rn(a, b))
        case M_F_MUL: AXB_BIN_F(__fmul_rn(a, b))
        case M_F_DIV: AXB_BIN_F(__fdiv_rn(a, b))
        case M_F_POW: AXB_BIN_F(hwym::Pow(a, b))
        case M_F_MOD: AXB_BIN_F(slow::f_fmod(a, b))
        case M_F_MAX: AXB_BIN_F((a < b) ? b : a) // std::max (cpu_operations.hpp:337-341)
        case M_F_MIN: AXB_BIN_F((b < a) ? b : a) // std::min
        case M_F_ATAN2: AXB_BIN_F(slow::f_atan2(a, b))
        case M_F_HYPOT: AXB_BIN_F(hwym::Hypot(a, b))
        case M_FA_POW: AXB_BIN_F(slow::f_pow(a, b))
        case M_FA_HYPOT: AXB_BIN_F(slow::f_hypot(a, b))
        case M_F_EQ: AXB_CMP_F(a == b)
        case M_F_NE: AXB_CMP_F(a != b)
        case M_F_LT: AXB_CMP_F(a < b)
        case M_F_LE: AXB_CMP_F(a <= b)
        case M_F_GT: AXB_CMP_F(a > b)
        case M_F_GE: AXB_CMP_F(a >= b)
        case M_F_LAND: AXB_CMP_F((a != 0.0f) && (b != 0.0f))
        case M_F_LOR: AXB_CMP_F((a != 0.0f) || (b != 0.0f))
        case M_F_LXOR: AXB_CMP_F((a != 0.0f) != (b != 0.0f))
        // ---- int32 ---------------------------------------------------------------------------------
        case M_I_NEG: AXB_UN_I(0u - ua)
        case M_I_ABS: AXB_UN_I(a < 0 ? (int32_t)(0u - ua) : a)
        case M_I_SIGN: AXB_UN_I(a > 0 ? 1 : (a < 0 ? -1 : 0))
        case M_I_SQUARE: AXB_UN_I(ua * ua)
        case M_I_RELU: AXB_UN_I(a > 0 ? a : 0)
        case M_I_RECIP: AXB_UN_I(a != 0 ? i_div(1, a) : 0)
        case M_U_RECIP: AXB_UN_I(ua != 0u ? 1u / ua : 0u)
        case M_I_LNOT: AXB_UN_I(ua == 0u ? 1 : 0)
        case M_I_ADD: AXB_BIN_I(ua + ub)
        case M_I_SUB: AXB_BIN_I(ua - ub)
        case M_I_MUL: AXB_BIN_I(ua * ub)
        case M_I_DIV: AXB_BIN_I(i_div(a, b))
        case M_U_DIV: AXB_BIN_I(ub == 0u ? 0u : ua / ub)
        case M_I_MOD: AXB_BIN_I(i_mod(a, b))
        case M_U_MOD: AXB_BIN_I(ub == 0u ? 0u : ua % ub)
        case M_I_MAX: AXB_BIN_I(a < b ? b : a)
        case M_U_MAX: AXB_BIN_I(ua < ub ? ub : ua)
        case M_I_MIN: AXB_BIN_I(b < a ? b : a)
        case M_U_MIN: AXB_BIN_I(ub < ua ? ub : ua)
        case M_I_AND: AXB_BIN_I(ua & ub)
        case M_I_OR: AXB_BIN_I(ua | ub)
        case M_I_XOR: AXB_BIN_I(ua ^ ub)
        case M_I_SHL: AXB_BIN_I(ua << (ub & 31u))
        case M_I_SHR: AXB_BIN_I(a >> (ub & 31u))
        case M_U_SHR: AXB_BIN_I(ua >> (ub & 31u))
        // ---- fp64 ----------------------------------------------------------------------------------
        case M_D_NEG: AXB_UN_D(-a)
        case M_D_ABS: AXB_UN_D(fabs(a))
        case M_D_SQRT: AXB_UN_D(sqrt(a))
        case M_D_EXP: AXB_UN_D(slow::d_exp(a))
        case M_D_LOG: AXB_UN_D(slow::d_log(a))
        case M_D_SIN: AXB_UN_D(slow::d_sin(a))
        case M_D_COS: AXB_UN_D(slow::d_cos(a))
        case M_D_TAN: AXB_UN_D(slow::d_tan(a))
        case M_D_TANH: AXB_UN_D(slow::d_tanh(a))
        case M_D_ERF: AXB_UN_D(slow::d_erf(a))
        case M_D_SIGN: AXB_UN_D(a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : 0.0))
        case M_D_FLOOR: AXB_UN_D(floor(a))
        case M_D_CEIL: AXB_UN_D(ceil(a))
        case M_D_TRUNC: AXB_UN_D(trunc(a))
        case M_D_ROUND: AXB_UN_D(rint(a))
        case M_D_RECIP: AXB_UN_D(1.0 / a)
        case M_D_SQUARE: AXB_UN_D(a * a)
        case M_D_RELU: AXB_UN_D(a > 0.0 ? a : 0.0)
        case M_D_SIGMOID: AXB_UN_D(slow::d_sigmoid(a))
        case M_D_ISNAN: AXB_UN_D2I(a != a)
        case M_D_ISINF: AXB_UN_D2I(isinf(a))
        case M_D_ISFINITE: AXB_UN_D2I(isfinite(a))
        case M_D_LNOT: AXB_UN_D2I(!(a != 0.0))
        case M_D_ADD: AXB_BIN_D(__dadd_rn(a, b))
        case M_D_SUB: AXB_BIN_D(__dsub_rn(a, b))
        case M_D_MUL: AXB_BIN_D(__dmul_rn(a, b))
        case M_D_DIV: AXB_BIN_D(__ddiv_rn(a, b))
        case M_D_POW: AXB_BIN_D(slow::d_pow(a, b))
        case M_D_MOD: AXB_BIN_D(slow::d_fmod(a, b))
        case M_D_MAX: AXB_BIN_D((a < b) ? b : a)
        case M_D_MIN: AXB_BIN_D((b < a) ? b : a)
        case M_D_HYPOT: AXB_BIN_D(slow::d_hypot(a, b))
        case M_D_ATAN2: AXB_BIN_D(slow::d_atan2(a, b))
        // ---- int64 ---------------------------------------------------------------------------------
        case M_L_NEG: AXB_UN_L(0ull - (uint64_t)a)
        case M_L_ABS: AXB_UN_L(a < 0 ? (int64_t)(0ull - (uint64_t)a) : a)
        case M_L_SIGN: AXB_UN_L(a > 0 ? 1 : (a < 0 ? -1 : 0))
        case M_L_SQUARE: AXB_UN_L((uint64_t)a * (uint64_t)a)
        case M_L_RELU: AXB_UN_L(a > 0 ? a : 0)
        case M_L_LNOT: AXB_UN_L(a == 0 ? 1 : 0)
        case M_L_ADD: AXB_BIN_L(ua + ub)
        case M_L_SUB: AXB_BIN_L(ua - ub)
        case M_L_MUL: AXB_BIN_L(ua * ub)
        case M_L_DIV: AXB_BIN_L(l_div(a, b))
        case M_UL_DIV: AXB_BIN_L(ub == 0 ? 0ull : ua / ub)
        case M_L_MOD: AXB_BIN_L(l_mod(a, b))
        case M_UL_MOD: AXB_BIN_L(ub == 0 ? 0ull : ua % ub)
        case M_L_MAX: AXB_BIN_L(a < b ? b : a)
        case M_UL_MAX: AXB_BIN_L(ua < ub ? ub : ua)
        case M_L_MIN: AXB_BIN_L(b < a ? b : a)
        case M_UL_MIN: AXB_BIN_L(ub < ua ? ub : ua)
        case M_L_AND: AXB_BIN_L(ua & ub)
        case M_L_OR: AXB_BIN_L(ua | ub)
        case M_L_XOR: AXB_BIN_L(ua ^ ub)
        case M_L_SHL: AXB_BIN_L(ua << (ub & 63u))
        case M_L_SHR: AXB_BIN_L(a >> (ub & 63u))
        case M_UL_SHR: AXB_BIN_L(ua >> (ub & 63u))
        default: break;
        }
        // post-op rounding of acc to the instruction's half dtype
        const uint32_t rnd = fl & 3u;
        if (rnd == RND_BF16) {
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                acc[i] = fromF<W>(round_bf16(asF<W>(acc[i])));
        } else if (rnd == RND_F16) {
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                acc[i] = fromF<W>(round_f16(asF<W>(acc[i])));
        }
    }
}

template <int VEC, typename W>
__global__ void __launch_bounds__(256, 3) elementwise_kernel(const __grid_constant__ Params p) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t vi = blockIdx.x * blockDim.x + threadIdx.x; vi < p.total_vec; vi += stride) {
        uint32_t row = fastdiv(vi, p.div_m[0], p.div_s[0]);
        uint32_t col = (vi - row * p.nvec_inner) * VEC;
        int valid = (int)min((uint32_t)VEC, p.inner_len - col);
        uint32_t c[kMaxOuter];
        uint32_t rem = row;
#pragma unroll
        for (int d = 0; d < kMaxOuter; ++d) {
            c[d] = 0;
            if (d < (int)p.n_outer) {
                uint32_t q = fastdiv(rem, p.div_m[1 + d], p.div_s[1 + d]);
                c[d] = rem - q * p.outer_shape[d];
                rem = q;
            }
        }
        W acc[VEC];
        run_program<VEC, W>(p, nullptr, c, col, valid, acc, true);
    }
}

} // namespace ew
} // namespace axb
