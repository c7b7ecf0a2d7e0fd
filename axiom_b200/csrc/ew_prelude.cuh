// ew_prelude.cuh — device side of kernel family 1: the fused elementwise / broadcast / where chain.
//
// One kernel runs a whole fusion group. The host (elementwise.cu) lowers the typed
// axb_instr program to a straight-line sequence of the EW_* statements defined here and the resulting
// translation unit (this prelude + ~20 generated lines) is specialised per program SIGNATURE (ops, dtypes,
// operand access modes — never shapes or immediates) by NVRTC for sm_100a, cached as a cubin in-tree
// (axiom_b200/kernel_cache/) and in memory. build() pre-compiles the signatures of the BASELINE configs, so
// the named hot path never compiles at run time. Everything is static in the specialised kernel: every thread
// owns VEC consecutive output elements of one innermost run and keeps the chain's live values in registers.
// Memory side: inputs whose innermost run is contiguous and 16-byte aligned are read with 128-bit streaming
// loads, broadcast inputs with one scalar load, anything else with per-lane strided loads; the output is written
// with 128-bit stores. HBM-bound by construction: every operand byte is read once.
//
// Register model (per thread): acc, t, u and NSLOT save slots, each VEC words wide.
//   canonical value forms   F: fp32 (bf16/fp16 values are fp32 values already rounded to that format)
//                           I: 32-bit integer (narrow ints sign/zero-extended, bool = 0/1)
//                           D: fp64, L: 64-bit integer          (only in the 64-bit-word instantiation)
//
// This file must stay self-contained (no system headers under NVRTC): it is embedded into libaxb.so as a string.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#ifndef INT32_MIN
#define INT32_MIN (-2147483647 - 1)
#endif
#ifndef INT64_MIN
#define INT64_MIN (-9223372036854775807LL - 1)
#endif
#endif

#include "hwy_math.cuh"

#ifndef AXB_H
// == axb_dtype (include/axb.h) == axiom::DType
enum { AXB_BOOL = 0, AXB_I8 = 1, AXB_I16 = 2, AXB_I32 = 3, AXB_I64 = 4, AXB_U8 = 5, AXB_U16 = 6, AXB_U32 = 7,
       AXB_U64 = 8, AXB_F16 = 9, AXB_BF16 = 10, AXB_F32 = 11, AXB_F64 = 12, AXB_C64 = 13, AXB_C128 = 14 };
#endif

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
    // iteration space: innermost run of `inner_len` elements, n_outer collapsed outer dims
    uint32_t inner_len;
    uint32_t nvec_inner; // ceil(inner_len / VEC)
    uint32_t total_vec;  // nvec_inner * prod(outer)  (< 2^31)
    uint32_t n_outer;
    uint32_t outer_shape[kMaxOuter];
    uint32_t div_m[kMaxOuter + 1]; // fast-div magics: [0] = nvec_inner, [1+d] = outer_shape[d]
    uint32_t div_s[kMaxOuter + 1];
    int32_t pad0;
    int32_t reduce_op; // fused full reduction (axb_elementwise_reduce): 0 = none
    void *reduce_ws;   // per-block partials
    const float *aux0, *aux1; // fused layer_norm / rms_norm: weight and bias rows (fp32, contiguous) or nullptr
    float eps;
    int32_t pad1;
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t m, uint32_t s) {
    return (__umulhi(n, m) + n) >> s;
}

__device__ __forceinline__ float h2f(unsigned short h) {
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f;
}
__device__ __forceinline__ unsigned short f2h(float f) {
    unsigned short h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(f));
    return h;
}
// ---- canonicalisation of storage bits -------------------------------------------------------------
template <typename W> __device__ __forceinline__ W canon(uint64_t raw, int dt) {
    switch (dt) {
    case AXB_F32: case AXB_I32: case AXB_U32: return (W)(uint32_t)raw;
    case AXB_BF16: return (W)((uint32_t)raw << 16);
    case AXB_F16: return (W)__float_as_uint(h2f((unsigned short)raw));
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
__device__ __forceinline__ float round_f16(float f) { return h2f(f2h(f)); }
template <typename W, int VEC> __device__ __forceinline__ void round_bf16_vec(W (&acc)[VEC]) {
    bool any_nan = false;
#pragma unroll
    for (int i = 0; i + 1 < VEC; i += 2) {
        const float a0 = __uint_as_float((uint32_t)acc[i]), a1 = __uint_as_float((uint32_t)acc[i + 1]);
        any_nan |= (a0 != a0) | (a1 != a1);
        uint32_t h;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a1), "f"(a0));
        acc[i] = (W)(h << 16);
        acc[i + 1] = (W)(h & 0xFFFF0000u);
    }
    if (VEC & 1)
        acc[VEC - 1] = (W)__float_as_uint(round_bf16(__uint_as_float((uint32_t)acc[VEC - 1])));
    if (any_nan) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float a = __uint_as_float((uint32_t)acc[i]);
            if (a != a)
                acc[i] = (W)0x7FC00000u;
        }
    }
}

template <typename W> __device__ __forceinline__ uint64_t decanon(W v, int dt) {
    switch (dt) {
    case AXB_BF16: return f32_to_bf16_bits(__uint_as_float((uint32_t)v));
    case AXB_F16: return f2h(__uint_as_float((uint32_t)v));
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
__device__ __host__ constexpr int dsize(int dt) {
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

__device__ __forceinline__ uint32_t ldg32(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg32(void *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Load VEC elements of operand `o` starting at element column `col` of the current run.
// MODE_VEC: the VEC*sizeof(elem) bytes (4, 8, 16 or 32) are contiguous and aligned -> one (two for 32 B) streaming
// vector load; consecutive threads own consecutive vectors, so a warp instruction covers one contiguous span.
template <int VEC, typename W, int DT, int MODE>
__device__ __forceinline__ void load_vec(const Operand &o, int64_t outer_off, uint32_t col, int valid,
                                         W (&r)[VEC]) {
    constexpr int dt = DT;
    constexpr int sz = dsize(dt);
    constexpr int BYTES = VEC * sz;
    constexpr int NW = BYTES >= 4 ? BYTES / 4 : 1;
    const char *base = o.ptr + outer_off;
    if (MODE == MODE_BCAST) {
        W v = canon<W>(ld_raw(base, sz), dt);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            r[i] = v;
        return;
    }
    if (MODE == MODE_VEC && valid == VEC && (BYTES == 4 || BYTES == 8 || BYTES == 16 || BYTES == 32)) {
        const char *p = base + (int64_t)col * sz;
        uint32_t w[NW];
        if (BYTES == 32) {
            uint4 a = ldg128(p), b = ldg128(p + 16);
            w[0] = a.x; w[1 % NW] = a.y; w[2 % NW] = a.z; w[3 % NW] = a.w;
            w[4 % NW] = b.x; w[5 % NW] = b.y; w[6 % NW] = b.z; w[7 % NW] = b.w;
        } else if (BYTES == 16) {
            uint4 a = ldg128(p);
            w[0] = a.x; w[1 % NW] = a.y; w[2 % NW] = a.z; w[3 % NW] = a.w;
        } else if (BYTES == 8) {
            uint2 a = ldg64(p);
            w[0] = a.x; w[1 % NW] = a.y;
        } else {
            w[0] = ldg32(p);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            uint64_t raw;
            if (sz == 8)
                raw = ((uint64_t)w[(2 * i + 1) % NW] << 32) | w[(2 * i) % NW];
            else if (sz == 4)
                raw = w[i % NW];
            else if (sz == 2)
                raw = (w[(i / 2) % NW] >> (16 * (i % 2))) & 0xFFFFu;
            else
                raw = (w[(i / 4) % NW] >> (8 * (i % 4))) & 0xFFu;
            r[i] = canon<W>(raw, dt);
        }
        return;
    }
    const int64_t is = (MODE == MODE_VEC) ? (int64_t)sz : o.inner_stride;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        r[i] = 0;
        if (i < valid)
            r[i] = canon<W>(ld_raw(base + (int64_t)(col + i) * is, sz), dt);
    }
}

template <int VEC, typename W, int DT, int MODE>
__device__ __forceinline__ void store_vec(const Operand &o, int64_t outer_off, uint32_t col, int valid,
                                          const W (&r)[VEC]) {
    constexpr int dt = DT;
    constexpr int sz = dsize(dt);
    constexpr int BYTES = VEC * sz;
    constexpr int NW = BYTES >= 4 ? BYTES / 4 : 1;
    char *base = const_cast<char *>(o.ptr) + outer_off;
    if (MODE == MODE_VEC && valid == VEC && (BYTES == 4 || BYTES == 8 || BYTES == 16 || BYTES == 32)) {
        char *p = base + (int64_t)col * sz;
        uint32_t w[NW];
        if (dt == AXB_BF16 && (VEC % 2) == 0) {
            // every bf16-typed value of a program is already rounded (low 16 bits clear, NaN canonical): the store takes
            // the high halves (one PRMT per pair). A value that is NOT pre-rounded sends the vector through the exact
            // conversion instead, so the shortcut cannot change a result.
            uint32_t low = 0;
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                low |= (uint32_t)r[i];
            if ((low & 0xFFFFu) == 0u) {
#pragma unroll
                for (int i = 0; i < VEC; i += 2)
                    w[(i / 2) % NW] = __byte_perm((uint32_t)r[i], (uint32_t)r[i + 1], 0x7632);
                if (BYTES == 32) {
                    stg128(p, make_uint4(w[0], w[1 % NW], w[2 % NW], w[3 % NW]));
                    stg128(p + 16, make_uint4(w[4 % NW], w[5 % NW], w[6 % NW], w[7 % NW]));
                } else if (BYTES == 16) {
                    stg128(p, make_uint4(w[0], w[1 % NW], w[2 % NW], w[3 % NW]));
                } else if (BYTES == 8) {
                    stg64(p, make_uint2(w[0], w[1 % NW]));
                } else {
                    stg32(p, w[0]);
                }
                return;
            }
        }
#pragma unroll
        for (int k = 0; k < NW; ++k)
            w[k] = 0;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const uint64_t raw = decanon<W>(r[i], dt);
            if (sz == 8) {
                w[(2 * i) % NW] = (uint32_t)raw;
                w[(2 * i + 1) % NW] = (uint32_t)(raw >> 32);
            } else if (sz == 4) {
                w[i % NW] = (uint32_t)raw;
            } else if (sz == 2) {
                w[(i / 2) % NW] |= ((uint32_t)raw & 0xFFFFu) << (16 * (i % 2));
            } else {
                w[(i / 4) % NW] |= ((uint32_t)raw & 0xFFu) << (8 * (i % 4));
            }
        }
        if (BYTES == 32) {
            stg128(p, make_uint4(w[0], w[1 % NW], w[2 % NW], w[3 % NW]));
            stg128(p + 16, make_uint4(w[4 % NW], w[5 % NW], w[6 % NW], w[7 % NW]));
        } else if (BYTES == 16) {
            stg128(p, make_uint4(w[0], w[1 % NW], w[2 % NW], w[3 % NW]));
        } else if (BYTES == 8) {
            stg64(p, make_uint2(w[0], w[1 % NW]));
        } else {
            stg32(p, w[0]);
        }
        return;
    }
    const int64_t is = (MODE == MODE_VEC) ? (int64_t)sz : o.inner_stride;
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
// Half-type transcendentals without the double-precision detour. The reference evaluates bf16 / fp16 unary functors in
// DOUBLE and rounds once to the half type (cpu_operations.hpp:432-710: the non-is_floating_point branches); slow:: above
// restates that literally and is FP64-pipe bound (bf16 gelu 0.75 TB/s). hfast:: first evaluates the function in fp32
// with a known error bound (CUDA libm: expf/sinf/cosf/erff/tanhf <= 2 ulp, logf/cbrtf 1, tanf/erfcf 4) and asks whether
// that error could change the rounding to the half type: it can only if the fp32 value lies within MARGIN ulps of a
// rounding boundary of the half format (low B bits == 1 << (B-1); B = 16 for bf16, 13 for fp16), or outside the range
// where the half format is normal. Only then (0.04 % - 0.3 % of the elements) the literal double path runs, so the
// result is bit-identical to the reference on EVERY input — tests/test_elementwise_gpu.py checks all 65 536 bit patterns
// of both half types for each function.
namespace hfast {
template <int B> __device__ __forceinline__ bool unsafe(float y, uint32_t margin) {
    const uint32_t u = __float_as_uint(y);
    const uint32_t e = (u >> 23) & 0xFFu;
    const uint32_t lo = u & ((1u << B) - 1u);
    const uint32_t mid = 1u << (B - 1);
    const uint32_t d = lo > mid ? lo - mid : mid - lo;
    // bf16 shares fp32's exponent range (fp32 subnormals / inf / nan go the slow way); fp16 is normal for 2^-14 <= |y| < 2^16
    const bool range_ok = B == 16 ? (e != 0u && e != 255u) : (e >= 113u && e <= 142u);
    return d <= margin || !range_ok;
}
#define AXB_HFAST1(NAME, FAST, SLOW, MARGIN)                                                            \
    template <int B> __device__ __forceinline__ float NAME(float a) {                                   \
        float y = (FAST);                                                                               \
        if (unsafe<B>(y, MARGIN))                                                                       \
            y = slow::SLOW(a);                                                                          \
        return y;                                                                                       \
    }
AXB_HFAST1(Exp, expf(a), f_exp, 12u)
AXB_HFAST1(Log, logf(a), f_log, 12u)
AXB_HFAST1(Sin, sinf(a), f_sin, 12u)
AXB_HFAST1(Cos, cosf(a), f_cos, 12u)
AXB_HFAST1(Tan, tanf(a), f_tan, 12u)
AXB_HFAST1(Erf, erff(a), f_erf, 12u)
AXB_HFAST1(Cbrt, cbrtf(a), f_cbrt, 12u)
AXB_HFAST1(Tanh, tanhf(a), f_tanh, 12u)
#undef AXB_HFAST1
// sigmoid / silu: the reference's own formula (exp(-|x|), 1 + e, one division) in fp32: <= 4.5 ulp
template <int B> __device__ __forceinline__ float Sigmoid(float a) {
    const float e = expf(-fabsf(a)), den = __fadd_rn(1.0f, e);
    float y = a >= 0.0f ? __fdiv_rn(1.0f, den) : __fdiv_rn(e, den);
    if (unsafe<B>(y, 12u))
        y = slow::f_sigmoid(a);
    return y;
}
template <int B> __device__ __forceinline__ float Silu(float a) {
    const float e = expf(-fabsf(a)), den = __fadd_rn(1.0f, e);
    float y = a >= 0.0f ? __fdiv_rn(a, den) : __fdiv_rn(__fmul_rn(a, e), den);
    if (unsafe<B>(y, 12u) || a == 0.0f)
        y = slow::f_silu(a);
    return y;
}
// gelu = 0.5 x (1 + erf(x / sqrt 2)) = 0.5 x erfc(-x / sqrt 2): erfc keeps the relative accuracy the 1 + erf form loses
// for negative x. Error: 4 ulp (erfcf) + 2 (products) + 1.5 x^2 for x < 0 (the rounding of the argument, amplified by
// z erfc'(z) / erfc(z) ~ 2 z^2) — under 44 ulp for x >= -5, doubled for a binade edge: margin 96. Below -5: slow path.
template <int B> __device__ __forceinline__ float Gelu(float a) {
    float y = __fmul_rn(__fmul_rn(0.5f, a), erfcf(__fmul_rn(a, -0.70710678118654752440f)));
    if (unsafe<B>(y, 96u) || a < -5.0f || a == 0.0f)
        y = slow::f_gelu(a);
    return y;
}
} // namespace hfast
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
    }                                                                                                   
#define AXB_UN_F2I(EXPR)                                                                                \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            float a = asF<W>(acc[i]);                                                                   \
            acc[i] = fromI<W>((EXPR) ? 1 : 0);                                                          \
        }                                                                                               \
    }                                                                                                   
#define AXB_BIN_F(EXPR)                                                                                 \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            float a = asF<W>(swap ? t[i] : acc[i]);                                                     \
            float b = asF<W>(swap ? acc[i] : t[i]);                                                     \
            acc[i] = fromF<W>(EXPR);                                                                    \
        }                                                                                               \
    }                                                                                                   
#define AXB_CMP_F(EXPR)                                                                                 \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            float a = asF<W>(swap ? t[i] : acc[i]);                                                     \
            float b = asF<W>(swap ? acc[i] : t[i]);                                                     \
            acc[i] = fromI<W>((EXPR) ? 1 : 0);                                                          \
        }                                                                                               \
    }                                                                                                   
#define AXB_UN_I(EXPR)                                                                                  \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            int32_t a = asI<W>(acc[i]);                                                                 \
            uint32_t ua = (uint32_t)a;                                                                  \
            (void)ua;                                                                                   \
            acc[i] = fromI<W>((int32_t)(EXPR));                                                         \
        }                                                                                               \
    }                                                                                                   
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
    }                                                                                                   
#define AXB_UN_D(EXPR)                                                                                  \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                double a = asD(acc[i]);                                                                 \
                acc[i] = fromD(EXPR);                                                                   \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   
#define AXB_UN_D2I(EXPR)                                                                                \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                double a = asD(acc[i]);                                                                 \
                acc[i] = (W)((EXPR) ? 1 : 0);                                                           \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   
#define AXB_BIN_D(EXPR)                                                                                 \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                double a = asD(swap ? t[i] : acc[i]);                                                   \
                double b = asD(swap ? acc[i] : t[i]);                                                   \
                acc[i] = fromD(EXPR);                                                                   \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   
#define AXB_UN_L(EXPR)                                                                                  \
    {                                                                                                   \
        if constexpr (sizeof(W) == 8) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                int64_t a = (int64_t)acc[i];                                                            \
                acc[i] = (W)(int64_t)(EXPR);                                                            \
            }                                                                                           \
        }                                                                                               \
    }                                                                                                   
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
    }                                                                                                   
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
    }                                                                                                   
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
    }                                                                                                   


// ---- statements emitted by the host code generator (elementwise.cu) --------------------------------------
// A specialised kernel is two lambdas over a per-vector context `x` (outer coordinates, column, valid count and
// the raw input vectors):  preload(x) issues every global load of one vector, compute(x) runs the chain and stores.
// ew_run() calls preload for U vectors before the first compute, so U x (inputs) 128-bit loads are in flight per thread.
template <typename W, int VEC, int NIN> struct Ctx {
    uint32_t c[kMaxOuter];
    uint32_t col;
    int valid;
    bool ok;
    W in[NIN > 0 ? NIN : 1][VEC];
};
template <int NOUT> __device__ __forceinline__ int64_t outer_off(const Operand &o, const uint32_t (&c)[kMaxOuter]) {
    int64_t off = 0;
#pragma unroll
    for (int d = 0; d < NOUT; ++d)
        off += (int64_t)c[d] * o.stride[d];
    return off;
}
#define EW_PRELOAD(K, DT, MODE) load_vec<VEC, W, DT, MODE>(p.in[K], outer_off<NOUT>(p.in[K], x.c), x.col, x.valid, x.in[K]);
#define EW_USE(K, DST)                                                                                  \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) DST[i] = x.in[K][i];                            \
    }
#define EW_STORE(DT, MODE) store_vec<VEC, W, DT, MODE>(p.out, outer_off<NOUT>(p.out, x.c), x.col, x.valid, acc);
#define EW_IMM(K, DST)                                                                                  \
    {                                                                                                   \
        const W v_ = (W)p.imm[K];                                                                       \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) DST[i] = v_;                                    \
    }
#define EW_MOV(SRC, DST)                                                                                \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) DST[i] = SRC[i];                                \
    }
#define EW_SEL                                                                                          \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) acc[i] = (acc[i] != 0) ? t[i] : u[i];           \
    }
#define EW_LEAKY(K)                                                                                     \
    {                                                                                                   \
        const float slope = __uint_as_float((uint32_t)p.imm[K]);                                        \
        AXB_UN_F(a > 0.0f ? a : __fmul_rn(slope, a))                                                    \
    }
// bf16 rounding of the whole vector: hardware pair conversion (cvt.rn.bf16x2.f32: RNE, exact on subnormals / overflow)
// and the reference's NaN encoding 0x7FC0 (bfloat16.hpp:16-27) restored on the rare vector that holds a NaN — 2.5
// instructions per element instead of the 7 of the integer formulation (bf16 kernels were issue-bound at ~3.4 TB/s).
#define EW_RND_BF16 round_bf16_vec<W, VEC>(acc);
#define EW_RND_F16                                                                                      \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) acc[i] = fromF<W>(round_f16(asF<W>(acc[i])));   \
    }
// binary statements take the operand order as a literal: EW_OP(SWAP, stmt)
#define EW_OP(SWAP, STMT)                                                                               \
    {                                                                                                   \
        constexpr bool swap = SWAP;                                                                     \
        (void)swap;                                                                                     \
        STMT                                                                                            \
    }
// conversions act on register RA (R_T / R_ACC / R_U)
#define EW_CVT_ON(RA, STMT)                                                                             \
    {                                                                                                   \
        constexpr int ra = RA;                                                                          \
        STMT                                                                                            \
    }

// One CTA owns 256*U consecutive vectors (VEC elements of one innermost run each): thread t handles vectors
// t, t+256, ... so every warp-level access is one contiguous span, and all U x (inputs) loads are issued before the
// first use. No grid-stride loop: a large grid of short CTAs measured faster than a persistent one on B200
// (tools/stream_probe.cu, profiles/r1_stream_probe.txt: 6.8 TB/s vs 6.0 TB/s for y = relu(x)).
template <int VEC, typename W, int NOUT, int NIN, int U, class FL, class FC>
__device__ __forceinline__ void ew_run(const Params &p, FL &&preload, FC &&compute) {
    const uint32_t v0 = blockIdx.x * (256u * U) + threadIdx.x;
    Ctx<W, VEC, NIN> x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t vi = v0 + u * 256u;
        x[u].ok = vi < p.total_vec;
        if (x[u].ok) {
            uint32_t row = (NOUT > 0) ? fastdiv(vi, p.div_m[0], p.div_s[0]) : 0u;
            x[u].col = (vi - row * p.nvec_inner) * VEC;
            x[u].valid = (int)min((uint32_t)VEC, p.inner_len - x[u].col);
#pragma unroll
            for (int d = 0; d < kMaxOuter; ++d) {
                x[u].c[d] = 0;
                if (d < NOUT) {
                    if (d == NOUT - 1) {
                        x[u].c[d] = row;
                    } else {
                        uint32_t q = fastdiv(row, p.div_m[1 + d], p.div_s[1 + d]);
                        x[u].c[d] = row - q * p.outer_shape[d];
                        row = q;
                    }
                }
            }
            preload(x[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (x[u].ok)
            compute(x[u]);
}

#define EW_KERNEL_BEGIN(NAME, VEC_, WTYPE, NOUT_, NIN_, U_)                                             \
    extern "C" __global__ void __launch_bounds__(256) NAME(const __grid_constant__ axb::ew::Params p) { \
        using namespace axb::ew;                                                                        \
        constexpr int VEC = VEC_, NOUT = NOUT_, NIN = NIN_, U = U_;                                     \
        typedef WTYPE W;                                                                                \
        typedef Ctx<W, VEC, NIN> X;
#define EW_PRELOAD_BEGIN auto preload = [&](X &x) {
#define EW_PRELOAD_END                                                                                  \
    }                                                                                                   \
    ;
#define EW_COMPUTE_BEGIN                                                                                \
    auto compute = [&](X &x) {                                                                          \
        W acc[VEC], t[VEC], u[VEC], s0[VEC], s1[VEC], s2[VEC], s3[VEC];                                 \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                               \
            acc[i] = t[i] = u[i] = s0[i] = s1[i] = s2[i] = s3[i] = 0;                                   \
        }
#define EW_COMPUTE_END                                                                                  \
    }                                                                                                   \
    ;
#define EW_KERNEL_END                                                                                   \
        ew_run<VEC, W, NOUT, NIN, U>(p, preload, compute);                                              \
    }

// ---- chain + FULL reduction in one pass (FusedReductionStep, include/axiom/graph/compiled_graph.hpp:61-66) --------------
// The last statement of the program is EW_RED instead of EW_STORE: every thread folds the fp32 values of its vectors into
// one register, the CTA reduces (shuffles + one shared-memory hop) and writes ONE partial to p.out.ptr[blockIdx.x]; the
// partials are combined by the reduction family's finish pass. The chain's value never reaches HBM.
// OPK: 0 = sum (also mean), 1 = max, 2 = min — same combine rules as reduce_common.cuh SumOp / MaxOp / MinOp.
template <int OPK> __device__ __forceinline__ float red_identity() {
    return OPK == 0 ? 0.0f : OPK == 1 ? -3.402823466e+38f : 3.402823466e+38f;
}
template <int OPK> __device__ __forceinline__ float red_combine(float a, float b) {
    return OPK == 0 ? __fadd_rn(a, b) : OPK == 1 ? ((a < b) ? b : a) : ((b < a) ? b : a);
}
#define EW_RED(OPK_)                                                                                    \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) if (i < x.valid)                                \
            racc = red_combine<OPK_>(racc, asF<W>(acc[i]));                                             \
    }
#define EW_RKERNEL_BEGIN(NAME, VEC_, WTYPE, NOUT_, NIN_, U_, OPK_)                                      \
    extern "C" __global__ void __launch_bounds__(256) NAME(const __grid_constant__ axb::ew::Params p) { \
        using namespace axb::ew;                                                                        \
        constexpr int VEC = VEC_, NOUT = NOUT_, NIN = NIN_, U = U_, OPK = OPK_;                         \
        typedef WTYPE W;                                                                                \
        typedef Ctx<W, VEC, NIN> X;                                                                     \
        float racc = red_identity<OPK>();
#define EW_RKERNEL_END                                                                                  \
        ew_run<VEC, W, NOUT, NIN, U>(p, preload, compute);                                              \
        __shared__ float red_sm[8];                                                                     \
        _Pragma("unroll") for (int d = 16; d >= 1; d >>= 1)                                             \
            racc = red_combine<OPK>(racc, __shfl_xor_sync(0xFFFFFFFFu, racc, d));                       \
        if ((threadIdx.x & 31) == 0)                                                                    \
            red_sm[threadIdx.x >> 5] = racc;                                                            \
        __syncthreads();                                                                                \
        if (threadIdx.x == 0) {                                                                         \
            float tot = red_sm[0];                                                                      \
            _Pragma("unroll") for (int w = 1; w < 8; ++w) tot = red_combine<OPK>(tot, red_sm[w]);       \
            reinterpret_cast<float *>(const_cast<char *>(p.out.ptr))[blockIdx.x] = tot;                 \
        }                                                                                               \
    }


// ---- chain + LAST-AXIS reduction in one pass (the reference fuses elementwise producers into ANY reduction,
// src/graph/graph_compiler.cpp:376-478; `(a * b).sum({1})`, `x.square().mean({-1})`) ------------------------------------------
// The iteration space is planned so that the innermost run IS the reduced axis (elementwise.cu launch_chain_rowreduce). A
// group of G threads owns one output row: G = 256 (one CTA per row) for long rows, G = 32 (one warp per row, 8 rows per CTA,
// no block barrier) for rows of at most 32 * U vectors. Thread t folds vectors t, t + G, ... of the row through the same
// EW_RED statement as the full reduction, the group combines, and ONE fp32 result per row is written (Mean divides by the
// row length in fp32, like reduce_common.cuh finalize). The chain's value never reaches HBM.
template <int VEC, typename W, int NOUT, int NIN, int U, int G, class FL, class FC>
__device__ __forceinline__ void ew_run_row(const Params &p, uint32_t row, uint32_t t, FL &&preload, FC &&compute) {
    uint32_t c[kMaxOuter];
    {
        uint32_t r = row;
#pragma unroll
        for (int d = 0; d < kMaxOuter; ++d) {
            c[d] = 0;
            if (d < NOUT) {
                if (d == NOUT - 1) {
                    c[d] = r;
                } else {
                    uint32_t q = fastdiv(r, p.div_m[1 + d], p.div_s[1 + d]);
                    c[d] = r - q * p.outer_shape[d];
                    r = q;
                }
            }
        }
    }
    for (uint32_t v0 = t; v0 < p.nvec_inner; v0 += (uint32_t)(G * U)) {
        Ctx<W, VEC, NIN> x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t vi = v0 + (uint32_t)(u * G);
            x[u].ok = vi < p.nvec_inner;
            if (x[u].ok) {
                x[u].col = vi * VEC;
                x[u].valid = (int)min((uint32_t)VEC, p.inner_len - x[u].col);
#pragma unroll
                for (int d = 0; d < kMaxOuter; ++d)
                    x[u].c[d] = c[d];
                preload(x[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (x[u].ok)
                compute(x[u]);
    }
}
#define EW_ROWKERNEL_BEGIN(NAME, VEC_, WTYPE, NOUT_, NIN_, U_, OPK_, G_)                                \
    extern "C" __global__ void __launch_bounds__(256) NAME(const __grid_constant__ axb::ew::Params p) { \
        using namespace axb::ew;                                                                        \
        constexpr int VEC = VEC_, NOUT = NOUT_, NIN = NIN_, U = U_, OPK = OPK_, G = G_;                 \
        typedef WTYPE W;                                                                                \
        typedef Ctx<W, VEC, NIN> X;                                                                     \
        float racc = red_identity<OPK>();
#define EW_ROWKERNEL_END                                                                                \
        const uint32_t rows = p.total_vec / p.nvec_inner;                                               \
        const uint32_t row = G == 256 ? blockIdx.x : blockIdx.x * 8u + (threadIdx.x >> 5);              \
        const uint32_t tig = G == 256 ? threadIdx.x : (threadIdx.x & 31u);                              \
        if (row < rows)                                                                                 \
            ew_run_row<VEC, W, NOUT, NIN, U, G>(p, row, tig, preload, compute);                         \
        _Pragma("unroll") for (int d = 16; d >= 1; d >>= 1)                                             \
            racc = red_combine<OPK>(racc, __shfl_xor_sync(0xFFFFFFFFu, racc, d));                       \
        float *rout = reinterpret_cast<float *>(p.reduce_ws);                                           \
        if (G == 256) {                                                                                 \
            __shared__ float red_sm[8];                                                                 \
            if ((threadIdx.x & 31) == 0)                                                                \
                red_sm[threadIdx.x >> 5] = racc;                                                        \
            __syncthreads();                                                                            \
            if (threadIdx.x == 0) {                                                                     \
                float tot = red_sm[0];                                                                  \
                _Pragma("unroll") for (int w = 1; w < 8; ++w) tot = red_combine<OPK>(tot, red_sm[w]);   \
                if (p.reduce_op == 57 /* AXB_OP_MEAN */)                                                         \
                    tot = __fdiv_rn(tot, (float)p.inner_len);                                           \
                rout[row] = tot;                                                                        \
            }                                                                                           \
        } else if (tig == 0 && row < rows) {                                                            \
            rout[row] = p.reduce_op == 57 /* AXB_OP_MEAN */ ? __fdiv_rn(racc, (float)p.inner_len) : racc;        \
        }                                                                                               \
    }


// ---- chain -> softmax / log_softmax over the LAST axis [-> row sum / mean / max] in one pass ("softmax
// with elementwise prologue / epilogue"; `(x * s + m).softmax(-1)`, and the README chain's `.softmax(-1).sum({1})`) -----------
// Same row ownership as EW_ROWKERNEL (G = 256: one CTA per row, G = 32: one warp per row), but the chain's values STAY in
// registers: thread t computes vectors t, t + G, ..., t + (NV - 1) G of its row (NV <= 8, i.e. rows of at most G * NV * VEC
// elements; longer rows run chain and softmax as two launches), the group reduces max and sum(exp(x - max)) and the thread
// writes exp(x - max) / sum — or (x - max) - log(sum) — for its own elements. EPI: 0 = store the row, 1 = sum, 2 = mean,
// 3 = max of the softmax values, one fp32 per row. Numerics = the reference's op sequence (cpu_operations.cpp:3132-3158:
// max, exp(x - max), sum, divide) with expf / logf and a tree instead of the serial sum, inside the 1e-5 contract of the
// stand-alone softmax kernels.
#define EW_KEEP                                                                                         \
    {                                                                                                   \
        _Pragma("unroll") for (int i = 0; i < VEC; ++i) if (i < x.valid) kv[i] = asF<W>(acc[i]);        \
    }
template <int G> __device__ __forceinline__ float group_max(float v, float *sm) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1)
        v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    if (G == 256) {
        __syncthreads(); // the scratch may still be read by the previous reduction
        if ((threadIdx.x & 31) == 0)
            sm[threadIdx.x >> 5] = v;
        __syncthreads();
        v = sm[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            v = fmaxf(v, sm[w]);
    }
    return v;
}
template <int G> __device__ __forceinline__ float group_sum(float v, float *sm) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1)
        v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    if (G == 256) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0)
            sm[threadIdx.x >> 5] = v;
        __syncthreads();
        v = sm[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            v = __fadd_rn(v, sm[w]);
    }
    return v;
}
// MODE: 0 softmax, 1 log_softmax, 2 layer_norm (mean, centre, biased variance, / sqrt(var + eps), * w + b: the reference's op
// sequence, operations.cpp:1803-1817), 3 rms_norm (x / sqrt(mean(x^2) + eps) * w, :1841-1857); w / b = p.aux0 / p.aux1.
#define EW_SMKERNEL_BEGIN(NAME, VEC_, WTYPE, NOUT_, NIN_, NV_, G_, MODE_, EPI_)                         \
    extern "C" __global__ void __launch_bounds__(256) NAME(const __grid_constant__ axb::ew::Params p) { \
        using namespace axb::ew;                                                                        \
        constexpr int VEC = VEC_, NOUT = NOUT_, NIN = NIN_, NV = NV_, G = G_, MODE = MODE_, EPI = EPI_; \
        constexpr int LOG = MODE_ == 1;                                                                 \
        typedef WTYPE W;                                                                                \
        typedef Ctx<W, VEC, NIN> X;                                                                     \
        float keep[NV][VEC];                                                                            \
        float *kv = keep[0];
#define EW_SMKERNEL_END                                                                                 \
        __shared__ float red_sm[8];                                                                     \
        const float ninf = __int_as_float(0xff800000);                                                  \
        const uint32_t rows = p.total_vec / p.nvec_inner;                                               \
        const uint32_t row = G == 256 ? blockIdx.x : blockIdx.x * 8u + (threadIdx.x >> 5);              \
        const uint32_t tig = G == 256 ? threadIdx.x : (threadIdx.x & 31u);                              \
        const bool active = row < rows;                                                                 \
        uint32_t rc[kMaxOuter];                                                                         \
        {                                                                                               \
            uint32_t r = row;                                                                           \
            _Pragma("unroll") for (int d = 0; d < kMaxOuter; ++d) {                                     \
                rc[d] = 0;                                                                              \
                if (d < NOUT) {                                                                         \
                    if (d == NOUT - 1) {                                                                \
                        rc[d] = r;                                                                      \
                    } else {                                                                            \
                        uint32_t q = fastdiv(r, p.div_m[1 + d], p.div_s[1 + d]);                        \
                        rc[d] = r - q * p.outer_shape[d];                                               \
                        r = q;                                                                          \
                    }                                                                                   \
                }                                                                                       \
            }                                                                                           \
        }                                                                                               \
        _Pragma("unroll") for (int j = 0; j < NV; ++j) {                                                \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) keep[j][i] = MODE >= 2 ? 0.0f : ninf;       \
            const uint32_t vi = tig + (uint32_t)(j * G);                                                \
            if (active && vi < p.nvec_inner) {                                                          \
                X x;                                                                                    \
                x.ok = true;                                                                            \
                x.col = vi * VEC;                                                                       \
                x.valid = (int)min((uint32_t)VEC, p.inner_len - x.col);                                 \
                _Pragma("unroll") for (int d = 0; d < kMaxOuter; ++d) x.c[d] = rc[d];                   \
                kv = keep[j];                                                                           \
                preload(x);                                                                             \
                compute(x);                                                                             \
            }                                                                                           \
        }                                                                                               \
        float *rout = reinterpret_cast<float *>(p.reduce_ws);                                           \
        if (MODE >= 2) { /* layer_norm / rms_norm: padded lanes hold 0 and are never stored */          \
            const float n_ = (float)p.inner_len;                                                        \
            float mean = 0.0f;                                                                          \
            if (MODE == 2) {                                                                            \
                float s1 = 0.0f;                                                                        \
                _Pragma("unroll") for (int j = 0; j < NV; ++j)                                          \
                    _Pragma("unroll") for (int i = 0; i < VEC; ++i) s1 = __fadd_rn(s1, keep[j][i]);     \
                mean = __fdiv_rn(group_sum<G>(s1, red_sm), n_);                                         \
            }                                                                                           \
            float s2 = 0.0f;                                                                            \
            _Pragma("unroll") for (int j = 0; j < NV; ++j) {                                            \
                const uint32_t vi = tig + (uint32_t)(j * G);                                            \
                const int valid = (active && vi < p.nvec_inner) ? (int)min((uint32_t)VEC, p.inner_len - vi * VEC) : 0; \
                _Pragma("unroll") for (int i = 0; i < VEC; ++i) if (i < valid) {                        \
                    const float c_ = __fsub_rn(keep[j][i], mean);                                       \
                    keep[j][i] = c_;                                                                    \
                    s2 = __fadd_rn(s2, __fmul_rn(c_, c_));                                              \
                }                                                                                       \
            }                                                                                           \
            const float sd = __fsqrt_rn(__fadd_rn(__fdiv_rn(group_sum<G>(s2, red_sm), n_), p.eps));     \
            _Pragma("unroll") for (int j = 0; j < NV; ++j) {                                            \
                const uint32_t vi = tig + (uint32_t)(j * G);                                            \
                if (!(active && vi < p.nvec_inner))                                                     \
                    continue;                                                                           \
                const uint32_t col = vi * VEC;                                                          \
                const int valid = (int)min((uint32_t)VEC, p.inner_len - col);                           \
                float *dst = rout + (size_t)row * p.inner_len + col;                                    \
                _Pragma("unroll") for (int i = 0; i < VEC; ++i) if (i < valid) {                        \
                    float y_ = __fdiv_rn(keep[j][i], sd);                                               \
                    if (p.aux0)                                                                         \
                        y_ = __fmul_rn(y_, p.aux0[col + i]);                                            \
                    if (MODE == 2 && p.aux1)                                                            \
                        y_ = __fadd_rn(y_, p.aux1[col + i]);                                            \
                    dst[i] = y_;                                                                        \
                }                                                                                       \
            }                                                                                           \
            return;                                                                                     \
        }                                                                                               \
        float m = ninf;                                                                                 \
        _Pragma("unroll") for (int j = 0; j < NV; ++j)                                                  \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) m = fmaxf(m, keep[j][i]);                   \
        m = group_max<G>(m, red_sm);                                                                    \
        float ssum = 0.0f;                                                                              \
        _Pragma("unroll") for (int j = 0; j < NV; ++j)                                                  \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i) {                                           \
                const float d_ = __fsub_rn(keep[j][i], m);                                              \
                const float e_ = expf(d_);                                                              \
                ssum = __fadd_rn(ssum, e_);                                                             \
                keep[j][i] = LOG ? d_ : e_;                                                             \
            }                                                                                           \
        ssum = group_sum<G>(ssum, red_sm);                                                              \
        const float lg = LOG ? logf(ssum) : 0.0f;                                                       \
        float racc = EPI == 3 ? ninf : 0.0f;                                                            \
        _Pragma("unroll") for (int j = 0; j < NV; ++j) {                                                \
            const uint32_t vi = tig + (uint32_t)(j * G);                                                \
            if (!(active && vi < p.nvec_inner))                                                         \
                continue;                                                                               \
            const uint32_t col = vi * VEC;                                                              \
            const int valid = (int)min((uint32_t)VEC, p.inner_len - col);                               \
            float o_[VEC];                                                                              \
            _Pragma("unroll") for (int i = 0; i < VEC; ++i)                                             \
                o_[i] = LOG ? __fsub_rn(keep[j][i], lg) : __fdiv_rn(keep[j][i], ssum);                  \
            if (EPI == 0) {                                                                             \
                float *dst = rout + (size_t)row * p.inner_len + col;                                    \
                if (VEC == 4 && valid == 4 && (p.inner_len & 3u) == 0 && (reinterpret_cast<size_t>(rout) & 15) == 0) { \
                    *reinterpret_cast<float4 *>(dst) = make_float4(o_[0], o_[1], o_[2], o_[3]);         \
                } else {                                                                                \
                    _Pragma("unroll") for (int i = 0; i < VEC; ++i) if (i < valid) dst[i] = o_[i];      \
                }                                                                                       \
            } else {                                                                                    \
                _Pragma("unroll") for (int i = 0; i < VEC; ++i) if (i < valid)                          \
                    racc = EPI == 3 ? fmaxf(racc, o_[i]) : __fadd_rn(racc, o_[i]);                      \
            }                                                                                           \
        }                                                                                               \
        if (EPI != 0) {                                                                                 \
            racc = EPI == 3 ? group_max<G>(racc, red_sm) : group_sum<G>(racc, red_sm);                  \
            if (EPI == 2)                                                                               \
                racc = __fdiv_rn(racc, (float)p.inner_len);                                             \
            if (tig == 0 && active)                                                                     \
                rout[row] = racc;                                                                       \
        }                                                                                               \
    }

} // namespace ew
} // namespace axb
