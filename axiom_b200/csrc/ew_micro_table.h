// ew_micro_table.h — the micro-op vocabulary of the elementwise code generator (host side).
// Each entry maps a micro-op id to the statement macro (ew_prelude.cuh) and the C expression it expands.
#pragma once
#include <stdint.h>

namespace axb {
namespace ew {

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

struct MicroDef {
    uint8_t id;
    const char *kind; // statement macro suffix: AXB_<kind>(expr)
    const char *expr;
};

#define AXB_MICRO_TABLE(X) \
    X(M_CVT_RND_BF16, CVT, fromF<W>(round_bf16(asF<W>(x)))) \
    X(M_CVT_RND_F16, CVT, fromF<W>(round_f16(asF<W>(x)))) \
    X(M_CVT_I2F, CVT, fromF<W>(__int2float_rn(asI<W>(x)))) \
    X(M_CVT_U2F, CVT, fromF<W>(__uint2float_rn((uint32_t)x))) \
    X(M_CVT_F2I, CVT, fromI<W>(__float2int_rz(asF<W>(x)))) \
    X(M_CVT_F2U, CVT, (W)__float2uint_rz(asF<W>(x))) \
    X(M_CVT_F2BOOL, CVT, (W)(asF<W>(x) != 0.0f ? 1u : 0u)) \
    X(M_CVT_I2BOOL, CVT, (W)((uint32_t)x != 0u ? 1u : 0u)) \
    X(M_CVT_WRAP_I8, CVT, fromI<W>((int32_t)(int8_t)(uint32_t)x)) \
    X(M_CVT_WRAP_U8, CVT, (W)((uint32_t)x & 0xFFu)) \
    X(M_CVT_WRAP_I16, CVT, fromI<W>((int32_t)(int16_t)(uint32_t)x)) \
    X(M_CVT_WRAP_U16, CVT, (W)((uint32_t)x & 0xFFFFu)) \
    X(M_CVT_F2D, CVT64, fromD((double)asF<W>(x))) \
    X(M_CVT_D2F, CVT64, fromF<W>((float)asD(x))) \
    X(M_CVT_I2L, CVT64, (W)(int64_t)asI<W>(x)) \
    X(M_CVT_U2L, CVT64, (W)(uint64_t)(uint32_t)x) \
    X(M_CVT_L2I, CVT64, (W)(uint32_t)x) \
    X(M_CVT_L2F, CVT64, fromF<W>(__ll2float_rn((long long)x))) \
    X(M_CVT_UL2F, CVT64, fromF<W>(__ull2float_rn((unsigned long long)x))) \
    X(M_CVT_L2D, CVT64, fromD(__ll2double_rn((long long)x))) \
    X(M_CVT_UL2D, CVT64, fromD(__ull2double_rn((unsigned long long)x))) \
    X(M_CVT_D2L, CVT64, (W)__double2ll_rz(asD(x))) \
    X(M_CVT_D2UL, CVT64, (W)__double2ull_rz(asD(x))) \
    X(M_CVT_F2L, CVT64, (W)__float2ll_rz(asF<W>(x))) \
    X(M_CVT_F2UL, CVT64, (W)__float2ull_rz(asF<W>(x))) \
    X(M_CVT_D2BOOL, CVT64, (W)(asD(x) != 0.0 ? 1u : 0u)) \
    X(M_CVT_L2BOOL, CVT64, (W)(x != 0 ? 1u : 0u)) \
    X(M_CVT_I2D, CVT64, fromD((double)asI<W>(x))) \
    X(M_CVT_U2D, CVT64, fromD((double)(uint32_t)x)) \
    X(M_CVT_D2I, CVT64, fromI<W>(__double2int_rz(asD(x)))) \
    X(M_CVT_D2U, CVT64, (W)__double2uint_rz(asD(x))) \
    X(M_F_NEG, UN_F, __uint_as_float(__float_as_uint(a) ^ 0x80000000u)) \
    X(M_F_ABS, UN_F, fabsf(a)) \
    X(M_F_SQRT, UN_F, __fsqrt_rn(a)) \
    X(M_F_EXP, UN_F, hwym::Exp(a)) \
    X(M_F_LOG, UN_F, hwym::Log(a)) \
    X(M_F_SIN, UN_F, hwym::Sin(a)) \
    X(M_F_COS, UN_F, hwym::Cos(a)) \
    X(M_F_TAN, UN_F, hwym::Tan(a)) \
    X(M_F_ERF, UN_F, hwym::Erf(a)) \
    X(M_F_SIGN, UN_F, f_sign(a)) \
    X(M_F_FLOOR, UN_F, floorf(a)) \
    X(M_F_CEIL, UN_F, ceilf(a)) \
    X(M_F_TRUNC, UN_F, truncf(a)) \
    X(M_F_ROUND, UN_F, rintf(a)) \
    X(M_F_RECIP, UN_F, __fdiv_rn(1.0f, a)) \
    X(M_F_SQUARE, UN_F, __fmul_rn(a, a)) \
    X(M_F_CBRT, UN_F, hwym::Cbrt(a)) \
    X(M_F_RELU, UN_F, a > 0.0f ? a : 0.0f) \
    X(M_F_SILU, UN_F, hwym::Silu(a)) \
    X(M_F_SIGMOID, UN_F, hwym::Sigmoid(a)) \
    X(M_F_TANH, UN_F, hwym::Tanh(a)) \
    X(M_F_GELU, UN_F, hwym::Gelu(a)) \
    X(M_F_ISNAN, UN_F2I, a != a) \
    X(M_F_ISINF, UN_F2I, isinf(a)) \
    X(M_F_ISFINITE, UN_F2I, isfinite(a)) \
    X(M_F_LNOT, UN_F2I, !(a != 0.0f)) \
    X(M_FA_SQRT, UN_F, (float)sqrt((double)a)) \
    X(M_FA_EXP, UN_F, slow::f_exp(a)) \
    X(M_FA_LOG, UN_F, slow::f_log(a)) \
    X(M_FA_SIN, UN_F, slow::f_sin(a)) \
    X(M_FA_COS, UN_F, slow::f_cos(a)) \
    X(M_FA_TAN, UN_F, slow::f_tan(a)) \
    X(M_FA_ERF, UN_F, slow::f_erf(a)) \
    X(M_FA_CBRT, UN_F, slow::f_cbrt(a)) \
    X(M_FA_SILU, UN_F, slow::f_silu(a)) \
    X(M_FA_SIGMOID, UN_F, slow::f_sigmoid(a)) \
    X(M_FA_TANH, UN_F, slow::f_tanh(a)) \
    X(M_FA_GELU, UN_F, slow::f_gelu(a)) \
    X(M_F_ADD, BIN_F, __fadd_rn(a, b)) \
    X(M_F_SUB, BIN_F, __fsub_rn(a, b)) \
    X(M_F_MUL, BIN_F, __fmul_rn(a, b)) \
    X(M_F_DIV, BIN_F, __fdiv_rn(a, b)) \
    X(M_F_POW, BIN_F, hwym::Pow(a, b)) \
    X(M_F_MOD, BIN_F, slow::f_fmod(a, b)) \
    X(M_F_MAX, BIN_F, (a < b) ? b : a) \
    X(M_F_MIN, BIN_F, (b < a) ? b : a) \
    X(M_F_ATAN2, BIN_F, slow::f_atan2(a, b)) \
    X(M_F_HYPOT, BIN_F, hwym::Hypot(a, b)) \
    X(M_FA_POW, BIN_F, slow::f_pow(a, b)) \
    X(M_FA_HYPOT, BIN_F, slow::f_hypot(a, b)) \
    X(M_F_EQ, CMP_F, a == b) \
    X(M_F_NE, CMP_F, a != b) \
    X(M_F_LT, CMP_F, a < b) \
    X(M_F_LE, CMP_F, a <= b) \
    X(M_F_GT, CMP_F, a > b) \
    X(M_F_GE, CMP_F, a >= b) \
    X(M_F_LAND, CMP_F, (a != 0.0f) && (b != 0.0f)) \
    X(M_F_LOR, CMP_F, (a != 0.0f) || (b != 0.0f)) \
    X(M_F_LXOR, CMP_F, (a != 0.0f) != (b != 0.0f)) \
    X(M_I_NEG, UN_I, 0u - ua) \
    X(M_I_ABS, UN_I, a < 0 ? (int32_t)(0u - ua) : a) \
    X(M_I_SIGN, UN_I, a > 0 ? 1 : (a < 0 ? -1 : 0)) \
    X(M_I_SQUARE, UN_I, ua * ua) \
    X(M_I_RELU, UN_I, a > 0 ? a : 0) \
    X(M_I_RECIP, UN_I, a != 0 ? i_div(1, a) : 0) \
    X(M_U_RECIP, UN_I, ua != 0u ? 1u / ua : 0u) \
    X(M_I_LNOT, UN_I, ua == 0u ? 1 : 0) \
    X(M_I_ADD, BIN_I, ua + ub) \
    X(M_I_SUB, BIN_I, ua - ub) \
    X(M_I_MUL, BIN_I, ua * ub) \
    X(M_I_DIV, BIN_I, i_div(a, b)) \
    X(M_U_DIV, BIN_I, ub == 0u ? 0u : ua / ub) \
    X(M_I_MOD, BIN_I, i_mod(a, b)) \
    X(M_U_MOD, BIN_I, ub == 0u ? 0u : ua % ub) \
    X(M_I_MAX, BIN_I, a < b ? b : a) \
    X(M_U_MAX, BIN_I, ua < ub ? ub : ua) \
    X(M_I_MIN, BIN_I, b < a ? b : a) \
    X(M_U_MIN, BIN_I, ub < ua ? ub : ua) \
    X(M_I_AND, BIN_I, ua & ub) \
    X(M_I_OR, BIN_I, ua | ub) \
    X(M_I_XOR, BIN_I, ua ^ ub) \
    X(M_I_SHL, BIN_I, ua << (ub & 31u)) \
    X(M_I_SHR, BIN_I, a >> (ub & 31u)) \
    X(M_U_SHR, BIN_I, ua >> (ub & 31u)) \
    X(M_D_NEG, UN_D, -a) \
    X(M_D_ABS, UN_D, fabs(a)) \
    X(M_D_SQRT, UN_D, sqrt(a)) \
    X(M_D_EXP, UN_D, slow::d_exp(a)) \
    X(M_D_LOG, UN_D, slow::d_log(a)) \
    X(M_D_SIN, UN_D, slow::d_sin(a)) \
    X(M_D_COS, UN_D, slow::d_cos(a)) \
    X(M_D_TAN, UN_D, slow::d_tan(a)) \
    X(M_D_TANH, UN_D, slow::d_tanh(a)) \
    X(M_D_ERF, UN_D, slow::d_erf(a)) \
    X(M_D_SIGN, UN_D, a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : 0.0)) \
    X(M_D_FLOOR, UN_D, floor(a)) \
    X(M_D_CEIL, UN_D, ceil(a)) \
    X(M_D_TRUNC, UN_D, trunc(a)) \
    X(M_D_ROUND, UN_D, rint(a)) \
    X(M_D_RECIP, UN_D, 1.0 / a) \
    X(M_D_SQUARE, UN_D, a * a) \
    X(M_D_RELU, UN_D, a > 0.0 ? a : 0.0) \
    X(M_D_SIGMOID, UN_D, slow::d_sigmoid(a)) \
    X(M_D_ISNAN, UN_D2I, a != a) \
    X(M_D_ISINF, UN_D2I, isinf(a)) \
    X(M_D_ISFINITE, UN_D2I, isfinite(a)) \
    X(M_D_LNOT, UN_D2I, !(a != 0.0)) \
    X(M_D_ADD, BIN_D, __dadd_rn(a, b)) \
    X(M_D_SUB, BIN_D, __dsub_rn(a, b)) \
    X(M_D_MUL, BIN_D, __dmul_rn(a, b)) \
    X(M_D_DIV, BIN_D, __ddiv_rn(a, b)) \
    X(M_D_POW, BIN_D, slow::d_pow(a, b)) \
    X(M_D_MOD, BIN_D, slow::d_fmod(a, b)) \
    X(M_D_MAX, BIN_D, (a < b) ? b : a) \
    X(M_D_MIN, BIN_D, (b < a) ? b : a) \
    X(M_D_HYPOT, BIN_D, slow::d_hypot(a, b)) \
    X(M_D_ATAN2, BIN_D, slow::d_atan2(a, b)) \
    X(M_L_NEG, UN_L, 0ull - (uint64_t)a) \
    X(M_L_ABS, UN_L, a < 0 ? (int64_t)(0ull - (uint64_t)a) : a) \
    X(M_L_SIGN, UN_L, a > 0 ? 1 : (a < 0 ? -1 : 0)) \
    X(M_L_SQUARE, UN_L, (uint64_t)a * (uint64_t)a) \
    X(M_L_RELU, UN_L, a > 0 ? a : 0) \
    X(M_L_LNOT, UN_L, a == 0 ? 1 : 0) \
    X(M_L_ADD, BIN_L, ua + ub) \
    X(M_L_SUB, BIN_L, ua - ub) \
    X(M_L_MUL, BIN_L, ua * ub) \
    X(M_L_DIV, BIN_L, l_div(a, b)) \
    X(M_UL_DIV, BIN_L, ub == 0 ? 0ull : ua / ub) \
    X(M_L_MOD, BIN_L, l_mod(a, b)) \
    X(M_UL_MOD, BIN_L, ub == 0 ? 0ull : ua % ub) \
    X(M_L_MAX, BIN_L, a < b ? b : a) \
    X(M_UL_MAX, BIN_L, ua < ub ? ub : ua) \
    X(M_L_MIN, BIN_L, b < a ? b : a) \
    X(M_UL_MIN, BIN_L, ub < ua ? ub : ua) \
    X(M_L_AND, BIN_L, ua & ub) \
    X(M_L_OR, BIN_L, ua | ub) \
    X(M_L_XOR, BIN_L, ua ^ ub) \
    X(M_L_SHL, BIN_L, ua << (ub & 63u)) \
    X(M_L_SHR, BIN_L, a >> (ub & 63u)) \
    X(M_UL_SHR, BIN_L, ua >> (ub & 63u))

#define AXB_MICRO_ROW(ID, KIND, ...) {ID, #KIND, #__VA_ARGS__},
static const MicroDef kMicroDefs[] = {AXB_MICRO_TABLE(AXB_MICRO_ROW)};
#undef AXB_MICRO_ROW

} // namespace ew
} // namespace axb
