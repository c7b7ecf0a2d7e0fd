// elementwise.cu — host side of kernel family 1: lowers an axb_instr program to the straight-line statement
// list of ew_prelude.cuh (one specialised sm_100a kernel per program signature, see jit.cpp), collapses the
// iteration space and launches ONE kernel per fusion group.
//
// Semantics follow the reference op by op (file:line in each rule below); see include/axb.h.
#include <algorithm>
#include <unordered_map>
#include <mutex>
#include <cstring>
#include <string>
#include <vector>

#include "axb_internal.h"
#include "ew_micro_table.h"
#include "ew_prelude.cuh"
#include <cuda_fp16.h>

#include "jit.h"

namespace axb {
namespace ew {

namespace {

enum Cls { C_F, C_I, C_D, C_L, C_BAD };
Cls cls_of(int dt) {
    switch (dt) {
    case AXB_F32: case AXB_BF16: case AXB_F16: return C_F;
    case AXB_BOOL: case AXB_I8: case AXB_I16: case AXB_I32: case AXB_U8: case AXB_U16: case AXB_U32: return C_I;
    case AXB_F64: return C_D;
    case AXB_I64: case AXB_U64: return C_L;
    default: return C_BAD;
    }
}
bool is_unsigned(int dt) { return dt == AXB_U8 || dt == AXB_U16 || dt == AXB_U32 || dt == AXB_U64 || dt == AXB_BOOL; }

const MicroDef *micro_def(uint32_t op) {
    for (const MicroDef &d : kMicroDefs)
        if (d.id == op)
            return &d;
    return nullptr;
}
const char *reg_name(uint32_t r) {
    static const char *names[] = {"t", "acc", "u", "s0", "s1", "s2", "s3"};
    return r < 7 ? names[r] : "acc";
}
const char *mode_name(int m) { return m == MODE_VEC ? "MODE_VEC" : m == MODE_BCAST ? "MODE_BCAST" : "MODE_STRIDED"; }

// Emits the statement list of one specialised kernel. The text depends only on the program SIGNATURE (ops,
// dtypes, operand access modes): shapes, strides, pointers and immediates stay run-time parameters.
struct Builder {
    Params &p;
    bool wide;
    int n_micro = 0, n_imm = 0;
    std::string code, loads;
    bool preloaded[kMaxIn] = {false, false, false, false, false, false, false, false};
    axb_status st = AXB_OK;
    explicit Builder(Params &pp, bool w) : p(pp), wide(w) {}
    void line(const std::string &s) {
        code += "            ";
        code += s;
        code += "\n";
        ++n_micro;
    }
    void emit(uint32_t op, uint32_t a = 0, uint32_t b = 0, uint32_t fl = 0) {
        if (st != AXB_OK)
            return;
        if (n_micro >= kMaxMicro) {
            st = fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: program too long after lowering (> %d statements)", kMaxMicro);
            return;
        }
        char buf[512];
        switch (op) {
        case M_LOAD:
            if (!preloaded[a]) { // every input is fetched once per vector, up front (ew_run issues these for U vectors)
                preloaded[a] = true;
                snprintf(buf, sizeof buf, "            EW_PRELOAD(%u, %d, %s)\n", a, (int)p.in[a].dtype, mode_name(p.in[a].mode));
                loads += buf;
            }
            snprintf(buf, sizeof buf, "EW_USE(%u, %s)", a, reg_name(b));
            line(buf);
            return;
        case M_IMM:
            snprintf(buf, sizeof buf, "EW_IMM(%u, %s)", a, reg_name(b));
            line(buf);
            return;
        case M_MOV:
            snprintf(buf, sizeof buf, "EW_MOV(%s, %s)", reg_name(a), reg_name(b));
            line(buf);
            return;
        case M_STORE:
            snprintf(buf, sizeof buf, "EW_STORE(%d, %s)", (int)p.out.dtype, mode_name(p.out.mode));
            line(buf);
            return;
        case M_SEL: line("EW_SEL"); return;
        case M_F_LEAKY:
            snprintf(buf, sizeof buf, "EW_LEAKY(%u)", a);
            line(buf);
            break;
        default: {
            const MicroDef *d = micro_def(op);
            if (!d) {
                st = fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: internal: micro-op %u has no definition", op);
                return;
            }
            std::string s;
            if (!strcmp(d->kind, "CVT") || !strcmp(d->kind, "CVT64")) {
                s = std::string("EW_CVT_ON(") + (a == R_T ? "R_T" : a == R_ACC ? "R_ACC" : "R_U") + ", AXB_" + d->kind + "(" +
                    d->expr + "))";
                line(s);
                return; // conversions carry no rounding flag
            }
            std::string expr = d->expr;
            if ((fl & 3u) == RND_BF16 || (fl & 3u) == RND_F16) {
                // half-type transcendentals: fp32 evaluation + exact fallback near rounding boundaries (ew_prelude.cuh hfast::)
                const char *fast = nullptr;
                switch (op) {
                case M_FA_EXP: fast = "Exp"; break;
                case M_FA_LOG: fast = "Log"; break;
                case M_FA_SIN: fast = "Sin"; break;
                case M_FA_COS: fast = "Cos"; break;
                case M_FA_TAN: fast = "Tan"; break;
                case M_FA_ERF: fast = "Erf"; break;
                case M_FA_CBRT: fast = "Cbrt"; break;
                case M_FA_TANH: fast = "Tanh"; break;
                case M_FA_SIGMOID: fast = "Sigmoid"; break;
                case M_FA_SILU: fast = "Silu"; break;
                case M_FA_GELU: fast = "Gelu"; break;
                case M_FA_SQRT: expr = "__fsqrt_rn(a)"; break; // float(sqrt(double)) == correctly rounded fp32 sqrt
                default: break;
                }
                if (fast)
                    expr = std::string("hfast::") + fast + ((fl & 3u) == RND_BF16 ? "<16>(a)" : "<13>(a)");
            }
            s = std::string("EW_OP(") + ((fl & FL_SWAP) ? "true" : "false") + ", AXB_" + d->kind + "(" + expr + "))";
            line(s);
            break;
        }
        }
        if ((fl & 3u) == RND_BF16)
            line("EW_RND_BF16");
        else if ((fl & 3u) == RND_F16)
            line("EW_RND_F16");
    }
    int add_imm(uint64_t bits) {
        // no de-duplication: the slot numbering (and so the kernel text) must not depend on immediate VALUES
        if (n_imm >= kMaxImm) {
            st = fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: too many immediates (> %d)", kMaxImm);
            return 0;
        }
        p.imm[n_imm] = bits;
        return n_imm++;
    }
    // canonical-register conversion from dtype `from` to dtype `to` on register `r`
    void cvt(int from, int to, uint32_t r) {
        if (from == to || st != AXB_OK)
            return;
        Cls cf = cls_of(from), ct = cls_of(to);
        if (cf == C_BAD || ct == C_BAD) {
            st = fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: conversion %d -> %d not supported", from, to);
            return;
        }
        auto narrow_int = [&](int dt) {
            // value is a 32-bit int in register; wrap to the storage width (C++ integral conversion)
            switch (dt) {
            case AXB_I8: emit(M_CVT_WRAP_I8, r); break;
            case AXB_U8: emit(M_CVT_WRAP_U8, r); break;
            case AXB_I16: emit(M_CVT_WRAP_I16, r); break;
            case AXB_U16: emit(M_CVT_WRAP_U16, r); break;
            default: break;
            }
        };
        auto round_half = [&](int dt) {
            if (dt == AXB_BF16)
                emit(M_CVT_RND_BF16, r);
            else if (dt == AXB_F16)
                emit(M_CVT_RND_F16, r);
        };
        if (to == AXB_BOOL) { // static_cast<bool>(x): x != 0
            switch (cf) {
            case C_F: emit(M_CVT_F2BOOL, r); break;
            case C_I: emit(M_CVT_I2BOOL, r); break;
            case C_D: emit(M_CVT_D2BOOL, r); break;
            case C_L: emit(M_CVT_L2BOOL, r); break;
            default: break;
            }
            return;
        }
        switch (cf) {
        case C_F:
            switch (ct) {
            case C_F: round_half(to); break; // f32/bf16/f16 -> narrower float (widening is a no-op)
            case C_I:
                if (to == AXB_U32)
                    emit(M_CVT_F2U, r);
                else {
                    emit(M_CVT_F2I, r);
                    narrow_int(to);
                }
                break;
            case C_D: emit(M_CVT_F2D, r); break;
            case C_L: emit(to == AXB_U64 ? M_CVT_F2UL : M_CVT_F2L, r); break;
            default: break;
            }
            break;
        case C_I:
            switch (ct) {
            case C_F:
                emit(from == AXB_U32 ? M_CVT_U2F : M_CVT_I2F, r);
                round_half(to);
                break;
            case C_I: narrow_int(to); break;
            case C_D: emit(from == AXB_U32 ? M_CVT_U2D : M_CVT_I2D, r); break;
            case C_L: emit(from == AXB_U32 ? M_CVT_U2L : M_CVT_I2L, r); break;
            default: break;
            }
            break;
        case C_D:
            switch (ct) {
            case C_F:
                emit(M_CVT_D2F, r);
                round_half(to);
                break;
            case C_I:
                emit(to == AXB_U32 ? M_CVT_D2U : M_CVT_D2I, r);
                narrow_int(to);
                break;
            case C_D: break;
            case C_L: emit(to == AXB_U64 ? M_CVT_D2UL : M_CVT_D2L, r); break;
            default: break;
            }
            break;
        case C_L:
            switch (ct) {
            case C_F:
                emit(from == AXB_U64 ? M_CVT_UL2F : M_CVT_L2F, r);
                round_half(to);
                break;
            case C_I:
                emit(M_CVT_L2I, r);
                narrow_int(to);
                break;
            case C_D: emit(from == AXB_U64 ? M_CVT_UL2D : M_CVT_L2D, r); break;
            case C_L: break;
            default: break;
            }
            break;
        default: break;
        }
    }
};

// immediate -> canonical register bits for dtype dt
uint64_t imm_bits(double v, int dt) {
    switch (cls_of(dt)) {
    case C_F: {
        float f = (float)v;
        if (dt == AXB_BF16) {
            uint32_t b;
            memcpy(&b, &f, 4);
            if (f != f)
                b = 0x7FC00000u;
            else {
                b += 0x7FFFu + ((b >> 16) & 1u);
                b &= 0xFFFF0000u;
            }
            return b;
        }
        uint32_t b;
        memcpy(&b, &f, 4);
        return b; // (fp16 immediates are rounded on device by the consumer's rounding flag)
    }
    case C_I:
        if (dt == AXB_BOOL)
            return v != 0.0 ? 1u : 0u;
        return (uint64_t)(uint32_t)(int32_t)(int64_t)v;
    case C_D: {
        uint64_t b;
        memcpy(&b, &v, 8);
        return b;
    }
    case C_L: return (uint64_t)(int64_t)v;
    default: return 0;
    }
}

struct OpInfo {
    int arity;  // 1, 2, 3
    enum Kind { ARITH, CMP, LOGIC2, TEST, LNOT, CASTK, WHEREK } kind;
};
bool op_info(int op, OpInfo *o) {
    switch (op) {
    case AXB_OP_ADD: case AXB_OP_SUB: case AXB_OP_MUL: case AXB_OP_DIV: case AXB_OP_POW: case AXB_OP_MOD:
    case AXB_OP_MAXIMUM: case AXB_OP_MINIMUM: case AXB_OP_ATAN2: case AXB_OP_HYPOT: case AXB_OP_BIT_AND:
    case AXB_OP_BIT_OR: case AXB_OP_BIT_XOR: case AXB_OP_SHL: case AXB_OP_SHR:
        *o = {2, OpInfo::ARITH};
        return true;
    case AXB_OP_EQ: case AXB_OP_NE: case AXB_OP_LT: case AXB_OP_LE: case AXB_OP_GT: case AXB_OP_GE:
        *o = {2, OpInfo::CMP};
        return true;
    case AXB_OP_LOGICAL_AND: case AXB_OP_LOGICAL_OR: case AXB_OP_LOGICAL_XOR:
        *o = {2, OpInfo::LOGIC2};
        return true;
    case AXB_OP_LOGICAL_NOT:
        *o = {1, OpInfo::LNOT};
        return true;
    case AXB_OP_ISNAN: case AXB_OP_ISINF: case AXB_OP_ISFINITE:
        *o = {1, OpInfo::TEST};
        return true;
    case AXB_OP_NEG: case AXB_OP_ABS: case AXB_OP_SQRT: case AXB_OP_EXP: case AXB_OP_LOG: case AXB_OP_SIN:
    case AXB_OP_COS: case AXB_OP_TAN: case AXB_OP_ERF: case AXB_OP_SIGN: case AXB_OP_FLOOR: case AXB_OP_CEIL:
    case AXB_OP_TRUNC: case AXB_OP_ROUND: case AXB_OP_RECIPROCAL: case AXB_OP_SQUARE: case AXB_OP_CBRT:
    case AXB_OP_RELU: case AXB_OP_LEAKY_RELU: case AXB_OP_SILU: case AXB_OP_SIGMOID: case AXB_OP_TANH:
    case AXB_OP_GELU:
        *o = {1, OpInfo::ARITH};
        return true;
    case AXB_OP_CAST:
        *o = {1, OpInfo::CASTK};
        return true;
    case AXB_OP_WHERE:
        *o = {3, OpInfo::WHEREK};
        return true;
    default: return false;
    }
}

// micro opcode for a unary arithmetic op evaluated in dtype dt; 0 = unsupported
uint32_t unary_micro(int op, int dt) {
    Cls c = cls_of(dt);
    bool half = (dt == AXB_BF16 || dt == AXB_F16);
    if (c == C_F) {
        switch (op) {
        case AXB_OP_NEG: return M_F_NEG;
        case AXB_OP_ABS: return M_F_ABS;
        case AXB_OP_SQRT: return half ? M_FA_SQRT : M_F_SQRT;
        case AXB_OP_EXP: return half ? M_FA_EXP : M_F_EXP;
        case AXB_OP_LOG: return half ? M_FA_LOG : M_F_LOG;
        case AXB_OP_SIN: return half ? M_FA_SIN : M_F_SIN;
        case AXB_OP_COS: return half ? M_FA_COS : M_F_COS;
        case AXB_OP_TAN: return half ? M_FA_TAN : M_F_TAN;
        case AXB_OP_ERF: return half ? M_FA_ERF : M_F_ERF;
        case AXB_OP_SIGN: return M_F_SIGN;
        // bf16 floor/ceil/trunc/round are identity in the reference (cpu_operations.hpp:728-770: only
        // is_floating_point_v and float16_t branches act) -> handled by caller
        case AXB_OP_FLOOR: return M_F_FLOOR;
        case AXB_OP_CEIL: return M_F_CEIL;
        case AXB_OP_TRUNC: return M_F_TRUNC;
        case AXB_OP_ROUND: return M_F_ROUND;
        case AXB_OP_RECIPROCAL: return M_F_RECIP;
        case AXB_OP_SQUARE: return M_F_SQUARE;
        case AXB_OP_CBRT: return half ? M_FA_CBRT : M_F_CBRT;
        case AXB_OP_RELU: return M_F_RELU;
        case AXB_OP_LEAKY_RELU: return M_F_LEAKY;
        case AXB_OP_SILU: return half ? M_FA_SILU : M_F_SILU;
        case AXB_OP_SIGMOID: return half ? M_FA_SIGMOID : M_F_SIGMOID;
        case AXB_OP_TANH: return half ? M_FA_TANH : M_F_TANH;
        case AXB_OP_GELU: return half ? M_FA_GELU : M_F_GELU;
        default: return 0;
        }
    }
    if (c == C_I) {
        bool u = is_unsigned(dt);
        switch (op) {
        case AXB_OP_NEG: return M_I_NEG; // bool handled by caller (!a)
        case AXB_OP_ABS: return u ? (uint32_t)M_MOV : (uint32_t)M_I_ABS;
        case AXB_OP_SIGN: return u ? (uint32_t)M_I_LNOT /*placeholder, fixed by caller*/ : (uint32_t)M_I_SIGN;
        case AXB_OP_SQUARE: return M_I_SQUARE;
        case AXB_OP_RELU: return u ? (uint32_t)M_MOV : (uint32_t)M_I_RELU;
        case AXB_OP_RECIPROCAL: return u ? M_U_RECIP : M_I_RECIP;
        case AXB_OP_FLOOR: case AXB_OP_CEIL: case AXB_OP_TRUNC: case AXB_OP_ROUND: return M_MOV; // identity
        default: return 0;
        }
    }
    if (c == C_D) {
        switch (op) {
        case AXB_OP_NEG: return M_D_NEG;
        case AXB_OP_ABS: return M_D_ABS;
        case AXB_OP_SQRT: return M_D_SQRT;
        case AXB_OP_EXP: return M_D_EXP;
        case AXB_OP_LOG: return M_D_LOG;
        case AXB_OP_SIN: return M_D_SIN;
        case AXB_OP_COS: return M_D_COS;
        case AXB_OP_TAN: return M_D_TAN;
        case AXB_OP_TANH: return M_D_TANH;
        case AXB_OP_ERF: return M_D_ERF;
        case AXB_OP_SIGN: return M_D_SIGN;
        case AXB_OP_FLOOR: return M_D_FLOOR;
        case AXB_OP_CEIL: return M_D_CEIL;
        case AXB_OP_TRUNC: return M_D_TRUNC;
        case AXB_OP_ROUND: return M_D_ROUND;
        case AXB_OP_RECIPROCAL: return M_D_RECIP;
        case AXB_OP_SQUARE: return M_D_SQUARE;
        case AXB_OP_RELU: return M_D_RELU;
        case AXB_OP_SIGMOID: return M_D_SIGMOID;
        default: return 0;
        }
    }
    if (c == C_L) {
        bool u = is_unsigned(dt);
        switch (op) {
        case AXB_OP_NEG: return M_L_NEG;
        case AXB_OP_ABS: return u ? (uint32_t)M_MOV : (uint32_t)M_L_ABS;
        case AXB_OP_SIGN: return M_L_SIGN;
        case AXB_OP_SQUARE: return M_L_SQUARE;
        case AXB_OP_RELU: return u ? (uint32_t)M_MOV : (uint32_t)M_L_RELU;
        case AXB_OP_FLOOR: case AXB_OP_CEIL: case AXB_OP_TRUNC: case AXB_OP_ROUND: return M_MOV;
        default: return 0;
        }
    }
    return 0;
}

// simd = both operands same-shape & contiguous: the reference then runs its Highway kernels (pow = exp(b*log a),
// hypot = sqrt(a*a+b*b): cpu_operations.cpp:229-262); otherwise the scalar libm functors (:516-583).
uint32_t binary_micro(int op, int dt, bool simd) {
    Cls c = cls_of(dt);
    bool half = (dt == AXB_BF16 || dt == AXB_F16) || !simd;
    bool u = is_unsigned(dt);
    if (c == C_F) {
        switch (op) {
        case AXB_OP_ADD: return M_F_ADD;
        case AXB_OP_SUB: return M_F_SUB;
        case AXB_OP_MUL: return M_F_MUL;
        case AXB_OP_DIV: return M_F_DIV;
        case AXB_OP_POW: return half ? M_FA_POW : M_F_POW;
        case AXB_OP_MOD: return M_F_MOD;
        case AXB_OP_MAXIMUM: return M_F_MAX;
        case AXB_OP_MINIMUM: return M_F_MIN;
        case AXB_OP_ATAN2: return M_F_ATAN2;
        case AXB_OP_HYPOT: return half ? M_FA_HYPOT : M_F_HYPOT;
        default: return 0; // bitwise on floats: TypeError in the reference (cpu_operations.cpp:52-67)
        }
    }
    if (c == C_I) {
        switch (op) {
        case AXB_OP_ADD: return M_I_ADD;
        case AXB_OP_SUB: return M_I_SUB;
        case AXB_OP_MUL: return M_I_MUL;
        case AXB_OP_DIV: return u ? M_U_DIV : M_I_DIV;
        case AXB_OP_MOD: return u ? M_U_MOD : M_I_MOD;
        case AXB_OP_MAXIMUM: return u ? M_U_MAX : M_I_MAX;
        case AXB_OP_MINIMUM: return u ? M_U_MIN : M_I_MIN;
        case AXB_OP_BIT_AND: return M_I_AND;
        case AXB_OP_BIT_OR: return M_I_OR;
        case AXB_OP_BIT_XOR: return M_I_XOR;
        case AXB_OP_SHL: return M_I_SHL;
        case AXB_OP_SHR: return u ? M_U_SHR : M_I_SHR;
        default: return 0;
        }
    }
    if (c == C_D) {
        switch (op) {
        case AXB_OP_ADD: return M_D_ADD;
        case AXB_OP_SUB: return M_D_SUB;
        case AXB_OP_MUL: return M_D_MUL;
        case AXB_OP_DIV: return M_D_DIV;
        case AXB_OP_POW: return M_D_POW;
        case AXB_OP_MOD: return M_D_MOD;
        case AXB_OP_MAXIMUM: return M_D_MAX;
        case AXB_OP_MINIMUM: return M_D_MIN;
        case AXB_OP_ATAN2: return M_D_ATAN2;
        case AXB_OP_HYPOT: return M_D_HYPOT;
        default: return 0;
        }
    }
    if (c == C_L) {
        switch (op) {
        case AXB_OP_ADD: return M_L_ADD;
        case AXB_OP_SUB: return M_L_SUB;
        case AXB_OP_MUL: return M_L_MUL;
        case AXB_OP_DIV: return u ? M_UL_DIV : M_L_DIV;
        case AXB_OP_MOD: return u ? M_UL_MOD : M_L_MOD;
        case AXB_OP_MAXIMUM: return u ? M_UL_MAX : M_L_MAX;
        case AXB_OP_MINIMUM: return u ? M_UL_MIN : M_L_MIN;
        case AXB_OP_BIT_AND: return M_L_AND;
        case AXB_OP_BIT_OR: return M_L_OR;
        case AXB_OP_BIT_XOR: return M_L_XOR;
        case AXB_OP_SHL: return M_L_SHL;
        case AXB_OP_SHR: return u ? M_UL_SHR : M_L_SHR;
        default: return 0;
        }
    }
    return 0;
}

uint32_t cmp_micro(int op) {
    switch (op) {
    case AXB_OP_EQ: return M_F_EQ;
    case AXB_OP_NE: return M_F_NE;
    case AXB_OP_LT: return M_F_LT;
    case AXB_OP_LE: return M_F_LE;
    case AXB_OP_GT: return M_F_GT;
    case AXB_OP_GE: return M_F_GE;
    case AXB_OP_LOGICAL_AND: return M_F_LAND;
    case AXB_OP_LOGICAL_OR: return M_F_LOR;
    case AXB_OP_LOGICAL_XOR: return M_F_LXOR;
    default: return 0;
    }
}

} // namespace

// Lower the typed program to micro-ops. `src_dtype(k)` = dtype of input k.
axb_status lower_program(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, bool wide,
                         Params &p, int *final_dtype, bool emit_store, std::string *code_out, std::string *loads_out) {
    Builder B(p, wide);
    std::vector<int> last_use(n_instr, -1), n_uses(n_instr, 0), slot_of(n_instr, -1);
    for (int i = 0; i < n_instr; ++i) {
        OpInfo oi;
        if (!op_info(prog[i].op, &oi))
            return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: op %d is not an elementwise op", prog[i].op);
        for (int j = 0; j < oi.arity; ++j) {
            int s = prog[i].src[j];
            if (s == AXB_SRC_NONE)
                return fail(AXB_ERR_INVALID, "axb_elementwise: instr %d operand %d missing", i, j);
            if (s == AXB_SRC_IMM)
                continue;
            if (s >= 0) {
                if (s >= n_inputs)
                    return fail(AXB_ERR_INVALID, "axb_elementwise: instr %d references input %d of %d", i, s, n_inputs);
                continue;
            }
            int v = -1 - s;
            if (v >= i)
                return fail(AXB_ERR_INVALID, "axb_elementwise: instr %d references value %d (not yet defined)", i, v);
            last_use[v] = i;
            n_uses[v]++;
        }
    }
    for (int v = 0; v < n_instr - 1; ++v)
        if (n_uses[v] == 0)
            return fail(AXB_ERR_INVALID, "axb_elementwise: value %d is never used (dead code)", v);

    // shape of every value (broadcast of its operands) — needed for the reference's path-dependent numerics
    std::vector<std::vector<int64_t>> vshape(n_instr);
    auto src_shape = [&](int sx) -> std::vector<int64_t> {
        if (sx == AXB_SRC_IMM || sx == AXB_SRC_NONE)
            return {1};
        if (sx >= 0)
            return std::vector<int64_t>(inputs[sx].shape, inputs[sx].shape + inputs[sx].ndim);
        return vshape[-1 - sx];
    };
    auto src_contig = [&](int sx) -> bool { return (sx >= 0) ? is_contiguous(inputs[sx]) : true; };
    auto same_shape_contig = [&](int a, int b) -> bool {
        return src_shape(a) == src_shape(b) && src_contig(a) && src_contig(b);
    };
    auto bshape2 = [&](std::vector<int64_t> a, const std::vector<int64_t> &b) {
        size_t n = std::max(a.size(), b.size());
        std::vector<int64_t> r(n, 1);
        for (size_t i = 0; i < n; ++i) {
            int64_t ea = i + a.size() >= n ? a[i + a.size() - n] : 1;
            int64_t eb = i + b.size() >= n ? b[i + b.size() - n] : 1;
            r[i] = ea == 1 ? eb : ea;
        }
        return r;
    };
    bool slot_busy[kNSlot] = {false, false, false, false};
    int acc_holds = -1;
    auto value_dtype = [&](int v) { return prog[v].dtype; };

    // fetch operand selector `s` into register r (R_T/R_ACC/R_U); returns its dtype (or `want` for IMM)
    auto fetch = [&](int s, double imm, uint32_t r, int want, int *dt_out) -> axb_status {
        if (s == AXB_SRC_IMM) {
            int idx = B.add_imm(imm_bits(imm, want));
            B.emit(M_IMM, idx, r);
            *dt_out = (want == AXB_F16) ? AXB_F32 : want; // fp16 imm: rounded by cvt below
            return B.st;
        }
        if (s >= 0) {
            B.emit(M_LOAD, s, r);
            *dt_out = inputs[s].dtype;
            return B.st;
        }
        int v = -1 - s;
        *dt_out = value_dtype(v);
        if (r == R_ACC && acc_holds == v)
            return AXB_OK;
        if (slot_of[v] >= 0) {
            B.emit(M_MOV, R_S0 + slot_of[v], r);
            return B.st;
        }
        if (acc_holds == v) {
            B.emit(M_MOV, R_ACC, r);
            return B.st;
        }
        return fail(AXB_ERR_INVALID, "axb_elementwise: internal: value %d not resident", v);
    };

    for (int i = 0; i < n_instr && B.st == AXB_OK; ++i) {
        const axb_instr &in = prog[i];
        OpInfo oi;
        op_info(in.op, &oi);
        const int rdt = in.dtype;
        if (cls_of(rdt) == C_BAD)
            return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: result dtype %d unsupported (complex has no GPU path)", rdt);
        if (!wide && is_64bit_dtype(rdt))
            return fail(AXB_ERR_INVALID, "axb_elementwise: internal: 64-bit dtype in narrow program");
        int d0 = 0, d1 = 0, d2 = 0;
        uint32_t flags = 0;
        if (oi.kind == OpInfo::WHEREK) {
            // cond ? a : b, a/b converted to the result dtype (cpu_operations.cpp:2401-2589); cond: any
            // non-zero byte is true (:2503-2513) == canonical word != 0
            AXB_TRY(fetch(in.src[1], in.imm, R_T, rdt, &d1));
            B.cvt(d1, rdt, R_T);
            AXB_TRY(fetch(in.src[2], in.imm, R_U, rdt, &d2));
            B.cvt(d2, rdt, R_U);
            AXB_TRY(fetch(in.src[0], in.imm, R_ACC, AXB_BOOL, &d0));
            B.emit(M_SEL);
        } else if (oi.arity == 1) {
            int ct = rdt;
            int sdt;
            AXB_TRY(fetch(in.src[0], in.imm, R_ACC, rdt, &sdt));
            if (oi.kind == OpInfo::CASTK) {
                B.cvt(sdt, rdt, R_ACC); // Tensor::astype (tensor.cpp:1656-1701)
            } else if (oi.kind == OpInfo::LNOT || oi.kind == OpInfo::TEST) {
                // evaluated on the INPUT dtype, Bool out (cpu_operations.cpp:836-880)
                if (rdt != AXB_BOOL)
                    return fail(AXB_ERR_INVALID, "axb_elementwise: instr %d must produce Bool", i);
                Cls c = cls_of(sdt);
                if (oi.kind == OpInfo::LNOT)
                    B.emit(c == C_F ? M_F_LNOT : c == C_I ? M_I_LNOT : c == C_D ? M_D_LNOT : M_L_LNOT);
                else if (c == C_F)
                    B.emit(in.op == AXB_OP_ISNAN ? M_F_ISNAN : in.op == AXB_OP_ISINF ? M_F_ISINF : M_F_ISFINITE);
                else if (c == C_D)
                    B.emit(in.op == AXB_OP_ISNAN ? M_D_ISNAN : in.op == AXB_OP_ISINF ? M_D_ISINF : M_D_ISFINITE);
                else { // integers: isnan/isinf = false, isfinite = true
                    int idx = B.add_imm(in.op == AXB_OP_ISFINITE ? 1 : 0);
                    B.emit(M_IMM, idx, R_ACC);
                }
            } else {
                B.cvt(sdt, ct, R_ACC);
                uint32_t m = 0;
                if (ct == AXB_BOOL) {
                    // NegateFunc on bool = !a; Sign/Abs/ReLU/Square/Reciprocal on bool = a
                    if (in.op == AXB_OP_NEG)
                        m = M_I_LNOT;
                    else if (in.op == AXB_OP_SIGN || in.op == AXB_OP_ABS || in.op == AXB_OP_RELU ||
                             in.op == AXB_OP_SQUARE || in.op == AXB_OP_RECIPROCAL)
                        m = M_MOV;
                } else if (ct == AXB_BF16 && (in.op == AXB_OP_FLOOR || in.op == AXB_OP_CEIL ||
                                              in.op == AXB_OP_TRUNC || in.op == AXB_OP_ROUND)) {
                    m = M_MOV; // reference: identity for bfloat16 (cpu_operations.hpp:728-770)
                } else if (cls_of(ct) == C_I && is_unsigned(ct) && in.op == AXB_OP_SIGN) {
                    m = M_CVT_I2BOOL; // a > 0 ? 1 : 0
                } else if (cls_of(ct) == C_L && is_unsigned(ct) && in.op == AXB_OP_SIGN) {
                    m = M_CVT_L2BOOL;
                } else {
                    m = unary_micro(in.op, ct);
                }
                if (!m)
                    return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: unary op %d on dtype %d has no GPU kernel", in.op, ct);
                if (m == M_CVT_I2BOOL || m == M_CVT_L2BOOL)
                    B.emit(m, R_ACC);
                else if (m != M_MOV) {
                    uint32_t a = 0;
                    if (m == M_F_LEAKY) {
                        float sl = (float)in.imm; // static_cast<T>(negative_slope) (cpu_operations.hpp:656-661)
                        if (ct == AXB_BF16) {
                            uint64_t b = imm_bits(in.imm, AXB_BF16);
                            uint32_t bb = (uint32_t)b;
                            memcpy(&sl, &bb, 4);
                        } else if (ct == AXB_F16) {
                            sl = __half2float(__float2half_rn(sl)); // float16_t(float) rounds the slope first (float16.hpp:15)
                        }
                        uint32_t bits;
                        memcpy(&bits, &sl, 4);
                        a = B.add_imm(bits);
                    }
                    if (is_half_dtype(ct))
                        flags |= (ct == AXB_BF16 ? RND_BF16 : RND_F16);
                    B.emit(m, a, 0, flags);
                    if (cls_of(ct) == C_I && ct != AXB_I32 && ct != AXB_U32 && ct != AXB_BOOL)
                        B.cvt(AXB_I32, ct, R_ACC);
                }
            }
        } else {
            // binary
            int ct;
            if (oi.kind == OpInfo::CMP || oi.kind == OpInfo::LOGIC2) {
                // both operands converted to Float32 first, Bool out (cpu_operations.cpp:115-126,286-296)
                if (rdt != AXB_BOOL)
                    return fail(AXB_ERR_INVALID, "axb_elementwise: comparison/logical instr %d must produce Bool", i);
                ct = AXB_F32;
            } else if (rdt == AXB_BOOL) {
                // arithmetic with Bool result goes through the same Float32 path and is cast to bool
                if (in.op == AXB_OP_BIT_AND || in.op == AXB_OP_BIT_OR || in.op == AXB_OP_BIT_XOR ||
                    in.op == AXB_OP_SHL || in.op == AXB_OP_SHR)
                    return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: bitwise ops on Bool are not supported");
                ct = AXB_F32;
            } else {
                ct = rdt;
            }
            bool swap = false;
            int s0 = in.src[0], s1 = in.src[1];
            bool s1_in_acc = (s1 < 0 && s1 != AXB_SRC_IMM && s1 != AXB_SRC_NONE && acc_holds == -1 - s1);
            bool s0_in_acc = (s0 < 0 && s0 != AXB_SRC_IMM && s0 != AXB_SRC_NONE && acc_holds == -1 - s0);
            if (s1_in_acc && !s0_in_acc) {
                // keep the resident value in acc, bring operand 0 into t and evaluate f(t, acc)
                swap = true;
                AXB_TRY(fetch(s0, in.imm, R_T, ct, &d0));
                B.cvt(d0, ct, R_T);
                AXB_TRY(fetch(s1, in.imm, R_ACC, ct, &d1));
                B.cvt(d1, ct, R_ACC);
            } else {
                AXB_TRY(fetch(s1, in.imm, R_T, ct, &d1));
                B.cvt(d1, ct, R_T);
                AXB_TRY(fetch(s0, in.imm, R_ACC, ct, &d0));
                B.cvt(d0, ct, R_ACC);
            }
            uint32_t m = (oi.kind == OpInfo::ARITH) ? binary_micro(in.op, ct, same_shape_contig(in.src[0], in.src[1])) : cmp_micro(in.op);
            if (!m)
                return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: binary op %d on dtype %d has no GPU kernel", in.op, ct);
            if (swap)
                flags |= FL_SWAP;
            if (oi.kind == OpInfo::ARITH && is_half_dtype(ct))
                flags |= (ct == AXB_BF16 ? RND_BF16 : RND_F16);
            B.emit(m, 0, 0, flags);
            if (oi.kind == OpInfo::ARITH) {
                if (rdt == AXB_BOOL)
                    B.cvt(AXB_F32, AXB_BOOL, R_ACC);
                else if (cls_of(ct) == C_I && ct != AXB_I32 && ct != AXB_U32)
                    B.cvt(AXB_I32, ct, R_ACC);
            }
        }
        {
            std::vector<int64_t> sh = src_shape(in.src[0]);
            for (int j = 1; j < oi.arity; ++j)
                sh = bshape2(sh, src_shape(in.src[j]));
            vshape[i] = sh;
        }
        acc_holds = i;
        // release slots whose value is dead after this instruction
        for (int v = 0; v < i; ++v)
            if (slot_of[v] >= 0 && last_use[v] <= i) {
                slot_busy[slot_of[v]] = false;
                slot_of[v] = -1;
            }
        // keep value i in a slot if anything later than instr i+1 needs it (instr i+1 can always take it
        // from acc: operands destined for t/u are fetched before acc is overwritten)
        if (i < n_instr - 1 && last_use[i] > i + 1) {
            int sl = -1;
            for (int k = 0; k < kNSlot; ++k)
                if (!slot_busy[k]) {
                    sl = k;
                    break;
                }
            if (sl < 0)
                return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: more than %d live intermediates", kNSlot);
            slot_busy[sl] = true;
            slot_of[i] = sl;
            B.emit(M_MOV, R_ACC, R_S0 + sl);
        }
    }
    AXB_TRY(B.st);
    if (emit_store)
        B.emit(M_STORE);
    AXB_TRY(B.st);
    *code_out = B.code;
    *loads_out = B.loads;
    *final_dtype = prog[n_instr - 1].dtype;
    return AXB_OK;
}

// Collapse the broadcast iteration space and fill Params operands. ops[0..n_in) inputs, ops[n_in] = out (may
// be null for reduce). Returns inner_len/outer dims in p.
axb_status plan_iteration(const axb_tensor *inputs, int n_inputs, const axb_tensor *out, const int64_t *oshape,
                          int ondim, int vec, Params &p) {
    const int n_ops = n_inputs + (out ? 1 : 0);
    // expanded byte strides per operand, right-aligned to ondim
    std::vector<std::vector<int64_t>> st(n_ops, std::vector<int64_t>(ondim, 0));
    for (int k = 0; k < n_ops; ++k) {
        const axb_tensor &t = (k < n_inputs) ? inputs[k] : *out;
        int shift = ondim - t.ndim;
        if (shift < 0)
            return fail(AXB_ERR_SHAPE, "axb_elementwise: operand %d has more dims than the output", k);
        for (int d = 0; d < t.ndim; ++d) {
            int od = d + shift;
            if (t.shape[d] == oshape[od])
                st[k][od] = (oshape[od] == 1) ? 0 : t.strides[d];
            else if (t.shape[d] == 1)
                st[k][od] = 0;
            else
                return fail(AXB_ERR_SHAPE, "axb_elementwise: operand %d dim %d (%lld) does not broadcast to %lld", k, d,
                            (long long)t.shape[d], (long long)oshape[od]);
        }
    }
    // drop size-1 dims, then merge adjacent dims (outer i into inner j) when every operand is dense across them
    std::vector<int64_t> shp;
    std::vector<std::vector<int64_t>> cs(n_ops);
    for (int d = 0; d < ondim; ++d) {
        if (oshape[d] == 1)
            continue;
        shp.push_back(oshape[d]);
        for (int k = 0; k < n_ops; ++k)
            cs[k].push_back(st[k][d]);
    }
    if (shp.empty()) {
        shp.push_back(1);
        for (int k = 0; k < n_ops; ++k)
            cs[k].push_back(0);
    }
    for (int d = (int)shp.size() - 2; d >= 0; --d) {
        bool ok = true;
        for (int k = 0; k < n_ops && ok; ++k)
            ok = (cs[k][d] == cs[k][d + 1] * shp[d + 1]);
        if (ok) {
            shp[d + 1] *= shp[d];
            shp.erase(shp.begin() + d);
            for (int k = 0; k < n_ops; ++k)
                cs[k].erase(cs[k].begin() + d);
        }
    }
    const int nd = (int)shp.size();
    if (nd - 1 > kMaxOuter)
        return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: more than %d non-collapsible dims", kMaxOuter + 1);
    const int64_t inner = shp[nd - 1];
    if (inner >= (1ll << 31))
        return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: innermost run of %lld elements too long", (long long)inner);
    int64_t nvec_inner = (inner + vec - 1) / vec;
    int64_t total = nvec_inner;
    for (int d = 0; d < nd - 1; ++d)
        total *= shp[d];
    if (total >= (1ll << 31))
        return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: %lld vectors exceed the 2^31 per-launch limit", (long long)total);
    p.inner_len = (uint32_t)inner;
    p.nvec_inner = (uint32_t)nvec_inner;
    p.total_vec = (uint32_t)total;
    p.n_outer = nd - 1;
    FastDiv f0 = make_fastdiv((uint32_t)nvec_inner);
    p.div_m[0] = f0.m;
    p.div_s[0] = f0.s;
    for (int d = 0; d < kMaxOuter; ++d) {
        p.outer_shape[d] = 1;
        p.div_m[1 + d] = 1;
        p.div_s[1 + d] = 0;
    }
    for (int d = 0; d < nd - 1; ++d) { // outer index 0 = dim adjacent to the inner run
        int64_t e = shp[nd - 2 - d];
        if (e >= (1ll << 31))
            return fail(AXB_ERR_UNSUPPORTED, "axb_elementwise: dim too large");
        p.outer_shape[d] = (uint32_t)e;
        FastDiv f = make_fastdiv((uint32_t)e);
        p.div_m[1 + d] = f.m;
        p.div_s[1 + d] = f.s;
    }
    for (int k = 0; k < n_ops; ++k) {
        const axb_tensor &t = (k < n_inputs) ? inputs[k] : *out;
        Operand &o = (k < n_inputs) ? p.in[k] : p.out;
        memset(&o, 0, sizeof(o));
        o.ptr = tensor_ptr(t);
        o.dtype = (uint8_t)t.dtype;
        o.inner_stride = cs[k][nd - 1];
        for (int d = 0; d < nd - 1; ++d)
            o.stride[d] = cs[k][nd - 2 - d];
        const int64_t sz = (int64_t)dtype_size(t.dtype);
        const int64_t vbytes = std::min<int64_t>(16, (int64_t)vec * sz);
        bool aligned = ((uintptr_t)o.ptr % vbytes) == 0;
        for (int d = 0; d < nd - 1 && aligned; ++d)
            aligned = (o.stride[d] % vbytes) == 0;
        if (o.inner_stride == 0 && k < n_inputs)
            o.mode = MODE_BCAST;
        else if (o.inner_stride == sz && aligned)
            o.mode = MODE_VEC;
        else
            o.mode = MODE_STRIDED;
    }
    return AXB_OK;
}

axb_status broadcast_all(const axb_tensor *inputs, int n_inputs, int64_t *oshape, int *ondim) {
    int nd = 0;
    for (int k = 0; k < n_inputs; ++k)
        nd = std::max(nd, inputs[k].ndim);
    for (int d = 0; d < nd; ++d)
        oshape[d] = 1;
    for (int k = 0; k < n_inputs; ++k) {
        int shift = nd - inputs[k].ndim;
        for (int d = 0; d < inputs[k].ndim; ++d) {
            int64_t e = inputs[k].shape[d];
            int64_t &o = oshape[d + shift];
            if (e == o || e == 1)
                continue;
            if (o == 1)
                o = e;
            else
                return fail(AXB_ERR_SHAPE, "axb_elementwise: shapes are not broadcastable (%lld vs %lld)", (long long)e,
                            (long long)o);
        }
    }
    *ondim = nd;
    return AXB_OK;
}

// Program-specific tail of the translation unit of one specialised kernel (jit.cpp prepends the shared prelude).
std::string kernel_source(const std::string &loads, const std::string &body, bool wide, int vec, int n_outer, int n_inputs,
                          int unroll, int reduce_opk = -1) {
    std::string s;
    char hdr[200];
    if (reduce_opk < 0)
        snprintf(hdr, sizeof hdr, "\nEW_KERNEL_BEGIN(axb_ew, %d, %s, %d, %d, %d)\n", vec, wide ? "uint64_t" : "uint32_t", n_outer,
                 n_inputs, unroll);
    else if (reduce_opk < 4) // chain + full reduction: the body ends in EW_RED, the kernel in a CTA reduce (ew_prelude.cuh)
        snprintf(hdr, sizeof hdr, "\nEW_RKERNEL_BEGIN(axb_ew_red, %d, %s, %d, %d, %d, %d)\n", vec, wide ? "uint64_t" : "uint32_t",
                 n_outer, n_inputs, unroll, reduce_opk);
    else if (reduce_opk < 64) // chain + last-axis reduction: reduce_opk = 4 * mode + op, mode 1 = one CTA per row, 2 = one warp per row
        snprintf(hdr, sizeof hdr, "\nEW_ROWKERNEL_BEGIN(axb_ew_rowred, %d, %s, %d, %d, %d, %d, %d)\n", vec, wide ? "uint64_t" : "uint32_t",
                 n_outer, n_inputs, unroll, reduce_opk & 3, (reduce_opk >> 2) == 1 ? 256 : 32);
    else { // chain + softmax / norm over the last axis: 64 + (NV - 1) + 8 * (G == 256) + 16 * mode + 64 * epilogue
        const int k = reduce_opk - 64;
        snprintf(hdr, sizeof hdr, "\nEW_SMKERNEL_BEGIN(axb_ew_softmax, %d, %s, %d, %d, %d, %d, %d, %d)\n", vec, wide ? "uint64_t" : "uint32_t",
                 n_outer, n_inputs, (k & 7) + 1, (k & 8) ? 256 : 32, (k >> 4) & 3, (k >> 6) & 3);
    }
    s += hdr;
    s += "        EW_PRELOAD_BEGIN\n" + loads + "        EW_PRELOAD_END\n";
    s += "        EW_COMPUTE_BEGIN\n" + body + "        EW_COMPUTE_END\n";
    s += reduce_opk < 0 ? "EW_KERNEL_END\n" : reduce_opk < 4 ? "EW_RKERNEL_END\n" : reduce_opk < 64 ? "EW_ROWKERNEL_END\n" : "EW_SMKERNEL_END\n";
    return s;
}

struct Prepared {
    Params p;
    std::string source;
    int unroll = 1;
};

// Validates, plans the iteration space and generates the kernel text. No device needed.
axb_status prepare(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, const axb_tensor *out,
                   Prepared *P, bool *empty, int reduce_opk = -1) {
    AXB_REQUIRE(prog && n_instr > 0 && n_instr <= AXB_MAX_INSTRS, AXB_ERR_INVALID,
                "axb_elementwise: need 1..%d instructions", AXB_MAX_INSTRS);
    AXB_REQUIRE(n_inputs >= 0 && n_inputs <= AXB_MAX_INPUTS, AXB_ERR_INVALID, "axb_elementwise: at most %d inputs",
                AXB_MAX_INPUTS);
    AXB_REQUIRE(n_inputs == 0 || inputs, AXB_ERR_INVALID, "axb_elementwise: null inputs");
    AXB_TRY(validate_tensor(out, "axb_elementwise(out)"));
    bool wide = is_64bit_dtype(out->dtype);
    for (int k = 0; k < n_inputs; ++k) {
        AXB_TRY(validate_tensor(&inputs[k], "axb_elementwise(input)"));
        AXB_REQUIRE(!is_complex_dtype(inputs[k].dtype), AXB_ERR_UNSUPPORTED,
                    "axb_elementwise: complex dtypes have no GPU path");
        wide = wide || is_64bit_dtype(inputs[k].dtype);
    }
    for (int i = 0; i < n_instr; ++i)
        wide = wide || is_64bit_dtype(prog[i].dtype);
    // output shape must be the broadcast of the inputs (or of out itself when there are no inputs)
    int64_t oshape[AXB_MAX_DIMS];
    int ondim = 0;
    if (n_inputs > 0) {
        AXB_TRY(broadcast_all(inputs, n_inputs, oshape, &ondim));
        AXB_REQUIRE(out->ndim >= ondim, AXB_ERR_SHAPE, "axb_elementwise: out rank %d < broadcast rank %d", out->ndim,
                    ondim);
        for (int d = 0; d < out->ndim; ++d) {
            int id = d - (out->ndim - ondim);
            int64_t want = id >= 0 ? oshape[id] : 1;
            AXB_REQUIRE(out->shape[d] == want || want == 1, AXB_ERR_SHAPE,
                        "axb_elementwise: out dim %d is %lld, broadcast of inputs gives %lld", d,
                        (long long)out->shape[d], (long long)want);
        }
    }
    ondim = out->ndim;
    for (int d = 0; d < ondim; ++d)
        oshape[d] = out->shape[d];
    *empty = numel(*out) == 0;
    if (*empty)
        return AXB_OK;
    Params &p = P->p;
    memset(&p, 0, sizeof(p));
    // elements per vector: 16 bytes of the widest element type that is streamed (so its access is ONE 128-bit
    // instruction per thread and a warp covers a contiguous 512-byte span); narrower operands use 64/32-bit accesses
    size_t widest = dtype_size(out->dtype);
    for (int k = 0; k < n_inputs; ++k)
        if (numel(inputs[k]) > 1)
            widest = std::max(widest, dtype_size(inputs[k].dtype));
    const int vec = wide ? 2 : (int)(16 / std::max<size_t>(widest, 1));
    AXB_TRY(plan_iteration(inputs, n_inputs, out, oshape, ondim, vec, p));
    int fdt = 0;
    std::string body, loads;
    AXB_TRY(lower_program(prog, n_instr, inputs, n_inputs, wide, p, &fdt, reduce_opk < 0, &body, &loads));
    AXB_REQUIRE(fdt == out->dtype, AXB_ERR_INVALID, "axb_elementwise: last instruction yields dtype %d but out is %d", fdt,
                out->dtype);
    if (reduce_opk >= 64) {
        body += "            EW_KEEP\n";
    } else if (reduce_opk >= 0) {
        char line[64];
        snprintf(line, sizeof line, "            EW_RED(%d)\n", reduce_opk & 3);
        body += line;
    }
    // loads in flight: U vectors per thread while the streamed (non-broadcast) operands fit the register budget
    int streamed = 0;
    for (int k = 0; k < n_inputs; ++k)
        streamed += p.in[k].mode != MODE_BCAST;
    // ~16 elements (8 for 64-bit words) per thread, fewer when many operands are streamed (register budget)
    // measured on B200 (C2 chain: U=1 5.76, U=2 6.03, U=4 4.07 TB/s; single-input ops flat from U=1..8)
    P->unroll = streamed <= 1 ? 4 : streamed <= 4 ? 2 : 1;
    if (const char *env = getenv("AXB_EW_UNROLL")) // tuning knob (dev): overrides the vectors-per-thread heuristic
        P->unroll = std::max(1, std::min(8, atoi(env)));
    P->source = kernel_source(loads, body, wide, vec, (int)p.n_outer, n_inputs, P->unroll, reduce_opk);
    return AXB_OK;
}


// ---- launch-plan cache -------------------------------------------------------------------------------------------------
// A fused-chain launch is fully determined by the program (ops, dtypes, operand selectors, immediates) and by the LAYOUT
// of every operand (dtype, shape, byte strides, 16-byte alignment class of its address) — never by the addresses
// themselves. prepare() (validation, iteration-space collapse, statement generation, kernel lookup by source hash) costs
// tens of microseconds per call, as much as the kernels the graph lowering launches; the plan cache keeps its result
// (kernel handle + Params with stale pointers + grid) under a hash of exactly those inputs, and a hit only patches the
// operand pointers. This is the per-launch half of the plan cache of SURVEY.md §8 a3 (the graph-level half lives in the
// C++ glue, host/cuda_graph.cpp); reference counterpart: the per-signature MPSGraph cache, mpsgraph_fused.mm:373-377.
struct PlanKeyOperand {
    int32_t dtype, ndim;
    uint32_t align; // address & 15
    int64_t shape[AXB_MAX_DIMS], strides[AXB_MAX_DIMS];
};
struct PlanEntry {
    std::vector<unsigned char> key;
    Params p;
    jit::Kernel kernel;
    int unroll, device;
    uint64_t tick;
};
std::mutex g_plan_mu;
std::unordered_map<uint64_t, std::vector<PlanEntry>> g_plans; // hash -> entries (full key compared)
size_t g_plan_count = 0;
uint64_t g_plan_tick = 0;
axb_plan_cache_stats g_plan_stats{};
constexpr size_t kMaxPlans = 4096;

void plan_key(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, const axb_tensor *out, int reduce_opk,
              int device, std::vector<unsigned char> *key, uint64_t *hash) {
    key->clear();
    auto put = [&](const void *q, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(q);
        key->insert(key->end(), b, b + n);
    };
    const int32_t head[4] = {n_instr, n_inputs, reduce_opk, device};
    put(head, sizeof head);
    for (int i = 0; i < n_instr; ++i) {
        put(&prog[i].op, 4);
        put(&prog[i].dtype, 4);
        put(prog[i].src, 12);
        put(&prog[i].imm, 8);
    }
    for (int k = 0; k <= n_inputs; ++k) {
        const axb_tensor &t = k < n_inputs ? inputs[k] : *out;
        const int32_t h2[2] = {t.dtype, t.ndim};
        put(h2, sizeof h2);
        const uint32_t al = (uint32_t)(reinterpret_cast<uintptr_t>(tensor_ptr(t)) & 15u);
        put(&al, 4);
        put(t.shape, sizeof(int64_t) * (size_t)std::max(t.ndim, 0));
        put(t.strides, sizeof(int64_t) * (size_t)std::max(t.ndim, 0));
    }
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : *key) {
        h ^= c;
        h *= 1099511628211ull;
    }
    *hash = h;
}

// Looks the launch up (or prepares and inserts it). On return P holds Params with CURRENT pointers, *k the kernel.
axb_status plan_lookup(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, const axb_tensor *out,
                       int reduce_opk, const char *name, int device, Prepared *P, jit::Kernel *k, bool *empty) {
    *empty = false;
    static const bool off = getenv("AXB_NO_PLAN_CACHE") != nullptr; // A/B switch
    bool sane = !off && prog && out && n_instr > 0 && n_instr <= AXB_MAX_INSTRS && n_inputs >= 0 && n_inputs <= AXB_MAX_INPUTS &&
                (n_inputs == 0 || inputs) && out->ndim >= 0 && out->ndim <= AXB_MAX_DIMS && out->base;
    for (int i = 0; sane && i < n_inputs; ++i)
        sane = inputs[i].ndim >= 0 && inputs[i].ndim <= AXB_MAX_DIMS && inputs[i].base;
    std::vector<unsigned char> key;
    uint64_t h = 0;
    if (sane) {
        plan_key(prog, n_instr, inputs, n_inputs, out, reduce_opk, device, &key, &h);
        std::lock_guard<std::mutex> lk(g_plan_mu);
        auto it = g_plans.find(h);
        if (it != g_plans.end())
            for (auto &e : it->second)
                if (e.key == key) {
                    e.tick = ++g_plan_tick;
                    g_plan_stats.hits++;
                    P->p = e.p;
                    P->unroll = e.unroll;
                    for (int i = 0; i < n_inputs; ++i)
                        P->p.in[i].ptr = tensor_ptr(inputs[i]);
                    P->p.out.ptr = tensor_ptr(*out);
                    *k = e.kernel;
                    return AXB_OK;
                }
    }
    // miss (or an argument set the validation inside prepare() has to reject with its own message)
    AXB_TRY(prepare(prog, n_instr, inputs, n_inputs, out, P, empty, reduce_opk));
    if (*empty)
        return AXB_OK;
    AXB_TRY(jit::get_kernel(P->source, name, device, k));
    if (sane) {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        g_plan_stats.misses++;
        if (g_plan_count >= kMaxPlans) { // full: drop the least recently used half
            std::vector<uint64_t> ticks;
            for (auto &kv : g_plans)
                for (auto &e : kv.second)
                    ticks.push_back(e.tick);
            std::nth_element(ticks.begin(), ticks.begin() + ticks.size() / 2, ticks.end());
            const uint64_t cut = ticks[ticks.size() / 2];
            for (auto it = g_plans.begin(); it != g_plans.end();) {
                auto &v = it->second;
                v.erase(std::remove_if(v.begin(), v.end(), [&](const PlanEntry &e) { return e.tick < cut; }), v.end());
                it = v.empty() ? g_plans.erase(it) : std::next(it);
            }
            g_plan_count = 0;
            for (auto &kv : g_plans)
                g_plan_count += kv.second.size();
            g_plan_stats.evictions++;
        }
        PlanEntry e;
        e.key = std::move(key);
        e.p = P->p;
        e.kernel = *k;
        e.unroll = P->unroll;
        e.device = device;
        e.tick = ++g_plan_tick;
        g_plans[h].push_back(std::move(e));
        g_plan_stats.entries = ++g_plan_count;
    }
    return AXB_OK;
}

} // namespace ew
} // namespace axb

using namespace axb;
using namespace axb::ew;

extern "C" {

axb_status axb_elementwise(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                           const axb_tensor *out, axb_stream_t stream) {
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    {
        bool via_table = false; // one half-precision tensor through a transcendental: 65 536-entry lookup (half_lut.cu)
        AXB_TRY(lut::try_run(prog, n_instr, inputs, n_inputs, out, s, ctx, &via_table));
        if (via_table)
            return AXB_OK;
    }
    Prepared P;
    bool empty = false;
    jit::Kernel k;
    AXB_TRY(plan_lookup(prog, n_instr, inputs, n_inputs, out, -1, "axb_ew", ctx->device, &P, &k, &empty));
    if (empty)
        return AXB_OK;
    const int threads = 256;
    int64_t blocks = ((int64_t)P.p.total_vec + (int64_t)threads * P.unroll - 1) / ((int64_t)threads * P.unroll);
    if (blocks < 1)
        blocks = 1;
    void *args[] = {&P.p};
    AXB_TRY(jit::launch(k, (unsigned)blocks, threads, 0, s, args));
    count_launch();
    return AXB_OK;
}

axb_status axb_get_plan_cache_stats(axb_plan_cache_stats *out) {
    AXB_REQUIRE(out, AXB_ERR_INVALID, "axb_get_plan_cache_stats: null out");
    std::lock_guard<std::mutex> lk(g_plan_mu);
    *out = g_plan_stats;
    return AXB_OK;
}
axb_status axb_plan_cache_clear(void) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    g_plans.clear();
    g_plan_count = 0;
    g_plan_stats = axb_plan_cache_stats{};
    return AXB_OK;
}

} // extern "C"

// Chain + full reduction, first pass (host_api.cpp axb_elementwise_reduce): `shape_desc` describes the chain's (never
// materialised) fp32 value; one fp32 partial per CTA goes to a pooled workspace returned in *partials / *n_partials.
namespace axb {
namespace ew {
axb_status launch_chain_reduce(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                               const axb_tensor *shape_desc, int reduce_opk, cudaStream_t s, DeviceCtx *ctx, void **partials,
                               int64_t *n_partials) {
    Prepared P;
    bool empty = false;
    axb_tensor od = *shape_desc;
    od.base = reinterpret_cast<void *>(uintptr_t(256)); // never dereferenced: the kernel only writes the partials
    jit::Kernel k;
    AXB_TRY(plan_lookup(prog, n_instr, inputs, n_inputs, &od, reduce_opk, "axb_ew_red", ctx->device, &P, &k, &empty));
    AXB_REQUIRE(!empty, AXB_ERR_INVALID, "axb_elementwise_reduce: empty chain");
    const int threads = 256;
    int64_t blocks = ((int64_t)P.p.total_vec + (int64_t)threads * P.unroll - 1) / ((int64_t)threads * P.unroll);
    if (blocks < 1)
        blocks = 1;
    void *ws = nullptr;
    AXB_TRY(axb_alloc((size_t)blocks * sizeof(float), s, &ws));
    P.p.out.ptr = static_cast<char *>(ws);
    void *args[] = {&P.p};
    axb_status st = jit::launch(k, (unsigned)blocks, threads, 0, s, args);
    if (st != AXB_OK) {
        axb_free(ws, s);
        return st;
    }
    count_launch();
    *partials = ws;
    *n_partials = blocks;
    return AXB_OK;
}

// Padded-pitch descriptor of a chain value that is never written: no operand set is dense across its last two dims any more, so
// plan_iteration keeps the last axis as the innermost run while the outer dims still collapse.
static axb_tensor row_desc(const axb_tensor &shape_desc) {
    axb_tensor od = shape_desc;
    od.base = reinterpret_cast<void *>(uintptr_t(256));
    od.byte_offset = 0;
    const int nd = od.ndim;
    const int64_t esz = (int64_t)dtype_size(od.dtype);
    int64_t pitch = od.shape[nd - 1] * esz + 16;
    od.strides[nd - 1] = esz;
    for (int d = nd - 2; d >= 0; --d) {
        od.strides[d] = pitch;
        pitch *= std::max<int64_t>(od.shape[d], 1);
    }
    return od;
}

// Chain + softmax / log_softmax over the last axis [+ row sum / mean / max of the result] in one pass
// (axb_elementwise_softmax): rows of at most 256 * 8 vectors; `out` holds the fp32 rows (epilogue 0) or one fp32 per row.
axb_status launch_chain_softmax(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                const axb_tensor *shape_desc, int mode, int epilogue, float *out, cudaStream_t s, DeviceCtx *ctx,
                                bool *done, const float *weight, const float *bias, float eps) {
    *done = false;
    const int nd = shape_desc->ndim;
    if (nd < 1 || shape_desc->shape[nd - 1] < 2 || dtype_size(shape_desc->dtype) != 4)
        return AXB_OK;
    const int64_t len = shape_desc->shape[nd - 1];
    int64_t rows = 1;
    for (int d = 0; d < nd - 1; ++d)
        rows *= shape_desc->shape[d];
    const int64_t nvec = (len + 3) / 4; // fp32 chains: 4 elements per vector
    if (rows == 0 || rows >= (1ll << 31) || nvec > 256 * 8)
        return AXB_OK;
    const bool cta = nvec > 32 * 8 || (nvec > 64 && rows < 4096); // one warp per row while 8 vectors per lane cover it
    const int g = cta ? 256 : 32;
    const int nv = (int)((nvec + g - 1) / g);
    const axb_tensor od = row_desc(*shape_desc);
    Prepared P;
    bool empty = false;
    jit::Kernel k;
    const int code = 64 + (nv - 1) + (cta ? 8 : 0) + 16 * mode + 64 * epilogue;
    AXB_TRY(plan_lookup(prog, n_instr, inputs, n_inputs, &od, code, "axb_ew_softmax", ctx->device, &P, &k, &empty));
    if (empty || (int64_t)P.p.inner_len != len || (int64_t)(P.p.total_vec / P.p.nvec_inner) != rows || (int64_t)P.p.nvec_inner != nvec)
        return AXB_OK;
    P.p.reduce_ws = out;
    P.p.aux0 = weight;
    P.p.aux1 = bias;
    P.p.eps = eps;
    const int64_t blocks = cta ? rows : (rows + 7) / 8;
    void *args[] = {&P.p};
    AXB_TRY(jit::launch(k, (unsigned)blocks, 256, 0, s, args));
    count_launch();
    *done = true;
    return AXB_OK;
}

// Chain + reduction over the LAST axis of the chain's value in one pass (host_api.cpp axb_elementwise_reduce_axis): `out`
// holds one fp32 per row of `shape_desc`. *done = false leaves the call to the unfused path (the planner could not keep the
// reduced axis as the innermost run on its own, or the row count does not fit a grid).
axb_status launch_chain_rowreduce(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                  const axb_tensor *shape_desc, int opk, bool is_mean, float *out, cudaStream_t s, DeviceCtx *ctx,
                                  bool *done) {
    *done = false;
    const int nd = shape_desc->ndim;
    if (nd < 1 || shape_desc->shape[nd - 1] < 2)
        return AXB_OK;
    // a descriptor of the (never written) chain value whose rows are PADDED: no operand set is dense across the last two
    // dims any more, so plan_iteration keeps the reduced axis as the innermost run while outer dims still collapse
    axb_tensor od = *shape_desc;
    od.base = reinterpret_cast<void *>(uintptr_t(256));
    od.byte_offset = 0;
    const int64_t esz = (int64_t)dtype_size(od.dtype);
    int64_t pitch = od.shape[nd - 1] * esz + 16;
    od.strides[nd - 1] = esz;
    for (int d = nd - 2; d >= 0; --d) {
        od.strides[d] = pitch;
        pitch *= std::max<int64_t>(od.shape[d], 1);
    }
    const int64_t len = od.shape[nd - 1];
    int64_t rows = 1;
    for (int d = 0; d < nd - 1; ++d)
        rows *= od.shape[d];
    if (rows == 0 || rows >= (1ll << 31))
        return AXB_OK;
    Prepared P;
    bool empty = false;
    jit::Kernel k;
    // one warp per row while a warp covers the row in at most two trips (VEC = 4 for fp32 chains, U <= 4)
    const int mode = len <= 1024 ? 2 : 1;
    AXB_TRY(plan_lookup(prog, n_instr, inputs, n_inputs, &od, 4 * mode + opk, "axb_ew_rowred", ctx->device, &P, &k, &empty));
    if (empty || (int64_t)P.p.inner_len != len || (int64_t)(P.p.total_vec / P.p.nvec_inner) != rows)
        return AXB_OK;
    P.p.reduce_op = is_mean ? AXB_OP_MEAN : 0;
    P.p.reduce_ws = out;
    const int64_t blocks = mode == 1 ? rows : (rows + 7) / 8;
    void *args[] = {&P.p};
    AXB_TRY(jit::launch(k, (unsigned)blocks, 256, 0, s, args));
    count_launch();
    *done = true;
    return AXB_OK;
}
} // namespace ew
} // namespace axb

extern "C" {

// Compile-only entry (no device needed): makes sure the specialised kernel of this program signature is in
// the in-tree cubin cache. build() calls it for the BASELINE configs so the named hot path never JITs.
axb_status axb_elementwise_precompile(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                      const axb_tensor *out) {
    Prepared P;
    bool empty = false;
    AXB_TRY(prepare(prog, n_instr, inputs, n_inputs, out, &P, &empty));
    if (empty)
        return AXB_OK;
    return jit::precompile(P.source);
}

// Same for the chain + reduction kernels: reduce_kind 0 = full reduction (axb_elementwise_reduce), 1 = last axis with one CTA per
// row, 2 = last axis with one warp per row (axb_elementwise_reduce_axis); reduce_op sum / mean / max / min.
axb_status axb_elementwise_reduce_precompile(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                             const axb_tensor *shape_desc, int32_t reduce_op, int reduce_kind) {
    AXB_REQUIRE(shape_desc && reduce_kind >= 0, AXB_ERR_INVALID, "axb_elementwise_reduce_precompile: bad arguments");
    if (reduce_kind >= 64) { // a chain + softmax kernel code (see launch_chain_softmax), for build-time checks
        const axb_tensor od = row_desc(*shape_desc);
        Prepared P;
        bool empty = false;
        AXB_TRY(prepare(prog, n_instr, inputs, n_inputs, &od, &P, &empty, reduce_kind));
        return empty ? AXB_OK : jit::precompile(P.source);
    }
    AXB_REQUIRE(reduce_kind <= 2, AXB_ERR_INVALID, "axb_elementwise_reduce_precompile: bad reduce_kind");
    const int opk = (reduce_op == AXB_OP_SUM || reduce_op == AXB_OP_MEAN) ? 0 : reduce_op == AXB_OP_MAX ? 1 : reduce_op == AXB_OP_MIN ? 2 : -1;
    AXB_REQUIRE(opk >= 0, AXB_ERR_UNSUPPORTED, "axb_elementwise_reduce_precompile: sum / mean / max / min only");
    axb_tensor od = *shape_desc;
    od.base = reinterpret_cast<void *>(uintptr_t(256));
    od.byte_offset = 0;
    if (reduce_kind > 0 && od.ndim >= 1) { // padded rows, exactly as launch_chain_rowreduce plans them
        const int64_t esz = (int64_t)dtype_size(od.dtype);
        int64_t pitch = od.shape[od.ndim - 1] * esz + 16;
        od.strides[od.ndim - 1] = esz;
        for (int d = od.ndim - 2; d >= 0; --d) {
            od.strides[d] = pitch;
            pitch *= std::max<int64_t>(od.shape[d], 1);
        }
    }
    Prepared P;
    bool empty = false;
    AXB_TRY(prepare(prog, n_instr, inputs, n_inputs, &od, &P, &empty, 4 * reduce_kind + opk));
    if (empty)
        return AXB_OK;
    return jit::precompile(P.source);
}

axb_status axb_get_jit_stats(axb_jit_stats *out) {
    AXB_REQUIRE(out, AXB_ERR_INVALID, "axb_get_jit_stats: null");
    jit::Stats s = jit::stats();
    out->mem_hits = s.mem_hits;
    out->disk_hits = s.disk_hits;
    out->compiles = s.compiles;
    return AXB_OK;
}

} // extern "C"
