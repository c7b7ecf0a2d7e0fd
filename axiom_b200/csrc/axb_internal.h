// axb_internal.h — shared host-side plumbing for the axb C ABI implementation (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "axb.h"

namespace axb {

// ---- error reporting (never throw across the C ABI) ------------------------------------------------
std::string &last_error_ref();
axb_status fail(axb_status st, const char *fmt, ...);

#define AXB_CUDA(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return ::axb::fail(_e == cudaErrorMemoryAllocation ? AXB_ERR_OOM : AXB_ERR_CUDA,            \
                               "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,        \
                               __LINE__);                                                               \
    } while (0)

#define AXB_TRY(expr)                                                                                   \
    do {                                                                                                \
        axb_status _s = (expr);                                                                         \
        if (_s != AXB_OK)                                                                               \
            return _s;                                                                                  \
    } while (0)

#define AXB_REQUIRE(cond, st, ...)                                                                      \
    do {                                                                                                \
        if (!(cond))                                                                                    \
            return ::axb::fail(st, __VA_ARGS__);                                                        \
    } while (0)

// ---- device context --------------------------------------------------------------------------------
struct DeviceCtx {
    int device = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr; // the backend stream (non-blocking)
};
// Context of the calling thread's current device (lazily created). nullptr + error set on failure.
DeviceCtx *current_ctx();
// Context of device `dev` if it has been created (never creates one, never touches the thread's current device).
DeviceCtx *ctx_of_device(int dev);
// Resolve a user stream handle: NULL -> backend stream of the current device.
axb_status resolve_stream(axb_stream_t s, cudaStream_t *out, DeviceCtx **ctx = nullptr);

// ---- launch accounting -----------------------------------------------------------------------------
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- host<->device transfer accounting (axb_get_transfer_stats) ------------------------------------
struct TransferCounters {
    std::atomic<uint64_t> h2d_bytes{0}, h2d_calls{0}, d2h_bytes{0}, d2h_calls{0}, d2d_bytes{0}, d2d_calls{0};
};
extern TransferCounters g_xfer;
inline void count_h2d(size_t n) { g_xfer.h2d_bytes.fetch_add(n, std::memory_order_relaxed); g_xfer.h2d_calls.fetch_add(1, std::memory_order_relaxed); }
inline void count_d2h(size_t n) { g_xfer.d2h_bytes.fetch_add(n, std::memory_order_relaxed); g_xfer.d2h_calls.fetch_add(1, std::memory_order_relaxed); }
inline void count_d2d(size_t n) { g_xfer.d2d_bytes.fetch_add(n, std::memory_order_relaxed); g_xfer.d2d_calls.fetch_add(1, std::memory_order_relaxed); }

#define AXB_CHECK_LAUNCH()                                                                              \
    do {                                                                                                \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess)                                                                          \
            return ::axb::fail(AXB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                     \
        ::axb::count_launch();                                                                          \
    } while (0)

// ---- dtype helpers ---------------------------------------------------------------------------------
inline size_t dtype_size(int32_t dt) {
    switch (dt) {
    case AXB_BOOL: case AXB_I8: case AXB_U8: return 1;
    case AXB_I16: case AXB_U16: case AXB_F16: case AXB_BF16: return 2;
    case AXB_I32: case AXB_U32: case AXB_F32: return 4;
    case AXB_I64: case AXB_U64: case AXB_F64: case AXB_C64: return 8;
    case AXB_C128: return 16;
    default: return 0;
    }
}
inline bool is_float_dtype(int32_t dt) { return dt == AXB_F16 || dt == AXB_BF16 || dt == AXB_F32 || dt == AXB_F64; }
inline bool is_half_dtype(int32_t dt) { return dt == AXB_F16 || dt == AXB_BF16; }
inline bool is_int_dtype(int32_t dt) { return dt >= AXB_BOOL && dt <= AXB_U64; }
inline bool is_64bit_dtype(int32_t dt) { return dt == AXB_I64 || dt == AXB_U64 || dt == AXB_F64; }
inline bool is_complex_dtype(int32_t dt) { return dt == AXB_C64 || dt == AXB_C128; }
inline bool is_signed_int_dtype(int32_t dt) { return dt >= AXB_I8 && dt <= AXB_I64; }
inline bool is_unsigned_int_dtype(int32_t dt) { return dt >= AXB_U8 && dt <= AXB_U64; }

inline int64_t numel(const axb_tensor &t) {
    int64_t n = 1;
    for (int i = 0; i < t.ndim; ++i)
        n *= t.shape[i];
    return n;
}
inline bool is_contiguous(const axb_tensor &t) {
    int64_t expect = (int64_t)dtype_size(t.dtype);
    for (int i = t.ndim - 1; i >= 0; --i) {
        if (t.shape[i] != 1 && t.strides[i] != expect)
            return false;
        expect *= t.shape[i];
    }
    return true;
}
inline axb_status validate_tensor(const axb_tensor *t, const char *what) {
    if (!t)
        return fail(AXB_ERR_INVALID, "%s: null tensor", what);
    if (t->ndim < 0 || t->ndim > AXB_MAX_DIMS)
        return fail(AXB_ERR_SHAPE, "%s: ndim %d out of range (max %d)", what, t->ndim, AXB_MAX_DIMS);
    if (dtype_size(t->dtype) == 0)
        return fail(AXB_ERR_UNSUPPORTED, "%s: unknown dtype %d", what, t->dtype);
    for (int i = 0; i < t->ndim; ++i)
        if (t->shape[i] < 0)
            return fail(AXB_ERR_SHAPE, "%s: negative extent", what);
    if (numel(*t) > 0 && !t->base)
        return fail(AXB_ERR_INVALID, "%s: null base pointer", what);
    return AXB_OK;
}
inline char *tensor_ptr(const axb_tensor &t) { return static_cast<char *>(t.base) + t.byte_offset; }

// contiguous tensor descriptor helper
inline axb_tensor make_contiguous_desc(void *base, int32_t dtype, const int64_t *shape, int ndim) {
    axb_tensor t{};
    t.base = base;
    t.byte_offset = 0;
    t.ndim = ndim;
    t.dtype = dtype;
    int64_t st = (int64_t)dtype_size(dtype);
    for (int i = ndim - 1; i >= 0; --i) {
        t.shape[i] = shape[i];
        t.strides[i] = st;
        st *= shape[i];
    }
    return t;
}

// ---- fast integer division (host precompute, device use) -------------------------------------------
struct FastDiv {
    uint32_t d, m, s; // n / d == (umulhi(n, m) + n) >> s  for n < 2^31
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{d, 0, 0};
    if (d == 0)
        d = 1;
    uint32_t s = 0;
    while ((1ull << s) < d)
        ++s;
    uint64_t m = ((1ull << 32) * ((1ull << s) - d)) / d + 1;
    f.d = d;
    f.m = (uint32_t)m;
    f.s = s;
    return f;
}

// elementwise.cu: chain + full reduction, first pass (one fp32 partial per CTA into a pooled block)
namespace ew {
axb_status launch_chain_reduce(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                               const axb_tensor *shape_desc, int reduce_opk, cudaStream_t s, DeviceCtx *ctx, void **partials,
                               int64_t *n_partials);
axb_status launch_chain_rowreduce(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                  const axb_tensor *shape_desc, int opk, bool is_mean, float *out, cudaStream_t s, DeviceCtx *ctx,
                                  bool *done);
axb_status launch_chain_softmax(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                const axb_tensor *shape_desc, int mode, int epilogue, float *out, cudaStream_t s, DeviceCtx *ctx,
                                bool *done, const float *weight = nullptr, const float *bias = nullptr, float eps = 0.0f);
}
// half_lut.cu: table-lookup evaluation of a chain over one half-precision tensor (*done = false: not applicable)
namespace lut {
axb_status try_run(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, const axb_tensor *out, cudaStream_t s,
                   DeviceCtx *ctx, bool *done);
}
// reduce.cu: divisor used by the NEXT Mean reduction of this thread instead of its own element count (the partials of a
// fused chain + mean stand for `count` elements); consumed by that call
int64_t &reduce_count_override();
} // namespace axb
