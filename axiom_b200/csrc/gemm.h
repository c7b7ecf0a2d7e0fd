// gemm.h — normalised GEMM problem shared by the SIMT kernel (matmul.cu) and the tcgen05 kernel (gemm_tcgen05.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "axb_internal.h"

namespace axb {
namespace gemm {

constexpr int kMaxBatchDims = 6;

// C[b] = act(op(A)[b] . op(B)[b] + bias).  All strides are in BYTES; transposes are already folded into them.
struct Desc {
    const char *a, *b;
    char *c;
    const char *bias;
    int64_t M, N, K, batch;
    int64_t a_rs, a_cs; // A[m][k]
    int64_t b_rs, b_cs; // B[k][n]
    int64_t c_rs, c_cs; // C[m][n]
    int64_t bias_rs, bias_cs;
    int32_t nb;         // batch dims, index 0 = innermost
    int32_t act;        // 0 or a unary axb_op
    int32_t bias_dtype;
    int32_t const_ops;  // AXB_MM_CONST_A / _B: the operand's derived forms may be cached (see split cache)
    int64_t bshape[kMaxBatchDims];
    int64_t a_bs[kMaxBatchDims], b_bs[kMaxBatchDims], c_bs[kMaxBatchDims], bias_bs[kMaxBatchDims];
};

// Hint for the NEXT fp32 matmul issued by this thread: operand A has already been written in the three-piece bf16 layout
// [batch][M][3 * kp] (see gemm_tcgen05.cu try_split_f32), so its split pass is skipped and the tensor-core path is
// mandatory (the A tensor passed to axb_matmul only carries shape / batch information). Consumed by that call.
struct PresplitA {
    const uint16_t *ptr = nullptr;
    int64_t kp = 0;
};
PresplitA &presplit_a_hint();

int &last_path(); // per thread, see axb_matmul_last_path

// Cache of operand forms derived from CONSTANT operands (today: the three-piece bf16 split of fp32 weights). Entries are
// keyed by the operand's address and layout and dropped when the source memory is freed or written (allocator.cpp /
// runtime.cpp call invalidate_const_cache; the glue announces in-place kernels through axb_invalidate_cached).
void invalidate_const_cache(const void *ptr, size_t bytes);
axb_const_cache_stats const_cache_stats();
bool const_cache_active();

// Convolution as an implicit GEMM on the tensor cores (gemm_tcgen05.cu conv_tc): NC(H)W input, (C_out, C_in, KH, KW) weights,
// contiguous (N, C_out, OH, OW) output of the same dtype (float32 via the exact three-piece bf16 split, bf16, fp16),
// groups == 1. 1-D convolutions pass H = KH = OH = 1. Sets *done when it ran; small / unsupported problems are left alone.
struct ConvDesc {
    const char *x, *w, *bias;
    char *out;
    int32_t dtype, bias_dtype, act;
    int64_t N, Cin, H, W, Cout, KH, KW, OH, OW;
    int64_t x_s[4], w_s[4]; // byte strides
    int64_t bias_s;
    int32_t sh, sw, ph, pw, dh, dw; // dh / dw may be negative (the sub-convolutions of a transposed convolution walk their taps backwards)
    // optional: byte strides of the output over (N, C_out, OH, OW); all zero = contiguous. A transposed convolution writes each
    // output phase (oh % stride, ow % stride) as its own stride-1 convolution into an interleaved view of the result.
    int64_t o_s[4];
    int32_t force;  // skip the "too small to pay for the preparation passes" heuristics (the caller sized the whole problem)
    // optional: where the caller keeps the prepared (channels-last, split) form of x between calls that share the same input
    // and strides (the phases of a transposed convolution). *x_prepared == nullptr: this call prepares x and leaves the buffer
    // there; otherwise it is used as it is. The caller frees it (axb_free, stream-ordered).
    void **x_prepared;
};
axb_status conv_tc(const ConvDesc &d, DeviceCtx *ctx, cudaStream_t s, bool *done);

// Runs the tensor-core path when the problem qualifies (sets *done), otherwise leaves it to the SIMT kernel.
axb_status try_tcgen05(const Desc &d, int adt, int bdt, int cdt, DeviceCtx *ctx, cudaStream_t s, bool *done);

} // namespace gemm
} // namespace axb
