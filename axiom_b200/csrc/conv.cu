// conv.cu — data-movement halves of the convolution / padding path ("next" row §8f-2: conv1d as GEMM).
//
// The reference computes conv1d with a naive 6-deep loop accumulating in fp64 (src/tensor/operations.cpp:3046-3135); its
// Metal lowering hands the op to MPSGraph (src/backends/metal/mpsgraph_graph_compiler.mm:365-405). Here conv1d is
//   out[n] = W2d [C_out, C_in*K]  x  col[n] [C_in*K, L_out]        (one batched GEMM, tcgen05 for half types)
// where col is written once by axb_im2col_1d below (HBM-bound gather with implicit zero padding; consecutive threads
// write consecutive output positions, reads are unit-stride for stride 1 and stride-s otherwise).
// axb_pad_constant is the N-d constant pad of OpType::Pad (operations.cpp:2150-2260).
#include "axb_internal.h"

namespace axb {
namespace {

struct Im2colParams {
    const char *in;
    char *out;
    int64_t in_sn, in_sc, in_sl; // byte strides of the input (N, C, L)
    int64_t N, C, L, K, Lout;
    int32_t stride, padding, dilation, esize;
};

template <typename T> __global__ void __launch_bounds__(256) im2col1d_kernel(const __grid_constant__ Im2colParams p) {
    // grid: x over L_out tiles, y over (c, k) rows, z over batch
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.y; // c * K + k
    const int64_t n = blockIdx.z;
    if (l >= p.Lout)
        return;
    const int64_t c = row / p.K, k = row - c * p.K;
    const int64_t src_l = l * p.stride - p.padding + k * p.dilation;
    T v = T(0);
    if (src_l >= 0 && src_l < p.L)
        v = *reinterpret_cast<const T *>(p.in + n * p.in_sn + c * p.in_sc + src_l * p.in_sl);
    reinterpret_cast<T *>(p.out)[(n * p.C * p.K + row) * p.Lout + l] = v;
}

struct PadParams {
    const char *in;
    char *out;
    int32_t ndim, esize;
    int64_t out_shape[AXB_MAX_DIMS], in_shape[AXB_MAX_DIMS], lo[AXB_MAX_DIMS], in_stride[AXB_MAX_DIMS];
    int64_t total;
    uint64_t fill_bits;
};

template <typename T> __global__ void __launch_bounds__(256) pad_kernel(const __grid_constant__ PadParams p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = i, off = 0;
        bool inside = true;
#pragma unroll 1
        for (int d = p.ndim - 1; d >= 0; --d) {
            const int64_t q = rem / p.out_shape[d];
            const int64_t c = rem - q * p.out_shape[d] - p.lo[d];
            rem = q;
            inside = inside && c >= 0 && c < p.in_shape[d];
            off += c * p.in_stride[d];
        }
        T v;
        if (inside) {
            v = *reinterpret_cast<const T *>(p.in + off);
        } else {
            const uint64_t b = p.fill_bits;
            memcpy(&v, &b, sizeof(T));
        }
        reinterpret_cast<T *>(p.out)[i] = v;
    }
}

// value -> bit pattern of dtype (floating and integer types)
uint64_t fill_bits_of(double value, int32_t dt) {
    uint64_t bits = 0;
    switch (dt) {
    case AXB_F32: { float f = (float)value; memcpy(&bits, &f, 4); break; }
    case AXB_F64: memcpy(&bits, &value, 8); break;
    case AXB_BF16: {
        float f = (float)value; uint32_t u; memcpy(&u, &f, 4);
        if (f != f) { bits = 0x7FC0u; break; }
        u += 0x7FFFu + ((u >> 16) & 1u); bits = u >> 16; break;
    }
    case AXB_F16: {
        // round-to-nearest-even float -> half through the hardware-equivalent integer path
        float f = (float)value; uint32_t x; memcpy(&x, &f, 4);
        const uint32_t sign = (x >> 16) & 0x8000u; x &= 0x7FFFFFFFu;
        uint32_t h;
        if (x >= 0x7F800000u) h = x > 0x7F800000u ? 0x7E00u : 0x7C00u;
        else if (x >= 0x477FF000u) h = 0x7C00u;
        else if (x < 0x33000001u) h = 0;
        else if (x < 0x38800000u) { // subnormal half
            const int sh = 126 - (int)(x >> 23); uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
            const uint32_t r = m >> (sh + 1), rem = m & ((1u << (sh + 1)) - 1), half = 1u << sh;
            h = r + ((rem > half || (rem == half && (r & 1u))) ? 1u : 0u);
        } else {
            const uint32_t e = (x >> 23) - 112, m = x & 0x7FFFFFu; h = (e << 10) | (m >> 13);
            const uint32_t rem = m & 0x1FFFu; if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
        }
        bits = sign | h; break;
    }
    case AXB_BOOL: bits = value != 0.0 ? 1 : 0; break;
    default: { int64_t i = (int64_t)value; memcpy(&bits, &i, 8); break; }
    }
    return bits;
}

} // namespace
} // namespace axb

using namespace axb;

extern "C" {

axb_status axb_im2col_1d(const axb_tensor *in, int kernel_size, int stride, int padding, int dilation,
                         const axb_tensor *out, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_im2col_1d(in)"));
    AXB_TRY(validate_tensor(out, "axb_im2col_1d(out)"));
    AXB_REQUIRE(in->ndim == 3 && out->ndim == 3, AXB_ERR_SHAPE, "axb_im2col_1d: in must be (N, C, L), out (N, C*K, L_out)");
    AXB_REQUIRE(kernel_size > 0 && stride > 0 && dilation > 0 && padding >= 0, AXB_ERR_INVALID, "axb_im2col_1d: bad geometry");
    AXB_REQUIRE(in->dtype == out->dtype && is_contiguous(*out), AXB_ERR_INVALID, "axb_im2col_1d: out must be contiguous, same dtype");
    Im2colParams p{};
    p.N = in->shape[0]; p.C = in->shape[1]; p.L = in->shape[2]; p.K = kernel_size;
    const int64_t span = (int64_t)dilation * (kernel_size - 1) + 1;
    AXB_REQUIRE(p.L + 2 * (int64_t)padding >= span, AXB_ERR_SHAPE, "axb_im2col_1d: input shorter than the dilated kernel");
    p.Lout = (p.L + 2 * (int64_t)padding - span) / stride + 1;
    AXB_REQUIRE(out->shape[0] == p.N && out->shape[1] == p.C * p.K && out->shape[2] == p.Lout, AXB_ERR_SHAPE,
                "axb_im2col_1d: out shape must be {%lld, %lld, %lld}", (long long)p.N, (long long)(p.C * p.K), (long long)p.Lout);
    if (numel(*out) == 0)
        return AXB_OK;
    AXB_REQUIRE(p.C * p.K <= 65535 && p.N <= 65535, AXB_ERR_UNSUPPORTED, "axb_im2col_1d: C*K and N must be <= 65535");
    p.in = tensor_ptr(*in); p.out = tensor_ptr(*out);
    p.in_sn = in->strides[0]; p.in_sc = in->strides[1]; p.in_sl = in->strides[2];
    p.stride = stride; p.padding = padding; p.dilation = dilation;
    p.esize = (int)dtype_size(in->dtype);
    cudaStream_t s;
    AXB_TRY(resolve_stream(stream, &s));
    dim3 grid((unsigned)((p.Lout + 255) / 256), (unsigned)(p.C * p.K), (unsigned)p.N);
    switch (p.esize) {
    case 1: im2col1d_kernel<uint8_t><<<grid, 256, 0, s>>>(p); break;
    case 2: im2col1d_kernel<uint16_t><<<grid, 256, 0, s>>>(p); break;
    case 4: im2col1d_kernel<uint32_t><<<grid, 256, 0, s>>>(p); break;
    case 8: im2col1d_kernel<uint64_t><<<grid, 256, 0, s>>>(p); break;
    default: return fail(AXB_ERR_UNSUPPORTED, "axb_im2col_1d: element size %d", p.esize);
    }
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

axb_status axb_pad_constant(const axb_tensor *in, const int64_t *pad_lo, const int64_t *pad_hi, double value,
                            const axb_tensor *out, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_pad_constant(in)"));
    AXB_TRY(validate_tensor(out, "axb_pad_constant(out)"));
    AXB_REQUIRE(in->ndim == out->ndim && in->dtype == out->dtype && is_contiguous(*out), AXB_ERR_INVALID,
                "axb_pad_constant: out must be contiguous with the input's rank and dtype");
    AXB_REQUIRE(!is_complex_dtype(in->dtype), AXB_ERR_UNSUPPORTED, "axb_pad_constant: complex dtypes have no GPU path");
    PadParams p{};
    p.ndim = in->ndim;
    p.total = 1;
    for (int d = 0; d < in->ndim; ++d) {
        AXB_REQUIRE(pad_lo[d] >= 0 && pad_hi[d] >= 0 && out->shape[d] == in->shape[d] + pad_lo[d] + pad_hi[d], AXB_ERR_SHAPE,
                    "axb_pad_constant: out extent %d must be in + lo + hi", d);
        p.out_shape[d] = out->shape[d]; p.in_shape[d] = in->shape[d]; p.lo[d] = pad_lo[d]; p.in_stride[d] = in->strides[d];
        p.total *= out->shape[d];
    }
    if (p.total == 0)
        return AXB_OK;
    p.in = tensor_ptr(*in); p.out = tensor_ptr(*out);
    p.esize = (int)dtype_size(in->dtype);
    p.fill_bits = fill_bits_of(value, in->dtype);
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const unsigned grid = (unsigned)std::min<int64_t>((p.total + 255) / 256, (int64_t)ctx->sm_count * 16);
    switch (p.esize) {
    case 1: pad_kernel<uint8_t><<<grid, 256, 0, s>>>(p); break;
    case 2: pad_kernel<uint16_t><<<grid, 256, 0, s>>>(p); break;
    case 4: pad_kernel<uint32_t><<<grid, 256, 0, s>>>(p); break;
    case 8: pad_kernel<uint64_t><<<grid, 256, 0, s>>>(p); break;
    default: return fail(AXB_ERR_UNSUPPORTED, "axb_pad_constant: element size %d", p.esize);
    }
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

} // extern "C"
