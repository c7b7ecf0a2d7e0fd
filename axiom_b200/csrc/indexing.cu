// indexing.cu — gather / index_select (CPUGatherOperation src/backends/cpu/cpu_operations.cpp:2796-2880,
// CPUIndexSelectOperation :2964-3060; used by ops::embedding src/tensor/operations.cpp:1751-1772 and the parity
// tests tests/test_cpu_gpu_parity.cpp:422-462). Pure data movement: one thread per output element, the element is
// moved as an opaque 1/2/4/8/16-byte word so every dtype (complex included) is covered. Negative indices wrap.
#include <cstring>

#include "axb_internal.h"

namespace axb {
namespace {

struct IdxParams {
    const char *in;
    const char *idx;
    char *out;
    int32_t ndim;      // dims of the OUTPUT iteration space (innermost first)
    int32_t dim;       // position (innermost-first) of the indexed dim
    int32_t esize;     // bytes per element
    int32_t idx_dtype; // AXB_I32 / AXB_I64
    int32_t mode;      // 0 = gather (index tensor has the output's shape), 1 = index_select (1-D index along dim)
    uint32_t oshape[AXB_MAX_DIMS];
    int64_t in_stride[AXB_MAX_DIMS];  // bytes, per output dim
    int64_t idx_stride[AXB_MAX_DIMS]; // bytes, per output dim (gather) or [dim] only (index_select)
    int64_t dim_size;
    int64_t total;
};

__device__ __forceinline__ void copy_elem(char *dst, const char *src, int esize) {
    switch (esize) {
    case 1: *dst = *src; break;
    case 2: *reinterpret_cast<uint16_t *>(dst) = *reinterpret_cast<const uint16_t *>(src); break;
    case 4: *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src); break;
    case 8: *reinterpret_cast<uint64_t *>(dst) = *reinterpret_cast<const uint64_t *>(src); break;
    default: *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src); break;
    }
}

__global__ void __launch_bounds__(256) index_kernel(const __grid_constant__ IdxParams p) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < p.total; o += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = o, in_off = 0, idx_off = 0;
#pragma unroll 1
        for (int d = 0; d < p.ndim; ++d) {
            int64_t q = rem / p.oshape[d];
            int64_t c = rem - q * p.oshape[d];
            rem = q;
            if (d == p.dim) {
                idx_off += c * p.idx_stride[d];
            } else {
                in_off += c * p.in_stride[d];
                if (p.mode == 0)
                    idx_off += c * p.idx_stride[d];
            }
        }
        int64_t i = p.idx_dtype == AXB_I64 ? *reinterpret_cast<const int64_t *>(p.idx + idx_off)
                                           : (int64_t) * reinterpret_cast<const int32_t *>(p.idx + idx_off);
        if (i < 0)
            i += p.dim_size;
        i = i < 0 ? 0 : (i >= p.dim_size ? p.dim_size - 1 : i); // out-of-range is UB in the reference; stay in bounds
        in_off += i * p.in_stride[p.dim];
        copy_elem(p.out + o * p.esize, p.in + in_off, p.esize);
    }
}

axb_status run_index(int mode, const axb_tensor *in, int dim, const axb_tensor *indices, const axb_tensor *out,
                     axb_stream_t stream, const char *what) {
    AXB_TRY(validate_tensor(in, what));
    AXB_TRY(validate_tensor(indices, what));
    AXB_TRY(validate_tensor(out, what));
    AXB_REQUIRE(in->dtype == out->dtype, AXB_ERR_INVALID, "%s: out dtype must equal the input dtype", what);
    AXB_REQUIRE(indices->dtype == AXB_I64 || indices->dtype == AXB_I32, AXB_ERR_UNSUPPORTED,
                "%s: indices must be Int64 or Int32", what);
    AXB_REQUIRE(is_contiguous(*out), AXB_ERR_UNSUPPORTED, "%s: out must be contiguous", what);
    if (dim < 0)
        dim += in->ndim;
    AXB_REQUIRE(dim >= 0 && dim < in->ndim, AXB_ERR_INVALID, "%s: dim out of range", what);
    AXB_REQUIRE(out->ndim == in->ndim, AXB_ERR_SHAPE, "%s: out rank must equal the input rank", what);
    IdxParams p;
    memset(&p, 0, sizeof p);
    if (mode == 0) {
        AXB_REQUIRE(indices->ndim == in->ndim, AXB_ERR_SHAPE, "%s: indices rank must equal the input rank", what);
        for (int d = 0; d < in->ndim; ++d)
            AXB_REQUIRE(out->shape[d] == indices->shape[d], AXB_ERR_SHAPE, "%s: out shape must equal the indices shape", what);
    } else {
        AXB_REQUIRE(numel(*indices) == out->shape[dim], AXB_ERR_SHAPE, "%s: out dim %d must equal the number of indices", what, dim);
        for (int d = 0; d < in->ndim; ++d)
            AXB_REQUIRE(d == dim || out->shape[d] == in->shape[d], AXB_ERR_SHAPE, "%s: out shape differs from the input in dim %d", what, d);
    }
    p.total = numel(*out);
    if (p.total == 0)
        return AXB_OK;
    AXB_REQUIRE(in->shape[dim] > 0, AXB_ERR_INVALID, "%s: indexed dim is empty", what);
    p.in = tensor_ptr(*in);
    p.idx = tensor_ptr(*indices);
    p.out = tensor_ptr(*out);
    p.ndim = in->ndim;
    p.dim = in->ndim - 1 - dim;
    p.esize = (int)dtype_size(in->dtype);
    p.idx_dtype = indices->dtype;
    p.mode = mode;
    p.dim_size = in->shape[dim];
    int64_t flat_idx_stride = (int64_t)dtype_size(indices->dtype);
    if (mode == 1) { // any-rank index tensor is read in C order; only contiguous or 1-D strided supported
        if (indices->ndim == 1)
            flat_idx_stride = indices->strides[0];
        else
            AXB_REQUIRE(is_contiguous(*indices), AXB_ERR_UNSUPPORTED, "%s: multi-dim indices must be contiguous", what);
    }
    for (int d = 0; d < in->ndim; ++d) {
        int s = in->ndim - 1 - d; // source dim
        p.oshape[d] = (uint32_t)out->shape[s];
        p.in_stride[d] = in->strides[s];
        p.idx_stride[d] = mode == 0 ? indices->strides[s] : (s == dim ? flat_idx_stride : 0);
    }
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    int64_t blocks = std::min<int64_t>((p.total + 255) / 256, (int64_t)ctx->sm_count * 16);
    index_kernel<<<(unsigned)blocks, 256, 0, s>>>(p);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

// scatter (CPUScatterOperation src/backends/cpu/cpu_operations.cpp:2886-2960): out = copy(in);
// out[.., idx[c], ..] = src[c] for every coordinate c of the index tensor. Duplicate targets: one of the writers wins
// (the CPU loop keeps the last in C order; results only differ when the caller scatters two values to one slot).
struct ScatterParams {
    const char *idx, *src;
    char *out;
    int32_t ndim, dim, esize, idx_dtype;
    uint32_t ishape[AXB_MAX_DIMS];                                   // index tensor shape, innermost first
    int64_t idx_stride[AXB_MAX_DIMS], src_stride[AXB_MAX_DIMS], out_stride[AXB_MAX_DIMS];
    int64_t dim_size, total;
};

__global__ void __launch_bounds__(256) scatter_kernel(const __grid_constant__ ScatterParams p) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < p.total; o += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = o, idx_off = 0, src_off = 0, out_off = 0;
#pragma unroll 1
        for (int d = 0; d < p.ndim; ++d) {
            const int64_t q = rem / p.ishape[d];
            const int64_t c = rem - q * p.ishape[d];
            rem = q;
            idx_off += c * p.idx_stride[d];
            src_off += c * p.src_stride[d];
            if (d != p.dim)
                out_off += c * p.out_stride[d];
        }
        int64_t i = p.idx_dtype == AXB_I64 ? *reinterpret_cast<const int64_t *>(p.idx + idx_off)
                                           : (int64_t) * reinterpret_cast<const int32_t *>(p.idx + idx_off);
        if (i < 0)
            i += p.dim_size;
        if (i < 0 || i >= p.dim_size)
            continue; // out-of-range targets are dropped instead of written out of bounds
        copy_elem(p.out + out_off + i * p.out_stride[p.dim], p.src + src_off, p.esize);
    }
}

} // namespace
} // namespace axb

using namespace axb;

extern "C" {
axb_status axb_gather(const axb_tensor *in, int dim, const axb_tensor *indices, const axb_tensor *out,
                      axb_stream_t stream) {
    return run_index(0, in, dim, indices, out, stream, "axb_gather");
}
axb_status axb_index_select(const axb_tensor *in, int dim, const axb_tensor *indices, const axb_tensor *out,
                            axb_stream_t stream) {
    return run_index(1, in, dim, indices, out, stream, "axb_index_select");
}

axb_status axb_scatter(const axb_tensor *in, int dim, const axb_tensor *indices, const axb_tensor *src, const axb_tensor *out,
                       axb_stream_t stream) {
    const char *what = "axb_scatter";
    AXB_TRY(validate_tensor(in, what));
    AXB_TRY(validate_tensor(indices, what));
    AXB_TRY(validate_tensor(src, what));
    AXB_TRY(validate_tensor(out, what));
    AXB_REQUIRE(in->dtype == out->dtype && src->dtype == in->dtype, AXB_ERR_INVALID, "%s: in, src and out must share one dtype", what);
    AXB_REQUIRE(indices->dtype == AXB_I64 || indices->dtype == AXB_I32, AXB_ERR_UNSUPPORTED, "%s: indices must be Int64 or Int32", what);
    if (dim < 0)
        dim += in->ndim;
    AXB_REQUIRE(dim >= 0 && dim < in->ndim, AXB_ERR_INVALID, "%s: dim out of range", what);
    AXB_REQUIRE(indices->ndim == in->ndim && src->ndim == in->ndim && out->ndim == in->ndim, AXB_ERR_SHAPE,
                "%s: in, indices, src and out must have the same rank", what);
    for (int d = 0; d < in->ndim; ++d) {
        AXB_REQUIRE(out->shape[d] == in->shape[d], AXB_ERR_SHAPE, "%s: out shape must equal the input shape", what);
        AXB_REQUIRE(indices->shape[d] <= src->shape[d] && (d == dim || indices->shape[d] <= in->shape[d]), AXB_ERR_SHAPE,
                    "%s: index extent %d exceeds src / input", what, d);
    }
    AXB_TRY(axb_gather_strided(in, out, stream)); // out = copy(in)
    ScatterParams p;
    memset(&p, 0, sizeof p);
    p.total = numel(*indices);
    if (p.total == 0 || numel(*out) == 0)
        return AXB_OK;
    p.idx = tensor_ptr(*indices);
    p.src = tensor_ptr(*src);
    p.out = tensor_ptr(*out);
    p.ndim = in->ndim;
    p.dim = in->ndim - 1 - dim;
    p.esize = (int)dtype_size(in->dtype);
    p.idx_dtype = indices->dtype;
    p.dim_size = in->shape[dim];
    for (int d = 0; d < in->ndim; ++d) {
        const int s = in->ndim - 1 - d;
        p.ishape[d] = (uint32_t)indices->shape[s];
        p.idx_stride[d] = indices->strides[s];
        p.src_stride[d] = src->strides[s];
        p.out_stride[d] = out->strides[s];
    }
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const int64_t blocks = std::min<int64_t>((p.total + 255) / 256, (int64_t)ctx->sm_count * 16);
    scatter_kernel<<<(unsigned)blocks, 256, 0, s>>>(p);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}
}
