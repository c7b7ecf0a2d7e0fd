// host_api.cpp — host-buffer entry points: what `x.gpu()` ... `.cpu()` does end to end for one op group
// (Tensor::to src/tensor/tensor.cpp:1593-1639 + materialize + Tensor::cpu :1642-1650): stage the inputs
// into device blocks from the caching allocator, launch, copy the result back, synchronise once.
#include <algorithm>
#include <cstring>
#include <vector>

#include "axb_internal.h"

using namespace axb;

namespace {

// byte span [lo, hi) touched by a strided host view, relative to base+byte_offset
void view_span(const axb_tensor &t, int64_t *lo, int64_t *hi) {
    int64_t mn = 0, mx = 0;
    bool empty = false;
    for (int d = 0; d < t.ndim; ++d) {
        if (t.shape[d] == 0)
            empty = true;
        int64_t ext = (t.shape[d] - 1) * t.strides[d];
        if (ext < 0)
            mn += ext;
        else
            mx += ext;
    }
    *lo = mn;
    *hi = empty ? mn : mx + (int64_t)dtype_size(t.dtype);
}

struct Staged {
    void *dev = nullptr;
    axb_tensor desc{};
};

// copies the span of a host view to the device and returns the equivalent device view
axb_status stage_in(const axb_tensor &h, cudaStream_t s, Staged *out) {
    int64_t lo, hi;
    view_span(h, &lo, &hi);
    size_t bytes = (size_t)std::max<int64_t>(hi - lo, 0);
    AXB_TRY(axb_alloc(bytes, s, &out->dev));
    if (bytes)
        AXB_CUDA(cudaMemcpyAsync(out->dev, tensor_ptr(h) + lo, bytes, cudaMemcpyHostToDevice, s));
    out->desc = h;
    out->desc.base = out->dev;
    out->desc.byte_offset = -lo;
    return AXB_OK;
}

axb_status alloc_out(const axb_tensor &h, cudaStream_t s, Staged *out) {
    AXB_REQUIRE(is_contiguous(h), AXB_ERR_UNSUPPORTED, "host output tensor must be contiguous");
    size_t bytes = (size_t)numel(h) * dtype_size(h.dtype);
    AXB_TRY(axb_alloc(bytes, s, &out->dev));
    out->desc = h;
    out->desc.base = out->dev;
    out->desc.byte_offset = 0;
    return AXB_OK;
}

axb_status copy_back_and_release(const axb_tensor &h, Staged &o, std::vector<Staged> &ins, cudaStream_t s,
                                 axb_status run_status) {
    axb_status st = run_status;
    if (st == AXB_OK) {
        size_t bytes = (size_t)numel(h) * dtype_size(h.dtype);
        if (bytes) {
            cudaError_t e = cudaMemcpyAsync(tensor_ptr(h), o.dev, bytes, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess)
                st = fail(AXB_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
        }
    }
    for (auto &i : ins)
        axb_free(i.dev, s);
    axb_free(o.dev, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (st == AXB_OK && e != cudaSuccess)
        st = fail(AXB_ERR_CUDA, "stream synchronize failed: %s", cudaGetErrorString(e));
    return st;
}

} // namespace

extern "C" {

axb_status axb_elementwise_host(const axb_instr *prog, int n_instr, const axb_tensor *host_inputs, int n_inputs,
                                const axb_tensor *host_out) {
    AXB_REQUIRE(n_inputs >= 0 && n_inputs <= AXB_MAX_INPUTS, AXB_ERR_INVALID, "axb_elementwise_host: bad input count");
    AXB_TRY(validate_tensor(host_out, "axb_elementwise_host(out)"));
    cudaStream_t s;
    AXB_TRY(resolve_stream(nullptr, &s));
    std::vector<Staged> ins(n_inputs);
    std::vector<axb_tensor> descs(std::max(n_inputs, 1));
    axb_status st = AXB_OK;
    for (int k = 0; k < n_inputs && st == AXB_OK; ++k) {
        st = validate_tensor(&host_inputs[k], "axb_elementwise_host(input)");
        if (st == AXB_OK)
            st = stage_in(host_inputs[k], s, &ins[k]);
        descs[k] = ins[k].desc;
    }
    Staged o;
    if (st == AXB_OK)
        st = alloc_out(*host_out, s, &o);
    if (st == AXB_OK)
        st = axb_elementwise(prog, n_instr, descs.data(), n_inputs, &o.desc, s);
    return copy_back_and_release(*host_out, o, ins, s, st);
}

axb_status axb_reduce_host(int32_t op, const axb_tensor *host_in, const int32_t *axes, int n_axes, int keep_dims,
                           const axb_tensor *host_out) {
    AXB_TRY(validate_tensor(host_in, "axb_reduce_host(in)"));
    AXB_TRY(validate_tensor(host_out, "axb_reduce_host(out)"));
    cudaStream_t s;
    AXB_TRY(resolve_stream(nullptr, &s));
    std::vector<Staged> ins(1);
    axb_status st = stage_in(*host_in, s, &ins[0]);
    Staged o;
    if (st == AXB_OK)
        st = alloc_out(*host_out, s, &o);
    if (st == AXB_OK)
        st = axb_reduce(op, &ins[0].desc, axes, n_axes, keep_dims, &o.desc, s);
    return copy_back_and_release(*host_out, o, ins, s, st);
}

axb_status axb_matmul_host(const axb_tensor *host_a, const axb_tensor *host_b, int transpose_a, int transpose_b,
                           const axb_tensor *host_out) {
    AXB_TRY(validate_tensor(host_a, "axb_matmul_host(a)"));
    AXB_TRY(validate_tensor(host_b, "axb_matmul_host(b)"));
    AXB_TRY(validate_tensor(host_out, "axb_matmul_host(out)"));
    cudaStream_t s;
    AXB_TRY(resolve_stream(nullptr, &s));
    std::vector<Staged> ins(2);
    axb_status st = stage_in(*host_a, s, &ins[0]);
    if (st == AXB_OK)
        st = stage_in(*host_b, s, &ins[1]);
    Staged o;
    if (st == AXB_OK)
        st = alloc_out(*host_out, s, &o);
    if (st == AXB_OK)
        st = axb_matmul(&ins[0].desc, &ins[1].desc, transpose_a, transpose_b, nullptr, &o.desc, s);
    return copy_back_and_release(*host_out, o, ins, s, st);
}

// FusedReductionStep (include/axiom/graph/compiled_graph.hpp:61-66): elementwise chain + FULL reduction.
// Round 1: the chain's value is materialised once in a pooled scratch block and reduced by the row kernel (2 launches +
// finish); folding the reduction into the chain kernel's epilogue is the next step for this entry point.
axb_status axb_elementwise_reduce(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                  int32_t reduce_op, const axb_tensor *out, axb_stream_t stream) {
    AXB_REQUIRE(prog && n_instr > 0 && n_inputs > 0 && inputs, AXB_ERR_INVALID, "axb_elementwise_reduce: bad arguments");
    AXB_TRY(validate_tensor(out, "axb_elementwise_reduce(out)"));
    int64_t shape[AXB_MAX_DIMS];
    int nd = 0;
    {
        int64_t cur[AXB_MAX_DIMS];
        int ncur = inputs[0].ndim;
        memcpy(cur, inputs[0].shape, sizeof(int64_t) * ncur);
        for (int k = 1; k < n_inputs; ++k) {
            AXB_TRY(axb_broadcast_shape(cur, ncur, inputs[k].shape, inputs[k].ndim, shape, &nd));
            memcpy(cur, shape, sizeof(int64_t) * nd);
            ncur = nd;
        }
        memcpy(shape, cur, sizeof(int64_t) * ncur);
        nd = ncur;
    }
    cudaStream_t s;
    AXB_TRY(resolve_stream(stream, &s));
    const int32_t vdt = prog[n_instr - 1].dtype;
    axb_tensor tmp = make_contiguous_desc(nullptr, vdt, shape, nd);
    void *buf = nullptr;
    AXB_TRY(axb_alloc((size_t)numel(tmp) * dtype_size(vdt), s, &buf));
    tmp.base = buf;
    axb_status st = axb_elementwise(prog, n_instr, inputs, n_inputs, &tmp, s);
    if (st == AXB_OK)
        st = axb_reduce(reduce_op, &tmp, nullptr, 0, 0, out, s);
    axb_free(buf, s);
    return st;
}

} // extern "C"
