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

} // extern "C"
