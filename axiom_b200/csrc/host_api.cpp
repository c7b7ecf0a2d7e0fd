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
    {
        AXB_CUDA(cudaMemcpyAsync(out->dev, tensor_ptr(h) + lo, bytes, cudaMemcpyHostToDevice, s));
        count_h2d(bytes);
    }
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
            count_d2h(bytes);
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

// ---- pipelined variant --------------------------------------------------------------------------------------------
// A large elementwise group is PCIe-bound end to end (config 2: 268 MB up, 268 MB down, 0.09 ms of kernel). PCIe is
// full duplex, so the rows of the result are cut into chunks: chunk i+1 uploads on one stream while chunk i computes on
// the backend stream and chunk i-1 downloads on a third — uploads and downloads then overlap and the step costs
// max(up, down) instead of up + down. Operands that do not have the result's leading extent (broadcast rows, scalars,
// small vectors) are staged whole once.
struct SideStreams {
    cudaStream_t in = nullptr, out = nullptr;
};
SideStreams *side_streams(int device) {
    static SideStreams per_dev[64];
    if (device < 0 || device >= 64)
        return nullptr;
    SideStreams &ss = per_dev[device];
    if (!ss.in) {
        if (cudaStreamCreateWithFlags(&ss.in, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ss.out, cudaStreamNonBlocking) != cudaSuccess)
            return nullptr;
    }
    return &ss;
}

constexpr int kMaxChunks = 16;
constexpr int64_t kPipelineMinBytes = 32ll << 20, kSmallOperandBytes = 1ll << 20;

// true when operand `h` can be cut along the result's leading dim into non-overlapping host spans
bool row_sliceable(const axb_tensor &h, const axb_tensor &out) {
    if (h.ndim != out.ndim || h.ndim < 1 || h.shape[0] != out.shape[0])
        return false;
    int64_t inner_hi = (int64_t)dtype_size(h.dtype);
    for (int d = 0; d < h.ndim; ++d) {
        if (h.strides[d] < 0)
            return false;
        if (d > 0)
            inner_hi += (h.shape[d] - 1) * h.strides[d];
    }
    return h.shape[0] <= 1 || h.strides[0] >= inner_hi; // rows do not interleave
}

// returns AXB_OK and sets *handled when the pipelined path ran
axb_status pipelined_elementwise(const axb_instr *prog, int n_instr, const axb_tensor *hin, int n_in, const axb_tensor *hout,
                                 cudaStream_t s, DeviceCtx *ctx, bool *handled) {
    *handled = false;
    const int64_t out_bytes = numel(*hout) * (int64_t)dtype_size(hout->dtype);
    if (!is_contiguous(*hout) || hout->ndim < 2 || out_bytes < kPipelineMinBytes || hout->shape[0] < 2 * kMaxChunks)
        return AXB_OK;
    bool sliced[AXB_MAX_INPUTS];
    for (int k = 0; k < n_in; ++k) {
        int64_t lo, hi;
        view_span(hin[k], &lo, &hi);
        sliced[k] = (hi - lo) > kSmallOperandBytes;
        if (sliced[k] && !row_sliceable(hin[k], *hout))
            return AXB_OK; // a large operand that is broadcast along rows or oddly strided: plain path
    }
    SideStreams *ss = side_streams(ctx->device);
    if (!ss)
        return AXB_OK;
    *handled = true;

    std::vector<Staged> ins(n_in);
    Staged o;
    std::vector<cudaEvent_t> events;
    auto new_event = [&](cudaEvent_t *e) {
        cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
        if (r == cudaSuccess)
            events.push_back(*e);
        return r;
    };
    axb_status st = AXB_OK;
    auto cuda_ok = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && st == AXB_OK)
            st = fail(AXB_ERR_CUDA, "axb_elementwise_host (pipelined): %s failed: %s", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    // device blocks for everything (full size), small operands uploaded right away on the backend stream
    for (int k = 0; k < n_in && st == AXB_OK; ++k) {
        if (!sliced[k]) {
            st = stage_in(hin[k], s, &ins[k]);
        } else {
            int64_t lo, hi;
            view_span(hin[k], &lo, &hi);
            st = axb_alloc((size_t)(hi - lo), s, &ins[k].dev);
            ins[k].desc = hin[k];
            ins[k].desc.base = ins[k].dev;
            ins[k].desc.byte_offset = -lo; // lo == 0 here (non-negative strides)
        }
    }
    if (st == AXB_OK)
        st = alloc_out(*hout, s, &o);
    cudaEvent_t ready = nullptr;
    if (st == AXB_OK && cuda_ok(new_event(&ready), "cudaEventCreate")) {
        cuda_ok(cudaEventRecord(ready, s), "cudaEventRecord");
        cuda_ok(cudaStreamWaitEvent(ss->in, ready, 0), "cudaStreamWaitEvent");  // allocator blocks are ordered on s
        cuda_ok(cudaStreamWaitEvent(ss->out, ready, 0), "cudaStreamWaitEvent");
    }
    const int64_t rows = hout->shape[0];
    const int chunks = (int)std::min<int64_t>(kMaxChunks, rows);
    const int64_t out_row_bytes = out_bytes / rows;
    for (int c = 0; c < chunks && st == AXB_OK; ++c) {
        const int64_t r0 = rows * c / chunks, r1 = rows * (c + 1) / chunks;
        std::vector<axb_tensor> d(std::max(n_in, 1));
        for (int k = 0; k < n_in; ++k) {
            d[k] = ins[k].desc;
            // every operand that has the result's leading extent is VIEWED per chunk (a {rows, 1} column as well as the
            // big {rows, cols} operand); operands with fewer dims or extent 1 broadcast along rows and stay whole
            if (hin[k].ndim == hout->ndim && hin[k].shape[0] == rows && rows > 1) {
                d[k].shape[0] = r1 - r0;
                d[k].byte_offset += r0 * hin[k].strides[0];
            }
            if (!sliced[k])
                continue;
            axb_tensor hv = hin[k];
            hv.shape[0] = r1 - r0;
            int64_t lo, hi;
            view_span(hv, &lo, &hi);
            cuda_ok(cudaMemcpyAsync(static_cast<char *>(ins[k].dev) + r0 * hin[k].strides[0], tensor_ptr(hin[k]) + r0 * hin[k].strides[0],
                                    (size_t)(hi - lo), cudaMemcpyHostToDevice, ss->in),
                    "H2D");
            count_h2d((size_t)(hi - lo));
        }
        cudaEvent_t up = nullptr, done = nullptr;
        if (!cuda_ok(new_event(&up), "cudaEventCreate") || !cuda_ok(new_event(&done), "cudaEventCreate"))
            break;
        cuda_ok(cudaEventRecord(up, ss->in), "cudaEventRecord");
        cuda_ok(cudaStreamWaitEvent(s, up, 0), "cudaStreamWaitEvent");
        axb_tensor od = o.desc;
        od.shape[0] = r1 - r0;
        od.byte_offset = r0 * out_row_bytes;
        if (st == AXB_OK)
            st = axb_elementwise(prog, n_instr, d.data(), n_in, &od, s);
        cuda_ok(cudaEventRecord(done, s), "cudaEventRecord");
        cuda_ok(cudaStreamWaitEvent(ss->out, done, 0), "cudaStreamWaitEvent");
        if (st == AXB_OK) {
            cuda_ok(cudaMemcpyAsync(tensor_ptr(*hout) + r0 * out_row_bytes, static_cast<char *>(o.dev) + r0 * out_row_bytes,
                                    (size_t)((r1 - r0) * out_row_bytes), cudaMemcpyDeviceToHost, ss->out),
                    "D2H");
            count_d2h((size_t)((r1 - r0) * out_row_bytes));
        }
    }
    // the blocks go back to the pool in backend-stream order: that stream first waits for both side streams
    cudaEvent_t in_done = nullptr, out_done = nullptr;
    if (new_event(&in_done) == cudaSuccess && new_event(&out_done) == cudaSuccess) {
        cudaEventRecord(in_done, ss->in);
        cudaEventRecord(out_done, ss->out);
        cudaStreamWaitEvent(s, in_done, 0);
        cudaStreamWaitEvent(s, out_done, 0);
    } else {
        cudaStreamSynchronize(ss->in);
        cudaStreamSynchronize(ss->out);
    }
    for (auto &i : ins)
        if (i.dev)
            axb_free(i.dev, s);
    if (o.dev)
        axb_free(o.dev, s);
    cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    for (cudaEvent_t e : events)
        cudaEventDestroy(e);
    return st;
}

} // namespace

extern "C" {

axb_status axb_elementwise_host(const axb_instr *prog, int n_instr, const axb_tensor *host_inputs, int n_inputs,
                                const axb_tensor *host_out) {
    AXB_REQUIRE(n_inputs >= 0 && n_inputs <= AXB_MAX_INPUTS, AXB_ERR_INVALID, "axb_elementwise_host: bad input count");
    AXB_TRY(validate_tensor(host_out, "axb_elementwise_host(out)"));
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(nullptr, &s, &ctx));
    for (int k = 0; k < n_inputs; ++k)
        AXB_TRY(validate_tensor(&host_inputs[k], "axb_elementwise_host(input)"));
    bool handled = false;
    axb_status pst = pipelined_elementwise(prog, n_instr, host_inputs, n_inputs, host_out, s, ctx, &handled);
    if (handled)
        return pst;
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
// fp32 chains under sum / mean / max / min run in one pass over the operands (below); every other combination
// materialises the chain's value once in a pooled scratch block and reduces it (2 launches + finish).
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
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const int32_t vdt = prog[n_instr - 1].dtype;
    axb_tensor tmp = make_contiguous_desc(nullptr, vdt, shape, nd);
    // One pass: fp32 chains under sum / mean / max / min fold every CTA's values into ONE partial inside the chain kernel
    // (EW_RED, ew_prelude.cuh) and the reduction family combines the partials — the chain's value never reaches HBM
    // (operands read once, 4 bytes written per CTA, instead of + one write and one read of the whole intermediate).
    static const bool two_pass = getenv("AXB_EW_REDUCE_TWO_PASS") != nullptr; // A/B switch
    bool wide = false;
    for (int k = 0; k < n_inputs; ++k)
        wide = wide || is_64bit_dtype(inputs[k].dtype);
    for (int i = 0; i < n_instr; ++i)
        wide = wide || is_64bit_dtype(prog[i].dtype);
    const int opk = (reduce_op == AXB_OP_SUM || reduce_op == AXB_OP_MEAN) ? 0 : reduce_op == AXB_OP_MAX ? 1 : reduce_op == AXB_OP_MIN ? 2 : -1;
    if (!two_pass && !wide && opk >= 0 && vdt == AXB_F32 && out->dtype == AXB_F32 && numel(*out) == 1 && numel(tmp) > 0) {
        void *partials = nullptr;
        int64_t n_partials = 0;
        AXB_TRY(ew::launch_chain_reduce(prog, n_instr, inputs, n_inputs, &tmp, opk, s, ctx, &partials, &n_partials));
        const axb_tensor pd = make_contiguous_desc(partials, AXB_F32, &n_partials, 1);
        if (reduce_op == AXB_OP_MEAN)
            reduce_count_override() = numel(tmp);
        const axb_status st = axb_reduce(reduce_op, &pd, nullptr, 0, 0, out, s);
        reduce_count_override() = 0;
        axb_free(partials, s);
        return st;
    }
    if (!two_pass && !wide && opk >= 0 && is_half_dtype(vdt) && numel(*out) == 1 && numel(tmp) > 0 &&
        out->dtype == axb_reduce_result_dtype(reduce_op, vdt)) {
        // half-type chains: same one-pass kernel (every value is already rounded to the half type, the fold is fp32 as in
        // the reference's reduction, cpu_operations.cpp:1422-1424), then the reference's result conversion: one rounding to
        // the half type (Mean: rounded to the half type, then widened to Float32 — graph_executor.cpp:396-403)
        void *partials = nullptr, *scratch = nullptr;
        int64_t n_partials = 0, one = 1;
        AXB_TRY(ew::launch_chain_reduce(prog, n_instr, inputs, n_inputs, &tmp, opk, s, ctx, &partials, &n_partials));
        axb_status st = axb_alloc(512, s, &scratch);
        if (st == AXB_OK) {
            const axb_tensor pd = make_contiguous_desc(partials, AXB_F32, &n_partials, 1);
            const axb_tensor f32d = make_contiguous_desc(scratch, AXB_F32, &one, 1);
            const axb_tensor halfd = make_contiguous_desc(static_cast<char *>(scratch) + 256, vdt, &one, 1);
            if (reduce_op == AXB_OP_MEAN)
                reduce_count_override() = numel(tmp);
            st = axb_reduce(reduce_op, &pd, nullptr, 0, 0, &f32d, s);
            reduce_count_override() = 0;
            if (st == AXB_OK && out->dtype == vdt)
                st = axb_cast(&f32d, out, s);
            else if (st == AXB_OK) {
                st = axb_cast(&f32d, &halfd, s);
                if (st == AXB_OK)
                    st = axb_cast(&halfd, out, s);
            }
            axb_free(scratch, s);
        }
        axb_free(partials, s);
        return st;
    }
    void *buf = nullptr;
    AXB_TRY(axb_alloc((size_t)numel(tmp) * dtype_size(vdt), s, &buf));
    tmp.base = buf;
    axb_status st = axb_elementwise(prog, n_instr, inputs, n_inputs, &tmp, s);
    if (st == AXB_OK)
        st = axb_reduce(reduce_op, &tmp, nullptr, 0, 0, out, s);
    axb_free(buf, s);
    return st;
}

// Chain + reduction over ONE axis. The fused single-pass kernel covers fp32 chains reduced over their LAST axis by sum / mean /
// max / min into a contiguous fp32 result; every other case runs the chain and the reduction as two launches.
axb_status axb_elementwise_reduce_axis(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t reduce_op,
                                       int axis, int keep_dims, const axb_tensor *out, axb_stream_t stream) {
    AXB_REQUIRE(prog && n_instr > 0 && n_inputs > 0 && inputs, AXB_ERR_INVALID, "axb_elementwise_reduce_axis: bad arguments");
    AXB_TRY(validate_tensor(out, "axb_elementwise_reduce_axis(out)"));
    int64_t shape[AXB_MAX_DIMS];
    int nd = inputs[0].ndim;
    memcpy(shape, inputs[0].shape, sizeof(int64_t) * nd);
    for (int k = 1; k < n_inputs; ++k) {
        int64_t nxt[AXB_MAX_DIMS];
        int nn = 0;
        AXB_TRY(axb_broadcast_shape(shape, nd, inputs[k].shape, inputs[k].ndim, nxt, &nn));
        memcpy(shape, nxt, sizeof(int64_t) * nn);
        nd = nn;
    }
    if (axis < 0)
        axis += nd;
    AXB_REQUIRE(axis >= 0 && axis < nd, AXB_ERR_INVALID, "axb_elementwise_reduce_axis: axis out of range");
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const int32_t vdt = prog[n_instr - 1].dtype;
    axb_tensor tmp = make_contiguous_desc(nullptr, vdt, shape, nd);
    const int opk = (reduce_op == AXB_OP_SUM || reduce_op == AXB_OP_MEAN) ? 0 : reduce_op == AXB_OP_MAX ? 1 : reduce_op == AXB_OP_MIN ? 2 : -1;
    static const bool two_pass = getenv("AXB_EW_REDUCE_TWO_PASS") != nullptr; // A/B switch
    bool wide = false;
    for (int k = 0; k < n_inputs; ++k)
        wide = wide || is_64bit_dtype(inputs[k].dtype);
    for (int i = 0; i < n_instr; ++i)
        wide = wide || is_64bit_dtype(prog[i].dtype);
    int64_t rows = 1;
    for (int d = 0; d < nd - 1; ++d)
        rows *= shape[d];
    if (!two_pass && !wide && opk >= 0 && axis == nd - 1 && vdt == AXB_F32 && out->dtype == AXB_F32 && is_contiguous(*out) &&
        numel(*out) == rows && numel(tmp) > 0) {
        bool done = false;
        AXB_TRY(ew::launch_chain_rowreduce(prog, n_instr, inputs, n_inputs, &tmp, opk, reduce_op == AXB_OP_MEAN,
                                           reinterpret_cast<float *>(tensor_ptr(*out)), s, ctx, &done));
        if (done)
            return AXB_OK;
    }
    if (numel(tmp) == 0) {
        const int32_t ax = axis;
        tmp.base = reinterpret_cast<void *>(uintptr_t(256));
        return axb_reduce(reduce_op, &tmp, &ax, 1, keep_dims, out, s);
    }
    void *buf = nullptr;
    AXB_TRY(axb_alloc((size_t)numel(tmp) * dtype_size(vdt), s, &buf));
    tmp.base = buf;
    axb_status st = axb_elementwise(prog, n_instr, inputs, n_inputs, &tmp, s);
    const int32_t ax = axis;
    if (st == AXB_OK)
        st = axb_reduce(reduce_op, &tmp, &ax, 1, keep_dims, out, s);
    axb_free(buf, s);
    return st;
}

// softmax / log_softmax over the LAST axis of a chain's value, optionally followed by a sum / mean / max over that same axis.
// One kernel for fp32 chains with rows of at most 8192 elements (the chain's values stay in registers between the max, the
// sum and the normalisation); otherwise the chain, axb_softmax and axb_reduce run one after the other with the same result.
axb_status axb_elementwise_softmax(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int is_log,
                                   int32_t epilogue_op, const axb_tensor *out, axb_stream_t stream) {
    AXB_REQUIRE(prog && n_instr > 0 && n_inputs > 0 && inputs, AXB_ERR_INVALID, "axb_elementwise_softmax: bad arguments");
    AXB_TRY(validate_tensor(out, "axb_elementwise_softmax(out)"));
    AXB_REQUIRE(epilogue_op == 0 || epilogue_op == AXB_OP_SUM || epilogue_op == AXB_OP_MEAN || epilogue_op == AXB_OP_MAX, AXB_ERR_UNSUPPORTED,
                "axb_elementwise_softmax: the row epilogue is none / sum / mean / max");
    int64_t shape[AXB_MAX_DIMS];
    int nd = inputs[0].ndim;
    memcpy(shape, inputs[0].shape, sizeof(int64_t) * nd);
    for (int k = 1; k < n_inputs; ++k) {
        int64_t nxt[AXB_MAX_DIMS];
        int nn = 0;
        AXB_TRY(axb_broadcast_shape(shape, nd, inputs[k].shape, inputs[k].ndim, nxt, &nn));
        memcpy(shape, nxt, sizeof(int64_t) * nn);
        nd = nn;
    }
    AXB_REQUIRE(nd >= 1, AXB_ERR_SHAPE, "axb_elementwise_softmax: rank-0 chain");
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const int32_t vdt = prog[n_instr - 1].dtype;
    axb_tensor tmp = make_contiguous_desc(nullptr, vdt, shape, nd);
    int64_t rows = 1;
    for (int d = 0; d < nd - 1; ++d)
        rows *= shape[d];
    const int64_t want = epilogue_op ? rows : numel(tmp);
    AXB_REQUIRE(numel(*out) == want, AXB_ERR_SHAPE, "axb_elementwise_softmax: output has %lld elements, expected %lld", (long long)numel(*out),
                (long long)want);
    if (numel(tmp) == 0)
        return AXB_OK;
    static const bool two_pass = getenv("AXB_EW_REDUCE_TWO_PASS") != nullptr; // A/B switch
    bool wide = false;
    for (int k = 0; k < n_inputs; ++k)
        wide = wide || is_64bit_dtype(inputs[k].dtype);
    for (int i = 0; i < n_instr; ++i)
        wide = wide || is_64bit_dtype(prog[i].dtype);
    if (!two_pass && !wide && vdt == AXB_F32 && out->dtype == AXB_F32 && is_contiguous(*out) && !(is_log && epilogue_op)) {
        bool done = false;
        const int epi = epilogue_op == AXB_OP_SUM ? 1 : epilogue_op == AXB_OP_MEAN ? 2 : epilogue_op == AXB_OP_MAX ? 3 : 0;
        AXB_TRY(ew::launch_chain_softmax(prog, n_instr, inputs, n_inputs, &tmp, is_log ? 1 : 0, epi, reinterpret_cast<float *>(tensor_ptr(*out)), s,
                                         ctx, &done));
        if (done)
            return AXB_OK;
    }
    // unfused: chain -> softmax -> (row reduction)
    void *buf = nullptr, *buf2 = nullptr;
    const size_t bytes = (size_t)numel(tmp) * dtype_size(vdt);
    AXB_TRY(axb_alloc(bytes, s, &buf));
    tmp.base = buf;
    axb_status st = axb_elementwise(prog, n_instr, inputs, n_inputs, &tmp, s);
    if (st == AXB_OK && !epilogue_op) {
        axb_tensor od = *out; // same element order; give it the chain's shape
        axb_tensor shaped = make_contiguous_desc(tensor_ptr(*out), out->dtype, shape, nd);
        (void)od;
        st = is_contiguous(*out) ? axb_softmax(&tmp, nd - 1, is_log, &shaped, s) : fail(AXB_ERR_UNSUPPORTED, "axb_elementwise_softmax: output must be contiguous");
    } else if (st == AXB_OK) {
        st = axb_alloc(bytes, s, &buf2);
        if (st == AXB_OK) {
            axb_tensor sm = make_contiguous_desc(buf2, vdt, shape, nd);
            st = axb_softmax(&tmp, nd - 1, is_log, &sm, s);
            const int32_t ax = nd - 1;
            if (st == AXB_OK) {
                int64_t rshape[AXB_MAX_DIMS];
                for (int d = 0; d < nd - 1; ++d)
                    rshape[d] = shape[d];
                axb_tensor rd = make_contiguous_desc(tensor_ptr(*out), out->dtype, rshape, nd - 1);
                st = is_contiguous(*out) ? axb_reduce(epilogue_op, &sm, &ax, 1, 0, &rd, s) : fail(AXB_ERR_UNSUPPORTED, "axb_elementwise_softmax: output must be contiguous");
            }
        }
    }
    if (buf2)
        axb_free(buf2, s);
    axb_free(buf, s);
    return st;
}

// layer_norm / rms_norm over the LAST axis of a chain's value (post-LN transformers: layer_norm(x + sublayer(x))). One kernel for
// fp32 chains with rows of at most 8192 elements and contiguous fp32 weight / bias rows; otherwise the chain and the stand-alone
// norm kernel run back to back.
axb_status axb_elementwise_norm(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t kind,
                                const axb_tensor *weight, const axb_tensor *bias, float eps, const axb_tensor *out, axb_stream_t stream) {
    AXB_REQUIRE(prog && n_instr > 0 && n_inputs > 0 && inputs, AXB_ERR_INVALID, "axb_elementwise_norm: bad arguments");
    AXB_REQUIRE(kind == AXB_OP_LAYER_NORM || kind == AXB_OP_RMS_NORM, AXB_ERR_INVALID, "axb_elementwise_norm: kind is AXB_OP_LAYER_NORM or AXB_OP_RMS_NORM");
    AXB_REQUIRE(kind == AXB_OP_LAYER_NORM || !bias, AXB_ERR_INVALID, "axb_elementwise_norm: rms_norm has no bias");
    AXB_TRY(validate_tensor(out, "axb_elementwise_norm(out)"));
    int64_t shape[AXB_MAX_DIMS];
    int nd = inputs[0].ndim;
    memcpy(shape, inputs[0].shape, sizeof(int64_t) * nd);
    for (int k = 1; k < n_inputs; ++k) {
        int64_t nxt[AXB_MAX_DIMS];
        int nn = 0;
        AXB_TRY(axb_broadcast_shape(shape, nd, inputs[k].shape, inputs[k].ndim, nxt, &nn));
        memcpy(shape, nxt, sizeof(int64_t) * nn);
        nd = nn;
    }
    AXB_REQUIRE(nd >= 1, AXB_ERR_SHAPE, "axb_elementwise_norm: rank-0 chain");
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const int32_t vdt = prog[n_instr - 1].dtype;
    axb_tensor tmp = make_contiguous_desc(nullptr, vdt, shape, nd);
    AXB_REQUIRE(numel(*out) == numel(tmp) && is_contiguous(*out), AXB_ERR_SHAPE, "axb_elementwise_norm: output must be contiguous with the chain's shape");
    if (numel(tmp) == 0)
        return AXB_OK;
    const int64_t n = shape[nd - 1];
    auto row_ok = [&](const axb_tensor *t) { return !t || (t->dtype == AXB_F32 && numel(*t) == n && is_contiguous(*t)); };
    static const bool two_pass = getenv("AXB_EW_REDUCE_TWO_PASS") != nullptr; // A/B switch
    bool wide = false;
    for (int k = 0; k < n_inputs; ++k)
        wide = wide || is_64bit_dtype(inputs[k].dtype);
    for (int i = 0; i < n_instr; ++i)
        wide = wide || is_64bit_dtype(prog[i].dtype);
    if (!two_pass && !wide && vdt == AXB_F32 && out->dtype == AXB_F32 && row_ok(weight) && row_ok(bias)) {
        bool done = false;
        AXB_TRY(ew::launch_chain_softmax(prog, n_instr, inputs, n_inputs, &tmp, kind == AXB_OP_LAYER_NORM ? 2 : 3, 0,
                                         reinterpret_cast<float *>(tensor_ptr(*out)), s, ctx, &done,
                                         weight ? reinterpret_cast<const float *>(tensor_ptr(*weight)) : nullptr,
                                         bias ? reinterpret_cast<const float *>(tensor_ptr(*bias)) : nullptr, eps));
        if (done)
            return AXB_OK;
    }
    void *buf = nullptr;
    AXB_TRY(axb_alloc((size_t)numel(tmp) * dtype_size(vdt), s, &buf));
    tmp.base = buf;
    axb_status st = axb_elementwise(prog, n_instr, inputs, n_inputs, &tmp, s);
    if (st == AXB_OK) {
        const axb_tensor shaped = make_contiguous_desc(tensor_ptr(*out), out->dtype, shape, nd);
        st = kind == AXB_OP_LAYER_NORM ? axb_layer_norm(&tmp, weight, bias, nd - 1, eps, &shaped, s) : axb_rms_norm(&tmp, weight, nd - 1, eps, &shaped, s);
    }
    axb_free(buf, s);
    return st;
}

} // extern "C"
