// cuda_operations.cpp — seam 2: ops::Operation subclasses for Device::GPU, registered with the OperationRegistry
// (include/axiom/operations.hpp:210-302), and the free functions the shared code calls on its GPU paths
// (src/backends/metal/metal_operations.hpp:11-67). Every entry validates like the CPU operation it mirrors, allocates a
// fresh GPU result and ends in one or a few axb_* launches on the backend stream; nothing is computed on the host and
// nothing falls back to the CPU operation.
#include <algorithm>
#include <cmath>
#include <numeric>

#include "axiom/operations.hpp"
#include "backends/metal/metal_operations.hpp"
#include "cuda_backend.hpp"
#include "cuda_plan.hpp"

namespace axiom {
namespace backends {
namespace cuda {

using ops::OpType;

namespace {

Shape shape_of(const int64_t *s, int n) {
    Shape out(n);
    for (int i = 0; i < n; ++i)
        out[i] = static_cast<size_t>(s[i]);
    return out;
}

Shape broadcast2(const Shape &a, const Shape &b) {
    int64_t sa[AXB_MAX_DIMS], sb[AXB_MAX_DIMS], so[AXB_MAX_DIMS];
    if (a.size() > AXB_MAX_DIMS || b.size() > AXB_MAX_DIMS)
        throw ShapeError("CUDA backend supports at most 8 dimensions");
    for (size_t i = 0; i < a.size(); ++i) sa[i] = (int64_t)a[i];
    for (size_t i = 0; i < b.size(); ++i) sb[i] = (int64_t)b[i];
    int n = 0;
    if (axb_broadcast_shape(sa, (int)a.size(), sb, (int)b.size(), so, &n) != AXB_OK)
        throw ShapeError::broadcast_incompatible(axb_last_error());
    return shape_of(so, n);
}

// graph_registry.cpp:95-146 (normalises negative axes; all-reduced -> {1}, argmax/argmin -> rank 0)
Shape reduced_shape(const Shape &in, std::vector<int> axes, bool keep_dims, bool is_arg) {
    if (axes.empty())
        return keep_dims ? Shape(in.size(), 1) : (is_arg ? Shape{} : Shape{1});
    for (int &a : axes)
        if (a < 0)
            a += (int)in.size();
    Shape out;
    for (size_t i = 0; i < in.size(); ++i) {
        const bool red = std::find(axes.begin(), axes.end(), (int)i) != axes.end();
        if (!red)
            out.push_back(in[i]);
        else if (keep_dims)
            out.push_back(1);
    }
    if (out.empty())
        return is_arg ? Shape{} : Shape{1};
    return out;
}

} // namespace

Tensor gpu_unary(OpType op, const Tensor &x_in, double param) {
    const Tensor x = on_gpu(x_in);
    const int32_t rdt = axb_unary_result_dtype(to_axb_op(op), to_axb_dtype(x.dtype()));
    if (rdt < 0)
        throw TypeError::unsupported_dtype(x.dtype_name(), op_name(op) + " on GPU");
    Tensor out = empty_gpu(x.shape(), static_cast<DType>(rdt));
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    check(plan::unary(to_axb_op(op), &xd, &od, param), "axb_unary");
    return out;
}

Tensor gpu_binary(OpType op, const Tensor &a_in, const Tensor &b_in) {
    const Tensor a = on_gpu(a_in), b = on_gpu(b_in);
    const int32_t rdt = axb_binary_result_dtype(to_axb_op(op), to_axb_dtype(a.dtype()), to_axb_dtype(b.dtype()));
    if (rdt < 0)
        throw TypeError::unsupported_dtype(a.dtype_name() + "/" + b.dtype_name(), op_name(op) + " on GPU");
    Tensor out = empty_gpu(broadcast2(a.shape(), b.shape()), static_cast<DType>(rdt));
    if (out.size() == 0)
        return out;
    const axb_tensor ad = describe(a), bd = describe(b), od = describe(out);
    check(plan::binary(to_axb_op(op), &ad, &bd, &od), "axb_binary");
    return out;
}

Tensor gpu_where(const Tensor &c_in, const Tensor &a_in, const Tensor &b_in) {
    const Tensor c = on_gpu(c_in), a = on_gpu(a_in), b = on_gpu(b_in);
    const int32_t rdt = axb_promote_types(to_axb_dtype(a.dtype()), to_axb_dtype(b.dtype()));
    if (rdt < 0)
        throw TypeError("where: cannot promote " + a.dtype_name() + " and " + b.dtype_name());
    Tensor out = empty_gpu(broadcast2(c.shape(), broadcast2(a.shape(), b.shape())), static_cast<DType>(rdt));
    if (out.size() == 0)
        return out;
    const axb_tensor cd = describe(c), ad = describe(a), bd = describe(b), od = describe(out);
    check(plan::where(&cd, &ad, &bd, &od), "axb_where");
    return out;
}

Tensor gpu_masked_fill(const Tensor &x_in, const Tensor &m_in, double value) {
    const Tensor x = on_gpu(x_in), m = on_gpu(m_in);
    Tensor out = empty_gpu(broadcast2(x.shape(), m.shape()), x.dtype());
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), md = describe(m), od = describe(out);
    check(plan::masked_fill(&xd, &md, value, &od), "axb_masked_fill");
    return out;
}

Tensor gpu_cast(const Tensor &x_in, DType to) {
    const Tensor x = on_gpu(x_in);
    Tensor out = empty_gpu(x.shape(), to);
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    check(plan::cast(&xd, &od), "axb_cast");
    return out;
}

Tensor gpu_contiguous(const Tensor &x) {
    if (x.is_contiguous())
        return x;
    Tensor out = empty_gpu(x.shape(), x.dtype());
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    check(plan::gather_strided(&xd, &od), "axb_gather_strided");
    return out;
}

// fresh storage holding exactly the view's elements (the patched Tensor::to compacts offset views before a transfer)
Tensor gpu_compact(const Tensor &x) {
    Tensor out = empty_gpu(x.shape(), x.dtype());
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    check(plan::gather_strided(&xd, &od), "axb_gather_strided");
    return out;
}

Tensor gpu_reduce(OpType op, const Tensor &x_in, const std::vector<int> &axes_in, bool keep_dims) {
    Tensor x = on_gpu(x_in);
    const bool is_arg = op == OpType::ArgMax || op == OpType::ArgMin;
    if (op == OpType::Softmax || op == OpType::LogSoftmax)
        return gpu_softmax(x, axes_in.empty() ? -1 : axes_in[0], op == OpType::LogSoftmax);
    if (is_arg) {
        // cpu_operations.cpp:2240-2247: no axis, axis == -1 or several axes mean "flatten, then axis 0"
        int ax = axes_in.empty() ? -1 : axes_in[0];
        const bool flat = ax == -1 || axes_in.size() > 1;
        if (!flat && (ax < -(int)x.ndim() || ax >= (int)x.ndim()))
            throw ShapeError::invalid_axis(ax, (int)x.ndim());
        if (!flat && ax < 0)
            ax += (int)x.ndim();
        if (x.size() == 0)
            throw ValueError("argmax/argmin of an empty tensor");
        Shape oshape;
        if (flat) {
            x = gpu_contiguous(x);
            oshape = keep_dims ? Shape(x.ndim(), 1) : Shape{};
        } else {
            oshape = reduced_shape(x.shape(), {ax}, keep_dims, true);
        }
        Tensor out = empty_gpu(oshape, DType::Int64);
        const axb_tensor xd = describe(x), od = describe(out);
        check(plan::argreduce(to_axb_op(op), &xd, flat ? -1 : ax, keep_dims ? 1 : 0, &od), "axb_argreduce");
        return out;
    }
    std::vector<int32_t> axes;
    for (int a : axes_in) {
        if (a < -(int)x.ndim() || a >= (int)x.ndim())
            throw ShapeError::invalid_axis(a, (int)x.ndim());
        axes.push_back(a < 0 ? a + (int)x.ndim() : a);
    }
    std::sort(axes.begin(), axes.end());
    axes.erase(std::unique(axes.begin(), axes.end()), axes.end());
    const int32_t rdt = axb_reduce_result_dtype(to_axb_op(op), to_axb_dtype(x.dtype()));
    if (rdt < 0)
        throw TypeError::unsupported_dtype(x.dtype_name(), op_name(op) + " on GPU");
    DType odt = static_cast<DType>(rdt);
    Tensor out = empty_gpu(reduced_shape(x.shape(), axes_in, keep_dims, false), odt);
    const axb_tensor xd = describe(x), od = describe(out);
    check(plan::reduce(to_axb_op(op), &xd, axes.data(), (int)axes.size(), keep_dims ? 1 : 0, &od), "axb_reduce");
    // the graph declares Mean -> Float32 for every non-f64 input (graph_registry.cpp:278-287) and the executor upcasts
    // half results (graph_executor.cpp:396-403); integer means arrive as Float32 from axb_reduce already
    if (op == OpType::Mean && (out.dtype() == DType::Float16 || out.dtype() == DType::BFloat16))
        return gpu_cast(out, DType::Float32);
    return out;
}

Tensor gpu_softmax(const Tensor &x_in, int axis, bool is_log) {
    const Tensor x = on_gpu(x_in);
    if (axis < -(int)x.ndim() || axis >= (int)x.ndim())
        throw ShapeError::invalid_axis(axis, (int)x.ndim());
    Tensor out = empty_gpu(x.shape(), x.dtype());
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    check(plan::softmax(&xd, axis, is_log ? 1 : 0, &od), "axb_softmax");
    return out;
}

Tensor gpu_matmul(const Tensor &a_in, const Tensor &b_in, bool ta, bool tb, const Tensor *bias, OpType activation, int const_ops) {
    const Tensor a = on_gpu(a_in), b = on_gpu(b_in);
    const axb_tensor ad = describe(a), bd = describe(b);
    int64_t oshape[AXB_MAX_DIMS];
    int ond = 0;
    check(axb_matmul_shape(&ad, &bd, ta ? 1 : 0, tb ? 1 : 0, oshape, &ond), "matmul");
    const int32_t rdt = axb_promote_types(ad.dtype, bd.dtype);
    if (rdt < 0)
        throw TypeError("matmul: cannot promote " + a.dtype_name() + " and " + b.dtype_name());
    Tensor out = empty_gpu(shape_of(oshape, ond), static_cast<DType>(rdt));
    if (out.size() == 0)
        return out;
    const axb_tensor od = describe(out);
    axb_matmul_epilogue epi{};
    axb_tensor biasd{};
    Tensor bias_gpu;
    if (bias && bias->storage()) {
        bias_gpu = on_gpu(*bias);
        biasd = describe(bias_gpu);
        epi.bias = &biasd;
    }
    if (activation != OpType::_Count)
        epi.activation = to_axb_op(activation);
    epi.flags = const_ops;
    const bool use_epi = epi.bias || epi.activation || epi.flags;
    check(plan::matmul(&ad, &bd, ta ? 1 : 0, tb ? 1 : 0, use_epi ? &epi : nullptr, &od), "axb_matmul");
    return out;
}

Tensor gpu_pad_constant(const Tensor &x_in, const std::vector<std::pair<size_t, size_t>> &pad, double value) {
    const Tensor x = on_gpu(x_in);
    if (pad.size() != x.ndim())
        throw ShapeError("pad: pad_width must have one (before, after) pair per dimension");
    Shape oshape(x.ndim());
    int64_t lo[AXB_MAX_DIMS], hi[AXB_MAX_DIMS];
    for (size_t d = 0; d < x.ndim(); ++d) {
        lo[d] = (int64_t)pad[d].first;
        hi[d] = (int64_t)pad[d].second;
        oshape[d] = x.shape()[d] + pad[d].first + pad[d].second;
    }
    Tensor out = empty_gpu(oshape, x.dtype());
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    check(plan::pad_constant(&xd, lo, hi, value, &od), "axb_pad_constant");
    return out;
}

// Forward convolution, 1-D or 2-D (axb_conv): implicit GEMM on the tensor cores for groups == 1 and >= 1e8 flop (bias and
// `activation` in the GEMM epilogue), the fp64-accumulating gather kernel otherwise (bit-identical to the CPU loops of
// operations.cpp:3054-3130 / :3263-3318). x: (N, C, spatial...) or unbatched; operands are promoted to one float dtype.
Tensor gpu_conv_nd(const Tensor &x_in, const Tensor &w_in, const Tensor &b_in, int nsp, const int *stride, const int *padding,
                   const int *dilation, int groups, OpType activation) {
    const char *what = nsp == 1 ? "conv1d" : "conv2d";
    Tensor x = on_gpu(x_in);
    Tensor w = on_gpu(w_in);
    const bool has_batch = (int)x.ndim() == nsp + 2;
    if (!has_batch) {
        Shape sh{1};
        for (auto d : x.shape())
            sh.push_back(d);
        x = x.reshape(sh);
    }
    if ((int)x.ndim() != nsp + 2 || (int)w.ndim() != nsp + 2)
        throw ShapeError(std::string(what) + ": unexpected input / weight rank");
    const size_t N = x.shape()[0], C = x.shape()[1], Cout = w.shape()[0], Cg = w.shape()[1];
    if (groups <= 0 || C != Cg * (size_t)groups || Cout % (size_t)groups != 0)
        throw ShapeError(std::string(what) + ": channel/group mismatch");
    Shape oshape{N, Cout};
    for (int d = 0; d < nsp; ++d) {
        const size_t L = x.shape()[2 + d], K = w.shape()[2 + d];
        const size_t span = (size_t)dilation[d] * (K - 1) + 1;
        if (stride[d] <= 0 || dilation[d] <= 0 || padding[d] < 0 || L + 2 * (size_t)padding[d] < span)
            throw ShapeError(std::string(what) + ": input too small for given kernel/padding/dilation");
        oshape.push_back((L + 2 * (size_t)padding[d] - span) / (size_t)stride[d] + 1);
    }
    const DType dt = static_cast<DType>(axb_promote_types(to_axb_dtype(x.dtype()), to_axb_dtype(w.dtype())));
    if (x.dtype() != dt)
        x = gpu_cast(x, dt);
    if (w.dtype() != dt)
        w = gpu_cast(w, dt);
    const bool has_bias = b_in.storage() && b_in.ndim() > 0 && b_in.size() > 0;
    Tensor b;
    if (has_bias) {
        b = on_gpu(b_in);
        if (b.dtype() != dt)
            b = gpu_cast(b, dt);
    }
    Tensor out = empty_gpu(oshape, dt);
    if (out.size() != 0) {
        const axb_tensor xd = describe(x), wd = describe(w), od = describe(out);
        axb_tensor bd{};
        if (has_bias)
            bd = describe(b);
        int32_t st[2], pd[2], dl[2];
        for (int d = 0; d < nsp; ++d) {
            st[d] = stride[d];
            pd[d] = padding[d];
            dl[d] = dilation[d];
        }
        check(plan::conv(&xd, &wd, has_bias ? &bd : nullptr, nsp, st, pd, dl, groups,
                         activation == OpType::_Count ? 0 : to_axb_op(activation), &od),
              "axb_conv");
    }
    if (has_batch)
        return out;
    Shape s1(out.shape().begin() + 1, out.shape().end());
    return out.reshape(s1);
}

Tensor gpu_conv1d_impl(const Tensor &x, const Tensor &w, const Tensor &b, int stride, int padding, int dilation, int groups) {
    return gpu_conv_nd(x, w, b, 1, &stride, &padding, &dilation, groups, OpType::_Count);
}

// ---- Operation classes ---------------------------------------------------------------------------------------------
namespace {

class CudaOperation final : public ops::Operation {
  public:
    explicit CudaOperation(OpType t) : type_(t) {}
    OpType type() const override { return type_; }
    std::string name() const override { return "cuda_" + op_name(type_); }
    Device device() const override { return Device::GPU; }

    // broadcasting is handled in-kernel; returning false would send the eager path to the CPU op (operations.cpp:1122-1127)
    bool supports_binary(const Tensor &lhs, const Tensor &rhs) const override {
        return ops::are_broadcastable(lhs.shape(), rhs.shape());
    }
    Tensor execute_binary(const Tensor &lhs, const Tensor &rhs) const override { return gpu_binary(type_, lhs, rhs); }
    Tensor execute_unary(const Tensor &input) const override {
        return gpu_unary(type_, input, type_ == OpType::LeakyReLU ? 0.01 : 0.0);
    }
    Tensor execute_reduction(const Tensor &input, const std::vector<int> &axis, bool keep_dims) const override {
        if ((type_ == OpType::Softmax || type_ == OpType::LogSoftmax) && input.is_lazy() && input.device() == Device::GPU) {
            Tensor r = graph::fused_softmax(input, axis.empty() ? -1 : axis[0], type_ == OpType::LogSoftmax);
            if (r.storage())
                return r;
        }
        return gpu_reduce(type_, input, axis, keep_dims);
    }
    Tensor execute_matmul(const Tensor &a, const Tensor &b, bool ta, bool tb) const override {
        return gpu_matmul(a, b, ta, tb);
    }
    Tensor execute_where(const Tensor &c, const Tensor &a, const Tensor &b) const override {
        if (c.is_lazy() || a.is_lazy() || b.is_lazy())
            return graph::fused_where(c, a, b); // lazy operand trees + select in one launch
        return gpu_where(c, a, b);
    }
    Tensor execute_masked_fill(const Tensor &input, const Tensor &mask, const Tensor &value) const override {
        // value is a one-element tensor (operations.cpp:1588-1590); a device scalar costs one 4/8-byte read back
        Tensor v = value.device() == Device::CPU ? value : value.cpu();
        const double val = v.astype(DType::Float64).item<double>({0});
        return gpu_masked_fill(input, mask, val);
    }
    Tensor execute_gather(const Tensor &input, int dim, const Tensor &indices) const override {
        const Tensor x = on_gpu(input), idx = on_gpu(indices);
        Tensor out = empty_gpu(idx.shape(), x.dtype());
        if (out.size() == 0)
            return out;
        const axb_tensor xd = describe(x), id = describe(idx), od = describe(out);
        check(axb_gather(&xd, dim, &id, &od, nullptr), "axb_gather");
        return out;
    }
    Tensor execute_index_select(const Tensor &input, int dim, const Tensor &indices) const override {
        const Tensor x = on_gpu(input), idx = on_gpu(indices);
        int d = dim < 0 ? dim + (int)x.ndim() : dim;
        if (d < 0 || d >= (int)x.ndim())
            throw ShapeError::invalid_axis(dim, (int)x.ndim());
        Shape oshape = x.shape();
        oshape[d] = idx.size();
        Tensor out = empty_gpu(oshape, x.dtype());
        if (out.size() == 0)
            return out;
        const axb_tensor xd = describe(x), id = describe(idx), od = describe(out);
        check(axb_index_select(&xd, d, &id, &od, nullptr), "axb_index_select");
        return out;
    }
    Tensor execute_scatter(const Tensor &input, int dim, const Tensor &indices, const Tensor &src) const override {
        const Tensor x = on_gpu(input), idx = on_gpu(indices);
        Tensor v = on_gpu(src);
        if (v.dtype() != x.dtype())
            v = gpu_cast(v, x.dtype());
        Tensor out = empty_gpu(x.shape(), x.dtype());
        if (out.size() == 0)
            return out;
        const axb_tensor xd = describe(x), id = describe(idx), vd = describe(v), od = describe(out);
        check(axb_scatter(&xd, dim, &id, &vd, &od, nullptr), "axb_scatter");
        return out;
    }
    Tensor execute_masked_select(const Tensor &input, const Tensor &mask) const override {
        // CPU rule (cpu_operations.cpp:2719-2776): broadcast, first byte of the mask element, row-major, 1-D result.
        // The element count decides the result's shape: 8 bytes come back inside axb_masked_select, nothing else.
        const Tensor x = on_gpu(input), m = on_gpu(mask);
        const axb_tensor xd = describe(x), md = describe(m);
        void *data = nullptr;
        int64_t count = 0;
        check(axb_masked_select(&xd, &md, &data, &count, nullptr), "axb_masked_select");
        if (count == 0)
            return empty_gpu({0}, x.dtype());
        const size_t bytes = (size_t)count * dtype_size(x.dtype());
        auto st = std::make_shared<CudaStorage>(data, bytes, CudaStorage::Adopt{});
        return Tensor(st, Shape{(size_t)count}, Strides{(int64_t)dtype_size(x.dtype())}, x.dtype(), 0);
    }
    Tensor execute_cast(const Tensor &input, DType target) const override { return gpu_cast(input, target); }
    void execute_binary_inplace(Tensor &lhs, const Tensor &rhs) const override {
        const Tensor r = on_gpu(rhs);
        const int32_t rdt = axb_binary_result_dtype(to_axb_op(type_), to_axb_dtype(lhs.dtype()), to_axb_dtype(r.dtype()));
        if (rdt != to_axb_dtype(lhs.dtype()) || broadcast2(lhs.shape(), r.shape()) != lhs.shape())
            throw TypeError("in-place " + op_name(type_) + ": result does not fit the left operand");
        const axb_tensor ld = describe(lhs), rd = describe(r);
        // a kernel is about to overwrite lhs: cached forms of it (weight splits) are stale from here on
        if (auto *cs = dynamic_cast<CudaStorage *>(lhs.storage().get()))
            axb_invalidate_cached(cs->device_ptr(), cs->size_bytes());
        check(axb_binary(to_axb_op(type_), &ld, &rd, &ld, nullptr), "axb_binary (in place)");
    }

  private:
    OpType type_;
};

} // namespace

// ---- pooling / sort / topk: called from the patched ops::max_pool*/avg_pool*/sort/argsort/topk (patches/apply.py), which
// otherwise copy a GPU tensor to the host, run scalar loops and upload the result (operations.cpp:2367-2880, :4199-4257) --
Tensor gpu_pool(const Tensor &input, bool is_avg, const std::vector<int> &kernel, const std::vector<int> &stride,
                const std::vector<int> &padding, bool count_include_pad, const Shape &out_shape) {
    const int nsp = (int)kernel.size();
    Tensor x = on_gpu(input);
    if (!is_floating_dtype(x.dtype())) // dispatch_float (include/axiom/dispatch.hpp:151-156)
        throw TypeError::unsupported_dtype(dtype_name(x.dtype()), is_avg ? "avg_pool" : "max_pool");
    const bool has_batch = (int)x.ndim() == nsp + 2;
    if (!has_batch)
        x = x.unsqueeze(0);
    Shape oshape = out_shape;
    if (!has_batch)
        oshape.insert(oshape.begin(), 1);
    Tensor out = empty_gpu(oshape, x.dtype());
    if (out.size() != 0) {
        int32_t k[3], s[3], p[3];
        for (int d = 0; d < nsp; ++d) {
            k[d] = kernel[d];
            s[d] = stride[d];
            p[d] = padding[d];
        }
        const axb_tensor xd = describe(x), od = describe(out);
        check(plan::pool(&xd, is_avg ? 1 : 0, nsp, k, s, p, count_include_pad ? 1 : 0, &od), "axb_pool");
    }
    return has_batch ? out : out.squeeze(0);
}

namespace {
// keys the kernels sort: float32 / float64 / int32 / int64; everything else becomes float32 first (operations.cpp:4199-4257)
Tensor sort_keys(const Tensor &input) {
    Tensor x = gpu_contiguous(on_gpu(input));
    switch (x.dtype()) {
    case DType::Float32: case DType::Float64: case DType::Int32: case DType::Int64: return x;
    default: return gpu_cast(x, DType::Float32);
    }
}
int sort_axis(const Tensor &x, int axis) {
    const int nd = (int)x.ndim();
    const int a = axis < 0 ? axis + nd : axis;
    if (a < 0 || a >= nd)
        throw ShapeError::invalid_axis(axis, nd);
    return a;
}
} // namespace

Tensor gpu_sort(const Tensor &input, int axis, bool descending, bool want_indices) {
    const Tensor x = sort_keys(input);
    const int a = sort_axis(x, axis);
    Tensor out = empty_gpu(x.shape(), want_indices ? DType::Int64 : x.dtype());
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    check(plan::sort(&xd, a, descending ? 1 : 0, want_indices ? nullptr : &od, want_indices ? &od : nullptr), "axb_sort");
    return out;
}

std::pair<Tensor, Tensor> gpu_topk(const Tensor &input, int k, int axis, bool largest) {
    const Tensor x = sort_keys(input);
    const int a = sort_axis(x, axis);
    if (k < 0 || (size_t)k > x.shape()[a])
        throw ValueError("topk: k = " + std::to_string(k) + " out of range for an axis of " + std::to_string(x.shape()[a]));
    Shape oshape = x.shape();
    oshape[a] = (size_t)k;
    Tensor values = empty_gpu(oshape, x.dtype()), indices = empty_gpu(oshape, DType::Int64);
    if (values.size() != 0) {
        const axb_tensor xd = describe(x), vd = describe(values), id = describe(indices);
        check(plan::sort(&xd, a, largest ? 1 : 0, &vd, &id), "axb_sort (topk)");
    }
    return {values, indices};
}

void register_cuda_operations() {
    static const OpType kOps[] = {
        OpType::Add, OpType::Subtract, OpType::Multiply, OpType::Divide, OpType::Power, OpType::Modulo,
        OpType::Equal, OpType::NotEqual, OpType::Less, OpType::LessEqual, OpType::Greater, OpType::GreaterEqual,
        OpType::LogicalAnd, OpType::LogicalOr, OpType::LogicalXor, OpType::LogicalNot,
        OpType::BitwiseAnd, OpType::BitwiseOr, OpType::BitwiseXor, OpType::LeftShift, OpType::RightShift,
        OpType::Maximum, OpType::Minimum, OpType::Atan2, OpType::Hypot,
        OpType::Negate, OpType::Abs, OpType::Sqrt, OpType::Exp, OpType::Log, OpType::Sin, OpType::Cos, OpType::Tan,
        OpType::Erf, OpType::Sign, OpType::Floor, OpType::Ceil, OpType::Trunc, OpType::Round, OpType::Reciprocal,
        OpType::Square, OpType::Cbrt, OpType::IsNaN, OpType::IsInf, OpType::IsFinite,
        OpType::ReLU, OpType::LeakyReLU, OpType::SiLU, OpType::Sigmoid, OpType::Tanh, OpType::GELU,
        OpType::Softmax, OpType::LogSoftmax,
        OpType::Sum, OpType::Mean, OpType::Max, OpType::Min, OpType::ArgMax, OpType::ArgMin, OpType::Any, OpType::All,
        OpType::Prod,
        OpType::MatMul, OpType::BatchMatMul, OpType::Where, OpType::MaskedFill, OpType::Gather, OpType::Scatter,
        OpType::IndexSelect, OpType::MaskedSelect,
        OpType::Cast,
    };
    for (OpType t : kOps)
        ops::OperationRegistry::register_operation(t, Device::GPU, std::make_unique<CudaOperation>(t));
}

} // namespace cuda

// ---- free functions of the GPU seam (declared in src/backends/metal/metal_operations.hpp) --------------------------
namespace metal {
namespace plan = cuda::plan;

using cuda::check;
using cuda::describe;
using cuda::empty_gpu;
using cuda::on_gpu;

Tensor gpu_make_contiguous(const Tensor &tensor) { return cuda::gpu_contiguous(tensor); }

Tensor gpu_layer_norm(const Tensor &input, const Tensor &weight, const Tensor &bias, int axis, float eps) {
    const Tensor x = on_gpu(input);
    Tensor out = empty_gpu(x.shape(), x.dtype()); // the lazy LayerNorm node declares the input dtype (graph_registry.cpp:625)
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    axb_tensor wd{}, bd{};
    Tensor w, b;
    const bool has_w = weight.storage() && weight.size() > 0, has_b = bias.storage() && bias.size() > 0;
    if (has_w) { w = on_gpu(weight); wd = describe(w); }
    if (has_b) { b = on_gpu(bias); bd = describe(b); }
    check(plan::layer_norm(&xd, has_w ? &wd : nullptr, has_b ? &bd : nullptr, axis, eps, &od), "axb_layer_norm");
    return out;
}

Tensor gpu_rms_norm(const Tensor &input, const Tensor &weight, int axis, float eps) {
    if (input.is_lazy() && input.device() == Device::GPU) { // elementwise producers of the operand as the kernel's prologue
        Tensor r = graph::fused_rms_norm(input, weight, axis, eps);
        if (r.storage())
            return r;
    }
    const Tensor x = on_gpu(input);
    Tensor out = empty_gpu(x.shape(), x.dtype());
    if (out.size() == 0)
        return out;
    const axb_tensor xd = describe(x), od = describe(out);
    axb_tensor wd{};
    Tensor w;
    const bool has_w = weight.storage() && weight.size() > 0;
    if (has_w) { w = on_gpu(weight); wd = describe(w); }
    check(plan::rms_norm(&xd, has_w ? &wd : nullptr, axis, eps, &od), "axb_rms_norm");
    return out;
}

Tensor gpu_conv1d(const Tensor &input, const Tensor &weight, const Tensor &bias, int stride, int padding, int dilation,
                  int groups) {
    return cuda::gpu_conv1d_impl(input, weight, bias, stride, padding, dilation, groups);
}

// softmax(Q K^T * scale + mask) V as batched GEMM -> masked scale -> row softmax -> batched GEMM. Semantics follow
// cpu_flash_attention (src/backends/cpu/flash_attention_cpu.cpp:73-105): causal masks kv > q, an explicit mask marks
// positions to DROP (non-zero = masked out), both by -inf before the softmax.
Tensor gpu_scaled_dot_product_attention(const Tensor &query, const Tensor &key, const Tensor &value, const Tensor &mask,
                                        float scale, bool is_causal) {
    const Tensor q = on_gpu(query), k = on_gpu(key), v = on_gpu(value);
    const bool f32 = q.dtype() == DType::Float32, half = q.dtype() == DType::BFloat16 || q.dtype() == DType::Float16;
    const bool narrow = q.ndim() == 4 && v.ndim() == 4 && q.shape()[3] <= 64 && v.shape()[3] <= 64; // the flash kernel's head_dim range
    if (k.dtype() == q.dtype() && v.dtype() == q.dtype() && (f32 || (half && narrow))) {
        // one C-ABI call. head_dim <= 64: ONE tcgen05 kernel, online softmax, no score matrix (csrc/flash_attention.cu),
        // fp32 through the exact three-piece split, bf16 / fp16 native; wider fp32 heads: Q K^T, fused scale + mask +
        // softmax writing pre-split probabilities, P V (csrc/attention.cu)
        Tensor out = empty_gpu({q.shape()[0], q.shape()[1], q.shape()[2], v.shape()[3]}, q.dtype());
        if (out.size() == 0)
            return out;
        Tensor m;
        axb_tensor md{};
        const bool has_mask = mask.storage() && mask.size() > 0;
        if (has_mask) {
            m = on_gpu(mask);
            if (m.itemsize() != 1)
                m = cuda::gpu_binary(ops::OpType::NotEqual, m, Tensor::zeros({1}, m.dtype(), Device::CPU).to(Device::GPU));
            md = describe(m);
        }
        const axb_tensor qd = describe(q), kd = describe(k), vd = describe(v), od = describe(out);
        check(axb_sdpa(&qd, &kd, &vd, has_mask ? &md : nullptr, scale, is_causal ? 1 : 0, &od, nullptr), "axb_sdpa");
        return out;
    }
    // scale the (B, H, Lq, D) queries, not the (B, H, Lq, Lk) scores: Lk / D times less data through the kernel
    Tensor sc = Tensor::full({1}, scale, Device::CPU).astype(q.dtype()).to(Device::GPU);
    Tensor s = cuda::gpu_matmul(cuda::gpu_binary(ops::OpType::Multiply, q, sc), k, false, true); // (B, H, Lq, Lk)
    const size_t Lq = q.shape()[2], Lk = k.shape()[2];
    if (is_causal) {
        // kv position j is visible to query i iff j <= i (q_block_offset + i, flash_attention_cpu.cpp:76-89)
        Tensor cm = Tensor::zeros({Lq, Lk}, DType::Bool, Device::CPU);
        auto *p = cm.typed_data<bool>();
        for (size_t i = 0; i < Lq; ++i)
            for (size_t j = i + 1; j < Lk; ++j)
                p[i * Lk + j] = true;
        s = cuda::gpu_masked_fill(s, cm.to(Device::GPU), -INFINITY);
    }
    if (mask.storage() && mask.size() > 0) {
        Tensor m = on_gpu(mask);
        if (m.dtype() != DType::Bool)
            m = cuda::gpu_binary(ops::OpType::NotEqual, m, Tensor::zeros({1}, m.dtype(), Device::CPU).to(Device::GPU));
        s = cuda::gpu_masked_fill(s, m, -INFINITY);
    }
    Tensor p = cuda::gpu_softmax(s, -1, false);
    return cuda::gpu_matmul(p, v, false, false); // (B, H, Lq, D)
}

Tensor gpu_pad(const Tensor &input, const std::vector<std::pair<size_t, size_t>> &pad_width, const std::string &mode,
               double value) {
    if (mode != "constant")
        throw RuntimeError::not_implemented("GPU pad mode '" + mode + "' (constant only)");
    return cuda::gpu_pad_constant(input, pad_width, value);
}

Tensor gpu_fill_constant(const Shape &shape, DType dtype, double value) {
    Tensor out = empty_gpu(shape, dtype);
    if (out.size() == 0)
        return out;
    const axb_tensor od = describe(out);
    check(plan::fill(&od, value), "axb_fill");
    return out;
}

Tensor gpu_linear(const Tensor &input, const Tensor &weight, const Tensor &bias, bool has_bias) {
    return cuda::gpu_matmul(input, weight, false, true, has_bias ? &bias : nullptr);
}

// ---- spatial operators next to the path: csrc/spatial.cu ------------------------------------------
namespace {
// common dtype of input / weight / bias; tensors moved to the device and cast where needed
struct ConvArgs {
    Tensor x, w, b;
    bool has_bias = false, has_batch = true;
};
ConvArgs conv_args(const Tensor &input, const Tensor &weight, const Tensor &bias, size_t batched_ndim, const char *what) {
    ConvArgs a;
    a.x = cuda::on_gpu(input);
    a.w = cuda::on_gpu(weight);
    if (a.x.ndim() + 1 == batched_ndim) {
        a.has_batch = false;
        Shape s{1};
        for (auto d : a.x.shape())
            s.push_back(d);
        a.x = a.x.reshape(s);
    }
    if (a.x.ndim() != batched_ndim || a.w.ndim() != batched_ndim)
        throw ShapeError(std::string(what) + ": unexpected input / weight rank");
    const DType dt = static_cast<DType>(axb_promote_types(cuda::to_axb_dtype(a.x.dtype()), cuda::to_axb_dtype(a.w.dtype())));
    if (a.x.dtype() != dt)
        a.x = cuda::gpu_cast(a.x, dt);
    if (a.w.dtype() != dt)
        a.w = cuda::gpu_cast(a.w, dt);
    a.has_bias = bias.storage() && bias.ndim() > 0 && bias.size() > 0;
    if (a.has_bias) {
        a.b = cuda::on_gpu(bias);
        if (a.b.dtype() != dt)
            a.b = cuda::gpu_cast(a.b, dt);
    }
    return a;
}
Tensor drop_batch(const Tensor &out, bool has_batch) {
    if (has_batch)
        return out;
    Shape s(out.shape().begin() + 1, out.shape().end());
    return out.reshape(s);
}
} // namespace

// ops::conv2d on the GPU arrives as a lazy OpType::Conv2D node (operations.cpp:3211-3220) and is lowered here
Tensor gpu_conv2d(const Tensor &input, const Tensor &weight, const Tensor &bias, std::array<int, 2> stride, std::array<int, 2> padding,
                  std::array<int, 2> dilation, int groups) {
    return cuda::gpu_conv_nd(input, weight, bias, 2, stride.data(), padding.data(), dilation.data(), groups, ops::OpType::_Count);
}

Tensor gpu_conv_transpose2d(const Tensor &input, const Tensor &weight, const Tensor &bias, std::array<int, 2> stride,
                            std::array<int, 2> padding, std::array<int, 2> output_padding, std::array<int, 2> dilation, int groups) {
    ConvArgs a = conv_args(input, weight, bias, 4, "conv_transpose2d");
    const size_t H = a.x.shape()[2], W = a.x.shape()[3], KH = a.w.shape()[2], KW = a.w.shape()[3];
    // output extent exactly as the reference computes it (operations.cpp:3546-3553)
    const size_t OH = (H - 1) * (size_t)stride[0] - 2 * (size_t)padding[0] + (size_t)dilation[0] * (KH - 1) + (size_t)output_padding[0] + 1;
    const size_t OW = (W - 1) * (size_t)stride[1] - 2 * (size_t)padding[1] + (size_t)dilation[1] * (KW - 1) + (size_t)output_padding[1] + 1;
    Tensor out = cuda::empty_gpu({a.x.shape()[0], a.w.shape()[1] * (size_t)groups, OH, OW}, a.x.dtype());
    if (out.size() > 0) {
        const axb_tensor xd = cuda::describe(a.x), wd = cuda::describe(a.w), od = cuda::describe(out);
        axb_tensor bd{};
        if (a.has_bias)
            bd = cuda::describe(a.b);
        const int32_t st[2] = {stride[0], stride[1]}, pd[2] = {padding[0], padding[1]}, dl[2] = {dilation[0], dilation[1]};
        cuda::check(axb_conv_transpose(&xd, &wd, a.has_bias ? &bd : nullptr, 2, st, pd, dl, groups, &od, nullptr), "axb_conv_transpose");
    }
    return drop_batch(out, a.has_batch);
}

Tensor gpu_conv_transpose1d(const Tensor &input, const Tensor &weight, const Tensor &bias, int stride, int padding, int output_padding,
                            int dilation, int groups) {
    ConvArgs a = conv_args(input, weight, bias, 3, "conv_transpose1d");
    const size_t L = a.x.shape()[2], K = a.w.shape()[2];
    const size_t OL = (L - 1) * (size_t)stride - 2 * (size_t)padding + (size_t)dilation * (K - 1) + (size_t)output_padding + 1; // :3386-3389
    Tensor out = cuda::empty_gpu({a.x.shape()[0], a.w.shape()[1] * (size_t)groups, OL}, a.x.dtype());
    if (out.size() > 0) {
        const axb_tensor xd = cuda::describe(a.x), wd = cuda::describe(a.w), od = cuda::describe(out);
        axb_tensor bd{};
        if (a.has_bias)
            bd = cuda::describe(a.b);
        const int32_t st = stride, pd = padding, dl = dilation;
        cuda::check(axb_conv_transpose(&xd, &wd, a.has_bias ? &bd : nullptr, 1, &st, &pd, &dl, groups, &od, nullptr), "axb_conv_transpose");
    }
    return drop_batch(out, a.has_batch);
}

Tensor gpu_interpolate(const Tensor &input, const std::vector<size_t> &size, ops::InterpolateMode mode, bool align_corners) {
    const Tensor x = cuda::on_gpu(input);
    if (x.ndim() != 3 && x.ndim() != 4)
        throw ValueError("interpolate: only 1D and 2D spatial interpolation supported, got " + std::to_string(x.ndim() - 2) + "D");
    Shape os{x.shape()[0], x.shape()[1]};
    for (auto v : size)
        os.push_back(v);
    Tensor out = cuda::empty_gpu(os, x.dtype());
    if (out.size() > 0) {
        const axb_tensor xd = cuda::describe(x), od = cuda::describe(out);
        const int m = mode == ops::InterpolateMode::Nearest ? 0 : mode == ops::InterpolateMode::Bilinear ? 1 : 2;
        cuda::check(axb_interpolate(&xd, m, align_corners ? 1 : 0, &od, nullptr), "axb_interpolate");
    }
    return out;
}

} // namespace metal
} // namespace backends
} // namespace axiom
