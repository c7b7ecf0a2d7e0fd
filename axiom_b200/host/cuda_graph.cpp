// cuda_graph.cpp — seam 3: lowering of Axiom's lazy DAG onto the B200 kernel families.
//
//   materialize_gpu_graph(root)                 contract: src/backends/metal/mpsgraph_graph_compiler.mm:706-944
//   execute_gpu_fused_chain / _fused_reduction  contract: src/backends/metal/mpsgraph_fused.mm:409-560,
//                                               include/axiom/graph/gpu_fusion.hpp:14-22
//
// Lowering rules (B200-first, not a translation of the MPSGraph builder):
//   * After the shared sub-DAGs (ref_count > 1) are materialised, the remaining lazy nodes under `root` form a TREE.
//     Every maximal connected sub-tree of elementwise nodes (unary, binary, comparison, logical, masked_fill, and the
//     broadcasts between them) becomes ONE axb_elementwise launch: a typed straight-line program over <= 8 strided
//     operands. Intermediates never touch HBM; scalars and {4096,1}-style operands are read with stride 0. The
//     reference's CPU fusion only follows the first input of a chain and needs equal shapes
//     (src/graph/graph_compiler.cpp:103-140, graph_executor.cpp:544-546); here the fusion unit is the DAG.
//   * elementwise tree + FULL reduction on top -> axb_elementwise_reduce (one pass, no materialised chain result).
//   * MatMul [+ bias add] [+ activation] -> one tcgen05 GEMM launch with a fused epilogue (nn::Linear + GELU).
//   * Reshape / Transpose / Slice nodes are strided VIEWS (no kernel) whenever the strides allow; the kernels honour
//     byte strides, zero strides and offsets, so the Metal path's forced gathers are not reproduced.
//   * softmax / layer_norm / reductions / conv1d map 1:1 onto their kernel family.
// Unknown ops raise RuntimeError — no CPU fallback on this path.
#include <algorithm>
#include <cmath>
#include <unordered_map>
#include <unordered_set>

#include "axiom/graph/compiled_graph.hpp"
#include "axiom/graph/gpu_fusion.hpp"
#include "axiom/graph/graph_node.hpp"
#include "backends/metal/metal_operations.hpp"
#include "cuda_backend.hpp"
#include "cuda_plan.hpp"

namespace axiom {
namespace graph {

namespace plan = backends::cuda::plan;
using backends::cuda::check;
using backends::cuda::describe;
using backends::cuda::empty_gpu;
using backends::cuda::on_gpu;
using backends::cuda::to_axb_dtype;
using backends::cuda::to_axb_op;
using ops::OpType;

void materialize_gpu_graph(GraphNode *root);

namespace {

bool is_leaf(const GraphNode *n) { return n->is_constant || n->is_materialized_; }
bool is_absent(const GraphNode *n) { return n->output_shape.empty() && !n->constant_storage && !n->cached_result_; }

// ops that an axb_elementwise program can express
bool is_fusable(const GraphNode *n) {
    const OpType op = n->op_type;
    if (op == OpType::MaskedFill)
        return true;
    if (!(is_unary_op(op) || is_binary_op(op)) || !is_elementwise_op(op))
        return false;
    switch (op) { // elementwise in the traits table but not in the kernel family
    case OpType::Conj: case OpType::Real: case OpType::Imag: case OpType::Softmax: case OpType::LogSoftmax:
        return false;
    default:
        return !is_complex_dtype(n->output_dtype);
    }
}

Tensor leaf_tensor(const GraphNode *n) {
    if (n->is_constant && n->constant_storage)
        return Tensor(n->constant_storage, n->output_shape, n->constant_strides, n->output_dtype, n->constant_offset);
    if (n->cached_result_)
        return Tensor(n->cached_result_, n->cached_shape_, n->cached_strides_, n->output_dtype);
    return Tensor();
}

// Which operands of a MatMul node are CONSTANT leaves of the graph in the weight position: 2-D tensors captured by value
// (nn::Linear's weight, x.matmul(W)). Their storage only changes through paths that invalidate libaxb's caches (storage
// release, copy_from, in-place ops), so the fp32 tensor-core path may keep their three-piece split across calls.
int const_operands(const GraphNode *mm) {
    int f = 0;
    for (int i = 0; i < 2 && i < (int)mm->inputs.size(); ++i) {
        const GraphNode *c = mm->inputs[i].get();
        if (c && c->is_constant && c->constant_storage && c->output_shape.size() == 2 && c->output_dtype == DType::Float32 &&
            c->constant_storage->device() == Device::GPU)
            f |= i == 0 ? AXB_MM_CONST_A : AXB_MM_CONST_B;
        // an ACTIVATION that several matmuls consume (the q / k / v projections of one layer-norm output,
        // examples/whisper_poc.cpp:63-75): it was materialised as a shared sub-graph, so every consumer sees the same storage;
        // its split is kept until that storage is released or written, and the second and third projection reuse it
        else if (i == 0 && c && !c->is_constant && c->is_materialized_ && c->cached_result_ && c->ref_count > 0 &&
                 c->output_dtype == DType::Float32 && c->cached_result_->device() == Device::GPU)
            f |= AXB_MM_CONST_A;
    }
    return f;
}

struct Lowering;
bool try_linear_epilogue(Lowering &L, const GraphNode *n);

struct Lowering {
    std::unordered_map<const GraphNode *, Tensor> value;

    const Tensor &get(const GraphNode *n) {
        auto it = value.find(n);
        if (it != value.end())
            return it->second;
        if (!is_leaf(n))
            throw RuntimeError::internal("CUDA graph lowering: node evaluated before its input");
        Tensor t = leaf_tensor(n);
        if (t.storage() && t.device() != Device::GPU)
            t = t.to(Device::GPU); // scalars created on the host (Tensor::full({1}, v)) etc.
        return value.emplace(n, std::move(t)).first->second;
    }

    // ---- elementwise trees --------------------------------------------------------------------------------------
    struct Program {
        std::vector<axb_instr> instrs;
        std::vector<const GraphNode *> inputs; // operand nodes (leaves or already-evaluated nodes)
    };

    // Appends the instructions of the fusable sub-tree rooted at n; returns the operand selector of its value.
    // A child that would overflow the program / operand limits is evaluated on its own and enters as an input.
    int32_t emit(const GraphNode *n, Program &pg, const GraphNode *tree_root) {
        if (n != tree_root && !is_leaf(n) && !value.count(n))
            try_linear_epilogue(*this, n); // act(matmul + bias) inside a larger tree: GEMM epilogue, then an operand
        const bool inline_it = !is_leaf(n) && !value.count(n) && is_fusable(n) &&
                               (n == tree_root || pg.instrs.size() + 1 < AXB_MAX_INSTRS);
        if (!inline_it) {
            if (!is_leaf(n) && !value.count(n))
                eval(n); // non-fusable producer (matmul, reduction, ...) or an overflowing sub-tree
            for (size_t i = 0; i < pg.inputs.size(); ++i)
                if (pg.inputs[i] == n)
                    return (int32_t)i;
            if (pg.inputs.size() >= AXB_MAX_INPUTS)
                throw std::length_error("operands");
            pg.inputs.push_back(n);
            return (int32_t)pg.inputs.size() - 1;
        }
        axb_instr in{};
        in.op = to_axb_op(n->op_type);
        in.dtype = to_axb_dtype(n->output_dtype);
        in.src[0] = in.src[1] = in.src[2] = AXB_SRC_NONE;
        if (n->op_type == OpType::MaskedFill) { // mask ? value : input  (operations.cpp:1575-1584)
            const int32_t x = emit(n->inputs[0].get(), pg, tree_root);
            const int32_t m = emit(n->inputs[1].get(), pg, tree_root);
            in.op = AXB_OP_WHERE;
            in.src[0] = m;
            in.src[1] = AXB_SRC_IMM;
            in.src[2] = x;
            in.imm = (double)get_params<MaskedFillParams>(n->params).value;
        } else {
            for (size_t i = 0; i < n->inputs.size() && i < 2; ++i)
                in.src[i] = emit(n->inputs[i].get(), pg, tree_root);
            if (n->op_type == OpType::LeakyReLU) {
                // lazy unary nodes carry NoParams (operations.cpp:1394 create_lazy_unary): the slope is the CPU functor's
                // 0.01 (cpu_operations.hpp LeakyReLUFunc) unless the node has ActivationParams
                in.imm = 0.01;
                if (auto *ap = std::get_if<ActivationParams>(&n->params))
                    in.imm = (double)ap->alpha;
            }
        }
        pg.instrs.push_back(in);
        return AXB_SRC_VALUE((int32_t)pg.instrs.size() - 1);
    }

    // Builds the program of the tree rooted at n. If the operand limit is hit, the deepest fusable children are
    // evaluated separately (they become inputs) and the build is retried.
    Program build_program(const GraphNode *n) {
        for (int attempt = 0; attempt < 64; ++attempt) {
            Program pg;
            try {
                emit(n, pg, n);
                return pg;
            } catch (const std::length_error &) {
                // materialise one fusable child sub-tree and try again with fewer operands
                const GraphNode *victim = nullptr;
                for (auto &c : n->inputs)
                    if (c && !is_leaf(c.get()) && !value.count(c.get()) && is_fusable(c.get())) {
                        victim = c.get();
                        break;
                    }
                if (!victim)
                    throw RuntimeError::internal("CUDA graph lowering: elementwise node with more than 8 operands");
                eval(victim);
            }
        }
        throw RuntimeError::internal("CUDA graph lowering: could not split an elementwise tree");
    }

    // softmax / log_softmax over the last axis of a 32-bit float node, with the node's own elementwise producers as a prologue
    // and an optional row reduction of the result as an epilogue. `sm` is the Softmax / LogSoftmax node;
    // returns false when the pattern does not apply (other axis, other dtype, rows too long are decided inside libaxb).
    bool try_fused_softmax(const GraphNode *sm, OpType epilogue, const Shape &out_shape, Tensor *result) {
        if (sm->output_dtype != DType::Float32)
            return false;
        return fused_softmax_core(sm->inputs.at(0).get(), get_params<ActivationParams>(sm->params).axis,
                                  sm->op_type == OpType::LogSoftmax, epilogue, out_shape, result);
    }
    // c = the node whose value is normalised (its own elementwise producers become the prologue)
    bool fused_softmax_core(const GraphNode *c, int ax, bool is_log, OpType epilogue, const Shape &out_shape, Tensor *result) {
        const int nd = (int)c->output_shape.size();
        if (nd < 1 || c->output_dtype != DType::Float32)
            return false;
        if (!(ax == nd - 1 || ax == -1) || c->output_shape.back() < 2 || c->output_shape.back() > 8192 ||
            ShapeUtils::size(c->output_shape) == 0)
            return false;
        Program pg;
        if (!is_leaf(c) && !value.count(c) && is_fusable(c)) {
            pg = build_program(c);
        } else { // nothing to fuse in front: the program is the identity on the (evaluated) input
            if (epilogue == OpType::_Count)
                return false; // plain softmax of a tensor: the stand-alone kernel family
            if (!is_leaf(c) && !value.count(c))
                eval(c);
            axb_instr id{};
            id.op = AXB_OP_CAST;
            id.dtype = AXB_F32;
            id.src[0] = 0;
            id.src[1] = id.src[2] = AXB_SRC_NONE;
            pg.instrs.push_back(id);
            pg.inputs.push_back(c);
        }
        Tensor r = empty_gpu(out_shape, DType::Float32);
        std::vector<axb_tensor> ins(pg.inputs.size());
        for (size_t i = 0; i < pg.inputs.size(); ++i)
            ins[i] = describe(get(pg.inputs[i]));
        const axb_tensor od = describe(r);
        check(plan::elementwise_softmax(pg.instrs.data(), (int)pg.instrs.size(), ins.data(), (int)ins.size(),
                                        is_log ? 1 : 0, epilogue == OpType::_Count ? 0 : to_axb_op(epilogue), &od),
              "axb_elementwise_softmax");
        *result = std::move(r);
        return true;
    }

    // layer_norm / rms_norm over the last axis of a float32 node whose elementwise producers become the kernel's prologue
    // (post-LN: layer_norm(x + sublayer(x))). weight / bias: float32 rows or empty tensors.
    bool fused_norm_core(const GraphNode *c, bool layer, int ax, const Tensor &weight, const Tensor &bias, float eps, Tensor *result) {
        const int nd = (int)c->output_shape.size();
        if (nd < 1 || c->output_dtype != DType::Float32 || !(ax == nd - 1 || ax == -1) || c->output_shape.back() < 2 ||
            c->output_shape.back() > 8192 || ShapeUtils::size(c->output_shape) == 0)
            return false;
        if (is_leaf(c) || value.count(c) || !is_fusable(c))
            return false; // nothing to fuse in front: the stand-alone kernels
        const size_t n = c->output_shape.back();
        auto row = [&](const Tensor &t, Tensor *g) {
            if (!t.storage() || t.size() == 0)
                return true;
            if (t.dtype() != DType::Float32 || t.size() != n)
                return false;
            *g = backends::cuda::gpu_contiguous(on_gpu(t));
            return true;
        };
        Tensor w, b;
        if (!row(weight, &w) || !row(bias, &b))
            return false;
        Program pg = build_program(c);
        Tensor r = empty_gpu(c->output_shape, DType::Float32);
        std::vector<axb_tensor> ins(pg.inputs.size());
        for (size_t i = 0; i < pg.inputs.size(); ++i)
            ins[i] = describe(get(pg.inputs[i]));
        const axb_tensor od = describe(r);
        axb_tensor wd{}, bd{};
        if (w.storage())
            wd = describe(w);
        if (b.storage())
            bd = describe(b);
        check(plan::elementwise_norm(pg.instrs.data(), (int)pg.instrs.size(), ins.data(), (int)ins.size(),
                                     layer ? AXB_OP_LAYER_NORM : AXB_OP_RMS_NORM, w.storage() ? &wd : nullptr, b.storage() ? &bd : nullptr, eps,
                                     &od),
              "axb_elementwise_norm");
        *result = std::move(r);
        return true;
    }

    void run_program(const Program &pg, const axb_tensor &out, int32_t reduce_op /* <0: none */) {
        std::vector<axb_tensor> ins(pg.inputs.size());
        for (size_t i = 0; i < pg.inputs.size(); ++i)
            ins[i] = describe(get(pg.inputs[i]));
        if (reduce_op < 0)
            check(plan::elementwise(pg.instrs.data(), (int)pg.instrs.size(), ins.data(), (int)ins.size(), &out),
                  "axb_elementwise");
        else
            check(plan::elementwise_reduce(pg.instrs.data(), (int)pg.instrs.size(), ins.data(), (int)ins.size(), reduce_op, &out),
                  "axb_elementwise_reduce");
    }

    void eval_elementwise(const GraphNode *n) {
        Program pg = build_program(n);
        Tensor out = empty_gpu(n->output_shape, n->output_dtype);
        if (out.size() != 0)
            run_program(pg, describe(out), -1);
        value[n] = std::move(out);
    }

    // ---- everything else ----------------------------------------------------------------------------------------
    Tensor input(const GraphNode *n, size_t i) {
        if (i >= n->inputs.size() || !n->inputs[i] || is_absent(n->inputs[i].get()))
            return Tensor();
        const GraphNode *c = n->inputs[i].get();
        if (!is_leaf(c) && !value.count(c))
            eval(c);
        return get(c);
    }

    void eval(const GraphNode *n) {
        if (value.count(n) || is_leaf(n))
            return;
        const OpType op = n->op_type;
        if (try_linear_epilogue(*this, n))
            return;
        if (is_fusable(n)) {
            eval_elementwise(n);
            return;
        }
        Tensor r;
        if (op == OpType::MatMul || op == OpType::BatchMatMul) {
            const auto &mp = get_params<MatMulParams>(n->params);
            r = backends::cuda::gpu_matmul(input(n, 0), input(n, 1), mp.transpose_a, mp.transpose_b, nullptr, OpType::_Count,
                                           const_operands(n));
        } else if (op == OpType::Softmax || op == OpType::LogSoftmax) {
            if (!try_fused_softmax(n, OpType::_Count, n->output_shape, &r)) // elementwise producers as a prologue
                r = backends::cuda::gpu_softmax(input(n, 0), get_params<ActivationParams>(n->params).axis, op == OpType::LogSoftmax);
        } else if (op == OpType::LayerNorm) {
            const auto &np = get_params<NormParams>(n->params);
            const GraphNode *c0 = n->inputs.at(0).get();
            const Tensor w0 = n->inputs.size() > 1 && n->inputs[1] ? input(n, 1) : Tensor(),
                         b0 = n->inputs.size() > 2 && n->inputs[2] ? input(n, 2) : Tensor();
            if (!(n->output_dtype == DType::Float32 && c0 && c0->ref_count <= 1 && fused_norm_core(c0, true, np.axis, w0, b0, np.eps, &r)))
                r = backends::metal::gpu_layer_norm(input(n, 0), w0, b0, np.axis, np.eps);
        } else if (op == OpType::Conv1D) {
            const auto &cp = get_params<ConvParams>(n->params);
            r = backends::cuda::gpu_conv1d_impl(input(n, 0), input(n, 1), input(n, 2), cp.stride.at(0), cp.padding.at(0),
                                                cp.dilation.at(0), cp.groups);
        } else if (op == OpType::Conv2D) {
            const auto &cp = get_params<ConvParams>(n->params);
            r = backends::metal::gpu_conv2d(input(n, 0), input(n, 1), input(n, 2), {cp.stride.at(0), cp.stride.at(1)},
                                            {cp.padding.at(0), cp.padding.at(1)}, {cp.dilation.at(0), cp.dilation.at(1)}, cp.groups);
        } else if (op == OpType::BatchNorm1D) {
            // inference batch norm over the channel axis 1 (nn/normalization.cpp:88-106, lazy node graph_registry.cpp:637-655):
            // (x - running_mean) * 1/sqrt(running_var + eps) * weight + bias as ONE elementwise program in fp32
            Tensor x = input(n, 0), w = input(n, 1), b = input(n, 2), mean = input(n, 3), var = input(n, 4);
            if (x.ndim() < 2 || !mean.storage() || !var.storage())
                throw ShapeError("batch_norm: expected (N, C, ...) input and running statistics");
            Shape ps(x.ndim(), 1);
            ps[1] = x.shape()[1];
            auto chan = [&](const Tensor &t) { return backends::cuda::gpu_contiguous(t).reshape(ps); };
            std::vector<Tensor> ops_t{x, chan(mean), chan(var)};
            const int32_t F = AXB_F32;
            auto mk = [&](int32_t opc, int32_t s0, int32_t s1, double imm = 0.0) {
                axb_instr in{};
                in.op = opc; in.dtype = F; in.src[0] = s0; in.src[1] = s1; in.src[2] = AXB_SRC_NONE; in.imm = imm;
                return in;
            };
            std::vector<axb_instr> pr;
            pr.push_back(mk(AXB_OP_ADD, 2, AXB_SRC_IMM, (double)get_params<NormParams>(n->params).eps)); // var + eps
            pr.push_back(mk(AXB_OP_SQRT, AXB_SRC_VALUE(0), AXB_SRC_NONE));
            pr.push_back(mk(AXB_OP_RECIPROCAL, AXB_SRC_VALUE(1), AXB_SRC_NONE));
            pr.push_back(mk(AXB_OP_SUB, 0, 1));                                                            // x - mean
            pr.push_back(mk(AXB_OP_MUL, AXB_SRC_VALUE(3), AXB_SRC_VALUE(2)));
            if (w.storage() && w.size() > 0) {
                ops_t.push_back(chan(w));
                pr.push_back(mk(AXB_OP_MUL, AXB_SRC_VALUE((int32_t)pr.size() - 1), (int32_t)ops_t.size() - 1));
            }
            if (b.storage() && b.size() > 0) {
                ops_t.push_back(chan(b));
                pr.push_back(mk(AXB_OP_ADD, AXB_SRC_VALUE((int32_t)pr.size() - 1), (int32_t)ops_t.size() - 1));
            }
            if (n->output_dtype != DType::Float32) {
                axb_instr c = mk(AXB_OP_CAST, AXB_SRC_VALUE((int32_t)pr.size() - 1), AXB_SRC_NONE);
                c.dtype = to_axb_dtype(n->output_dtype);
                pr.push_back(c);
            }
            r = empty_gpu(n->output_shape, n->output_dtype);
            if (r.size() != 0) {
                std::vector<axb_tensor> ins;
                for (auto &t : ops_t)
                    ins.push_back(describe(t));
                const axb_tensor od = describe(r);
                check(plan::elementwise(pr.data(), (int)pr.size(), ins.data(), (int)ins.size(), &od), "axb_elementwise (batch_norm)");
            }
        } else if (op == OpType::Reshape) {
            Tensor x = input(n, 0);
            const Shape &ns = get_params<ReshapeParams>(n->params).new_shape;
            r = (x.is_contiguous() ? x : backends::cuda::gpu_contiguous(x)).reshape(ns); // a view of contiguous storage
        } else if (op == OpType::Transpose) {
            r = input(n, 0).transpose(get_params<TransposeParams>(n->params).axes); // permuted strides, no kernel
        } else if (op == OpType::Slice) {
            const auto &sp = get_params<SliceParams>(n->params);
            std::vector<Slice> sl;
            for (size_t d = 0; d < sp.starts.size(); ++d)
                sl.emplace_back(sp.starts[d], sp.ends[d], sp.strides[d]);
            r = input(n, 0).slice(sl); // strided view, no kernel
        } else if (op == OpType::Pad) {
            const auto &pp = get_params<PadParams>(n->params);
            r = backends::cuda::gpu_pad_constant(input(n, 0), pp.pad_widths, pp.value);
        } else if (op == OpType::GLU) { // a * sigmoid(b), halves along `axis` (operations.cpp:1452-1480)
            Tensor x = input(n, 0);
            int ax = get_params<ActivationParams>(n->params).axis;
            if (ax < 0)
                ax += (int)x.ndim();
            const int64_t half = (int64_t)x.shape()[ax] / 2;
            std::vector<Slice> lo(x.ndim()), hi(x.ndim());
            lo[ax] = Slice(0, half);
            hi[ax] = Slice(half, 2 * half);
            Tensor a = x.slice(lo), b = x.slice(hi);
            axb_instr pr[2] = {};
            pr[0].op = AXB_OP_SIGMOID; pr[0].dtype = to_axb_dtype(x.dtype());
            pr[0].src[0] = 1; pr[0].src[1] = pr[0].src[2] = AXB_SRC_NONE;
            pr[1].op = AXB_OP_MUL; pr[1].dtype = to_axb_dtype(x.dtype());
            pr[1].src[0] = 0; pr[1].src[1] = AXB_SRC_VALUE(0); pr[1].src[2] = AXB_SRC_NONE;
            r = empty_gpu(n->output_shape, n->output_dtype);
            const axb_tensor ins[2] = {describe(a), describe(b)}, od = describe(r);
            if (r.size() != 0)
                check(plan::elementwise(pr, 2, ins, 2, &od), "axb_elementwise (glu)");
        } else if (is_reduction_op(op)) {
            const auto &rp = get_params<ReductionParams>(n->params);
            const GraphNode *c = n->inputs.at(0).get();
            const bool full = rp.axes.empty() || rp.axes.size() == c->output_shape.size();
            const bool plain = op == OpType::Sum || op == OpType::Max || op == OpType::Min || op == OpType::Mean || op == OpType::Prod;
            if (full && plain && !rp.keep_dims && !is_leaf(c) && !value.count(c) && is_fusable(c) &&
                n->output_dtype == c->output_dtype && ShapeUtils::size(c->output_shape) > 0) {
                // elementwise tree + full reduction in one pass over the operands
                Program pg = build_program(c);
                r = empty_gpu(n->output_shape, n->output_dtype);
                run_program(pg, describe(r), to_axb_op(op));
            } else if (!full && rp.axes.size() == 1 && (op == OpType::Sum || op == OpType::Max || op == OpType::Mean) && !is_leaf(c) &&
                       !value.count(c) && c->op_type == OpType::Softmax && c->ref_count <= 1 && n->output_dtype == DType::Float32 &&
                       (rp.axes[0] == (int)c->output_shape.size() - 1 || rp.axes[0] == -1) &&
                       try_fused_softmax(c, op, n->output_shape, &r)) {
                // softmax(...).sum / mean / max over the softmax axis (the README chain): one kernel, the probabilities
                // never reach HBM
            } else if (!full && rp.axes.size() == 1 && (op == OpType::Sum || op == OpType::Max || op == OpType::Min || op == OpType::Mean) &&
                       !is_leaf(c) && !value.count(c) && is_fusable(c) && c->output_dtype == DType::Float32 &&
                       n->output_dtype == DType::Float32 && ShapeUtils::size(c->output_shape) > 0 &&
                       (rp.axes[0] == (int)c->output_shape.size() - 1 || rp.axes[0] == -1)) {
                // elementwise tree + reduction over its LAST axis in one pass (graph_compiler.cpp:376-478 fuses producers
                // into any reduction): one CTA / warp per output row folds the chain's values as it computes them
                Program pg = build_program(c);
                r = empty_gpu(n->output_shape, n->output_dtype);
                std::vector<axb_tensor> ins(pg.inputs.size());
                for (size_t i = 0; i < pg.inputs.size(); ++i)
                    ins[i] = describe(get(pg.inputs[i]));
                const axb_tensor od = describe(r);
                check(plan::elementwise_reduce_axis(pg.instrs.data(), (int)pg.instrs.size(), ins.data(), (int)ins.size(), to_axb_op(op),
                                                    (int)c->output_shape.size() - 1, rp.keep_dims ? 1 : 0, &od),
                      "axb_elementwise_reduce_axis");
            } else {
                r = backends::cuda::gpu_reduce(op, input(n, 0), rp.axes, rp.keep_dims);
            }
        } else {
            throw RuntimeError("CUDA graph lowering: unsupported op type " + std::to_string(static_cast<int>(op)) +
                               " (no CPU fallback on the GPU path)");
        }
        if (r.dtype() != n->output_dtype) // e.g. Mean of a half tensor is declared Float32 (graph_registry.cpp:278-287)
            r = backends::cuda::gpu_cast(r, n->output_dtype);
        if (r.shape() != n->output_shape)
            r = r.reshape(n->output_shape);
        value[n] = std::move(r);
    }
};

// MatMul -> (+ bias) -> (activation): recognised on the way DOWN from the consumer so the GEMM epilogue does the work.
// Returns true and stores the value of `n` when the pattern applied.
bool try_linear_epilogue(Lowering &L, const GraphNode *n) {
    auto fusable_act = [](OpType t) {
        return t == OpType::ReLU || t == OpType::GELU || t == OpType::SiLU || t == OpType::Sigmoid || t == OpType::Tanh;
    };
    const GraphNode *add = n;
    OpType act = OpType::_Count;
    if (fusable_act(n->op_type) && n->inputs.size() == 1 && n->inputs[0] && !is_leaf(n->inputs[0].get()) &&
        !L.value.count(n->inputs[0].get())) {
        act = n->op_type;
        add = n->inputs[0].get();
    }
    if (act != OpType::_Count && (add->op_type == OpType::Conv1D || add->op_type == OpType::Conv2D) &&
        add->output_dtype == n->output_dtype && add->ref_count <= 1) {
        // act(conv(x, w, b)): the convolution's GEMM epilogue applies bias and activation (whisper_poc.cpp:121-124 gelu(conv))
        const auto &cp = get_params<ConvParams>(add->params);
        const int nsp = add->op_type == OpType::Conv1D ? 1 : 2;
        int st[2] = {1, 1}, pd[2] = {0, 0}, dl[2] = {1, 1};
        for (int d = 0; d < nsp; ++d) {
            st[d] = cp.stride.at(d);
            pd[d] = cp.padding.at(d);
            dl[d] = cp.dilation.at(d);
        }
        Tensor r = backends::cuda::gpu_conv_nd(L.input(add, 0), L.input(add, 1), L.input(add, 2), nsp, st, pd, dl, cp.groups, act);
        if (r.shape() != n->output_shape)
            r = r.reshape(n->output_shape);
        L.value[n] = std::move(r);
        return true;
    }
    const GraphNode *mm = nullptr, *bias = nullptr;
    if (add->op_type == OpType::Add && add->inputs.size() == 2) {
        for (int s = 0; s < 2; ++s) {
            const GraphNode *a = add->inputs[s].get(), *b = add->inputs[1 - s].get();
            if (a && b && !is_leaf(a) && !L.value.count(a) && (a->op_type == OpType::MatMul || a->op_type == OpType::BatchMatMul) &&
                is_leaf(b) && b->output_shape.size() == 1 && a->output_shape.size() >= 2 &&
                b->output_shape[0] == a->output_shape.back() && b->output_dtype == a->output_dtype) {
                mm = a;
                bias = b;
                break;
            }
        }
    } else if (act != OpType::_Count && (add->op_type == OpType::MatMul || add->op_type == OpType::BatchMatMul)) {
        mm = add; // MatMul -> activation without bias
    }
    if (!mm || add->output_dtype != mm->output_dtype || n->output_dtype != mm->output_dtype)
        return false;
    if (mm->output_dtype != DType::Float32 && mm->output_dtype != DType::BFloat16 && mm->output_dtype != DType::Float16)
        return false;
    const auto &mp = get_params<MatMulParams>(mm->params);
    Tensor a = L.input(mm, 0), b = L.input(mm, 1);
    Tensor bias_t = bias ? L.get(bias) : Tensor();
    Tensor r = backends::cuda::gpu_matmul(a, b, mp.transpose_a, mp.transpose_b, bias ? &bias_t : nullptr, act, const_operands(mm));
    if (r.shape() != n->output_shape)
        r = r.reshape(n->output_shape);
    L.value[n] = std::move(r);
    return true;
}

// Materialises every not-yet-materialised node below `root` that has more than one consumer (contract step 1 of
// mpsgraph_graph_compiler.mm:709-744); afterwards the lazy nodes under `root` form a tree.
void materialize_shared_below(GraphNode *root) {
    std::vector<GraphNode *> shared_topo;
    std::unordered_set<GraphNode *> visited;
    std::vector<std::pair<GraphNode *, size_t>> stk{{root, 0}};
    while (!stk.empty()) {
        auto &[n, idx] = stk.back();
        if (visited.count(n)) {
            stk.pop_back();
            continue;
        }
        if (idx < n->inputs.size()) {
            GraphNode *c = n->inputs[idx++].get();
            if (c && !visited.count(c) && !is_leaf(c))
                stk.push_back({c, 0});
        } else {
            visited.insert(n);
            if (n != root && !is_leaf(n) && n->ref_count > 1)
                shared_topo.push_back(n);
            stk.pop_back();
        }
    }
    for (GraphNode *s : shared_topo)
        if (!s->is_materialized_)
            materialize_gpu_graph(s);
}

} // namespace

void materialize_gpu_graph(GraphNode *root) {
    if (root->is_materialized_)
        return;
    backends::cuda::ensure_initialized();
    materialize_shared_below(root); // 1. shared sub-DAGs first: their results stay available to other consumers
    // 2. plan cache (cuda_plan.hpp): signature + leaf layout -> recorded C-ABI call sequence. A hit replays it with the
    //    current leaf pointers; a miss lowers the remaining tree with the recorder on.
    Tensor out;
    if (!plan::try_replay(root, &out)) {
        plan::begin_record(root);
        try {
            Lowering L;
            L.eval(root);
            out = L.get(root);
            // 3. the node caches (storage, shape, strides) with offset 0 (graph_registry.cpp:323-331): views that start
            //    inside another tensor's storage are gathered into a fresh one
            if (out.offset() != 0) {
                Tensor c = empty_gpu(out.shape(), out.dtype());
                if (c.size() != 0) {
                    const axb_tensor sd = describe(out), dd = describe(c);
                    check(plan::gather_strided(&sd, &dd), "axb_gather_strided");
                }
                out = c;
            }
        } catch (...) {
            plan::end_record(nullptr, false);
            throw;
        }
        plan::end_record(&out, true);
    }
    root->cached_result_ = out.storage();
    root->cached_shape_ = out.shape();
    root->cached_strides_ = out.strides();
    root->is_materialized_ = true;
    for (auto &inp : root->inputs)
        if (inp && !inp->is_constant && inp->ref_count > 0)
            inp->ref_count--;
}

// softmax / log_softmax that arrive EAGERLY with a lazy operand (ops::log_softmax is eager on every device,
// operations.cpp:1437-1449): the operand's elementwise tree is lowered into the same kernel as the normalisation. Returns an
// empty Tensor when the pattern does not apply (the caller then materialises the operand and runs the stand-alone kernel).
Tensor fused_softmax(const Tensor &input, int axis, bool is_log) {
    auto ln = input.lazy_node();
    if (!ln || ln->is_materialized_ || ln->output_dtype != DType::Float32)
        return Tensor();
    backends::cuda::ensure_initialized();
    materialize_shared_below(ln.get());
    if (ln->ref_count > 1 || !is_fusable(ln.get()))
        return Tensor(); // someone else consumes the value as well, or there is nothing elementwise to fuse
    Lowering L;
    Tensor r;
    if (!L.fused_softmax_core(ln.get(), axis, is_log, OpType::_Count, ln->output_shape, &r))
        return Tensor();
    return r;
}

// rms_norm arrives eagerly through the gpu_rms_norm seam (operations.cpp:1834) with a possibly lazy operand
Tensor fused_rms_norm(const Tensor &input, const Tensor &weight, int axis, float eps) {
    auto ln = input.lazy_node();
    if (!ln || ln->is_materialized_ || ln->output_dtype != DType::Float32)
        return Tensor();
    backends::cuda::ensure_initialized();
    materialize_shared_below(ln.get());
    if (ln->ref_count > 1 || !is_fusable(ln.get()))
        return Tensor();
    Lowering L;
    Tensor r;
    if (!L.fused_norm_core(ln.get(), false, axis, weight, Tensor(), eps, &r))
        return Tensor();
    return r;
}

// where(cond, a, b) is eager in the reference (operations.cpp:1548-1570), which would cut `chain.where(c > 0, 0)` into
// three launches (chain, comparison, select). When the operands are still lazy, their elementwise trees are lowered
// into the SAME program as the select: BASELINE config 2 becomes one kernel through the unchanged Tensor API.
Tensor fused_where(const Tensor &cond, const Tensor &a, const Tensor &b) {
    backends::cuda::ensure_initialized();
    const Tensor *ops3[3] = {&cond, &a, &b};
    std::vector<std::shared_ptr<GraphNode>> keep; // wrappers for operands that are already materialised
    const GraphNode *nodes[3];
    for (int i = 0; i < 3; ++i) {
        auto ln = ops3[i]->lazy_node();
        if (ln && !ln->is_materialized_) {
            materialize_shared_below(ln.get());
            if (ln->ref_count > 1) // someone else consumes this value as well: compute it once, share the result
                materialize_gpu_graph(ln.get());
            nodes[i] = ln.get();
        } else {
            const Tensor t = on_gpu(*ops3[i]);
            auto c = std::make_shared<GraphNode>();
            c->is_constant = true;
            c->is_materialized_ = true;
            c->constant_storage = t.storage();
            c->constant_strides = t.strides();
            c->constant_offset = t.offset();
            c->output_shape = t.shape();
            c->output_dtype = t.dtype();
            c->target_device = Device::GPU;
            keep.push_back(c);
            nodes[i] = c.get();
        }
    }
    const int32_t rdt = axb_promote_types(to_axb_dtype(nodes[1]->output_dtype), to_axb_dtype(nodes[2]->output_dtype));
    if (rdt < 0)
        throw TypeError("where: operands cannot be promoted to a common dtype");
    int64_t sa[3][AXB_MAX_DIMS], so[AXB_MAX_DIMS], tmp[AXB_MAX_DIMS];
    int n_out = 0, n_tmp = 0;
    for (int i = 0; i < 3; ++i) {
        if (nodes[i]->output_shape.size() > AXB_MAX_DIMS)
            throw ShapeError("CUDA backend supports at most 8 dimensions");
        for (size_t d = 0; d < nodes[i]->output_shape.size(); ++d)
            sa[i][d] = (int64_t)nodes[i]->output_shape[d];
    }
    if (axb_broadcast_shape(sa[1], (int)nodes[1]->output_shape.size(), sa[2], (int)nodes[2]->output_shape.size(), tmp, &n_tmp) != AXB_OK ||
        axb_broadcast_shape(sa[0], (int)nodes[0]->output_shape.size(), tmp, n_tmp, so, &n_out) != AXB_OK)
        throw ShapeError::broadcast_incompatible(axb_last_error());
    Shape oshape(n_out);
    for (int d = 0; d < n_out; ++d)
        oshape[d] = (size_t)so[d];

    Lowering L;
    Lowering::Program pg;
    int32_t sel[3];
    for (int attempt = 0;; ++attempt) {
        pg = Lowering::Program();
        try {
            // a synthetic root that is never the tree_root of emit(): every operand may be inlined
            for (int i = 0; i < 3; ++i)
                sel[i] = L.emit(nodes[i], pg, nullptr);
            if (pg.instrs.size() + 1 > AXB_MAX_INSTRS)
                throw std::length_error("instructions");
            break;
        } catch (const std::length_error &) {
            // too many operands / instructions for one program: evaluate the largest lazy operand on its own
            const GraphNode *victim = nullptr;
            for (int i = 0; i < 3; ++i)
                if (!is_leaf(nodes[i]) && !L.value.count(nodes[i]))
                    victim = nodes[i];
            if (!victim || attempt > 3)
                throw RuntimeError::internal("CUDA where: operands do not fit one elementwise program");
            L.eval(victim);
        }
    }
    axb_instr w{};
    w.op = AXB_OP_WHERE;
    w.dtype = rdt;
    w.src[0] = sel[0];
    w.src[1] = sel[1];
    w.src[2] = sel[2];
    pg.instrs.push_back(w);
    Tensor out = empty_gpu(oshape, static_cast<DType>(rdt));
    if (out.size() != 0)
        L.run_program(pg, describe(out), -1);
    return out;
}

// ---- hooks of the compile-and-execute path (graph_executor.cpp:439-443,634-638,830-834) ----------------------------
namespace {

// Lowers op_chain with the slot convention of mpsgraph_fused.mm:424-470: input_slot_indices[i] lists op i's operands as
// indices into `buffers`, -1 = result of the previous op.
bool chain_program(const StepBase &step, const std::vector<OpType> &chain, std::vector<Tensor> &buffers, DType chain_dtype,
                   std::vector<axb_instr> &prog, std::vector<Tensor> &inputs) {
    if (chain.empty() || chain.size() > AXB_MAX_INSTRS || step.input_slot_indices.size() < chain.size())
        return false;
    std::vector<int> slot_of_input;
    for (size_t i = 0; i < chain.size(); ++i) {
        const OpType op = chain[i];
        if (!(is_unary_op(op) || is_binary_op(op)) || !is_elementwise_op(op))
            return false;
        axb_instr in{};
        in.op = to_axb_op(op);
        in.src[0] = in.src[1] = in.src[2] = AXB_SRC_NONE;
        if (op == OpType::LeakyReLU)
            in.imm = 0.01; // op chains carry no parameters; same default as the eager path and the CPU functor
        const auto &slots = step.input_slot_indices[i];
        int32_t dts[2] = {-1, -1};
        for (size_t k = 0; k < slots.size() && k < 2; ++k) {
            if (slots[k] < 0) {
                if (i == 0)
                    return false;
                in.src[k] = AXB_SRC_VALUE((int32_t)i - 1);
                dts[k] = prog[i - 1].dtype;
            } else {
                if (slots[k] >= (int)buffers.size())
                    return false;
                auto it = std::find(slot_of_input.begin(), slot_of_input.end(), slots[k]);
                if (it == slot_of_input.end()) {
                    if (inputs.size() >= AXB_MAX_INPUTS)
                        return false;
                    slot_of_input.push_back(slots[k]);
                    inputs.push_back(on_gpu(buffers[slots[k]]));
                    it = slot_of_input.end() - 1;
                }
                in.src[k] = (int32_t)(it - slot_of_input.begin());
                dts[k] = to_axb_dtype(inputs[in.src[k]].dtype());
            }
        }
        in.dtype = is_binary_op(op) ? axb_binary_result_dtype(in.op, dts[0], dts[1]) : axb_unary_result_dtype(in.op, dts[0]);
        if (in.dtype < 0)
            return false;
        prog.push_back(in);
    }
    (void)chain_dtype;
    return true;
}

} // namespace

bool execute_gpu_fused_chain(const StepBase &step, const std::vector<ops::OpType> &op_chain, std::vector<Tensor> &buffers) {
    std::vector<axb_instr> prog;
    std::vector<Tensor> inputs;
    if (step.output_slot < 0 || step.output_slot >= (int)buffers.size() ||
        !chain_program(step, op_chain, buffers, step.output_dtype, prog, inputs))
        return false;
    if (prog.back().dtype != to_axb_dtype(step.output_dtype)) {
        if (prog.size() >= AXB_MAX_INSTRS)
            return false;
        axb_instr cast{};
        cast.op = AXB_OP_CAST;
        cast.dtype = to_axb_dtype(step.output_dtype);
        cast.src[0] = AXB_SRC_VALUE((int32_t)prog.size() - 1);
        cast.src[1] = cast.src[2] = AXB_SRC_NONE;
        prog.push_back(cast);
    }
    Tensor out = empty_gpu(step.output_shape, step.output_dtype); // fresh result replaces the pre-bound arena view
    if (out.size() != 0) {
        std::vector<axb_tensor> ins(inputs.size());
        for (size_t i = 0; i < inputs.size(); ++i)
            ins[i] = describe(inputs[i]);
        const axb_tensor od = describe(out);
        check(plan::elementwise(prog.data(), (int)prog.size(), ins.data(), (int)ins.size(), &od), "axb_elementwise");
    }
    buffers[step.output_slot] = out;
    return true;
}

bool execute_gpu_fused_reduction(const FusedReductionStep &step, std::vector<Tensor> &buffers) {
    const OpType rop = step.reduction_op;
    if (!(rop == OpType::Sum || rop == OpType::Max || rop == OpType::Min || rop == OpType::Mean || rop == OpType::Prod))
        return false;
    if (auto *rp = std::get_if<ReductionParams>(&step.params))
        if (rp->keep_dims)
            return false;
    std::vector<axb_instr> prog;
    std::vector<Tensor> inputs;
    if (step.output_slot < 0 || step.output_slot >= (int)buffers.size() ||
        !chain_program(step, step.op_chain, buffers, step.chain_dtype, prog, inputs))
        return false;
    if (ShapeUtils::size(step.output_shape) != 1 || prog.back().dtype != to_axb_dtype(step.output_dtype))
        return false; // only the chain + FULL reduction pattern is fused (graph_compiler.cpp fusion rule)
    Tensor out = empty_gpu(step.output_shape, step.output_dtype);
    std::vector<axb_tensor> ins(inputs.size());
    for (size_t i = 0; i < inputs.size(); ++i)
        ins[i] = describe(inputs[i]);
    const axb_tensor od = describe(out);
    check(plan::elementwise_reduce(prog.data(), (int)prog.size(), ins.data(), (int)ins.size(), to_axb_op(rop), &od),
          "axb_elementwise_reduce");
    buffers[step.output_slot] = out;
    return true;
}

} // namespace graph
} // namespace axiom
