// cuda_backend.hpp — host-side glue between the UNMODIFIED Axiom C++ library and the axb C ABI (libaxb.so).
//
// This directory is the `src/backends/cuda` that SURVEY.md §8(b) describes: it is written in the reference's own
// language (C++20, compiled against the reference headers where they lie) and fills the three in-process seams:
//   1. axiom::Storage / make_storage            (include/axiom/storage.hpp:13-32, src/tensor/storage.cpp:26-48)
//   2. ops::Operation + OperationRegistry        (include/axiom/operations.hpp:210-302)
//   3. graph lowering hooks                      (src/graph/graph_registry.cpp:16-24,834-850,
//                                                 include/axiom/graph/gpu_fusion.hpp:14-22)
// plus the free functions the shared code calls on its GPU paths (src/backends/metal/metal_operations.hpp:11-67,
// metal_common.hpp:53-98, unified_storage.hpp:16-58). Those symbols live in namespace axiom::backends::metal because
// that is the name the reference's call sites use (tensor.cpp:587,1604-1648; operations.cpp:1837,4033); here they are
// implemented on CUDA. Nothing in this directory computes on the CPU: every tensor operation ends in an axb_* call.
#pragma once

#include <map>

#include <string>

#include "axb.h"
#include "axiom/error.hpp"
#include "axiom/operations.hpp"
#include "axiom/storage.hpp"
#include "axiom/tensor.hpp"

namespace axiom {
namespace backends {
namespace cuda {

// Device memory owned through the stream-ordered caching allocator of libaxb (axb_alloc / axb_free).
// Like MetalStorage (src/backends/metal/metal_storage.mm:54-60) data() throws: shared code only touches data() after
// checking device() == CPU; device access goes through device_ptr().
class CudaStorage : public Storage {
  public:
    explicit CudaStorage(size_t size_bytes);
    // takes ownership of a block that came from axb_alloc on the current device (plan replay, cuda_plan.cpp)
    struct Adopt {};
    CudaStorage(void *adopted, size_t size_bytes, Adopt);
    ~CudaStorage() override;
    CudaStorage(const CudaStorage &) = delete;
    CudaStorage &operator=(const CudaStorage &) = delete;

    void *data() override;
    const void *data() const override;
    size_t size_bytes() const override { return size_bytes_; }
    Device device() const override { return Device::GPU; }
    void copy_to(Storage &other) const override;
    void copy_from(const Storage &other) override;
    std::unique_ptr<Storage> clone() const override;

    void *device_ptr() const { return ptr_; }
    int device_index() const { return device_index_; }

  private:
    void *ptr_ = nullptr;
    size_t size_bytes_ = 0;
    int device_index_ = 0;
};

// axb status -> Axiom exception (include/axiom/error.hpp); never returns on failure
void check(axb_status st, const char *what);

// Binds the calling thread to the backend's device, and registers the CUDA operations on first use
void ensure_initialized();
void register_cuda_operations();
bool is_cuda_available();

// Descriptor of a materialised GPU tensor (shape, BYTE strides, byte offset — zero / negative strides pass through)
axb_tensor describe(const Tensor &t);
// Fresh contiguous GPU tensor
Tensor empty_gpu(const Shape &shape, DType dtype);
// Moves a tensor to the GPU if it is not there yet (materialises lazy tensors)
Tensor on_gpu(const Tensor &t);

inline std::string op_name(ops::OpType op) { return "op#" + std::to_string(static_cast<int>(op)); }
inline int32_t to_axb_dtype(DType d) { return static_cast<int32_t>(d); }   // same order (dtype.hpp:171-191)
inline int32_t to_axb_op(ops::OpType op) { return static_cast<int32_t>(op); } // same order (operations.hpp:49-209)

// Functional entry points shared by the Operation classes, the graph lowering and the seam functions
Tensor gpu_unary(ops::OpType op, const Tensor &x, double param = 0.0);
Tensor gpu_binary(ops::OpType op, const Tensor &a, const Tensor &b);
Tensor gpu_reduce(ops::OpType op, const Tensor &x, const std::vector<int> &axes, bool keep_dims);
// const_ops: AXB_MM_CONST_A / AXB_MM_CONST_B — the operand is a constant leaf of the graph (weights): libaxb may keep
// derived forms of it (the fp32 three-piece split) until its storage is freed or written
Tensor gpu_matmul(const Tensor &a, const Tensor &b, bool ta, bool tb, const Tensor *bias = nullptr,
                  ops::OpType activation = ops::OpType::_Count, int const_ops = 0);
Tensor gpu_softmax(const Tensor &x, int axis, bool is_log);
Tensor gpu_where(const Tensor &c, const Tensor &a, const Tensor &b);
Tensor gpu_masked_fill(const Tensor &x, const Tensor &mask, double value);
Tensor gpu_cast(const Tensor &x, DType to);
Tensor gpu_contiguous(const Tensor &x);
Tensor gpu_compact(const Tensor &x); // always a fresh, offset-free storage (patched Tensor::to)
Tensor gpu_conv1d_impl(const Tensor &x, const Tensor &w, const Tensor &b, int stride, int padding, int dilation,
                       int groups);
Tensor gpu_conv_nd(const Tensor &x, const Tensor &w, const Tensor &b, int nsp, const int *stride, const int *padding,
                   const int *dilation, int groups, ops::OpType activation);
Tensor gpu_pad_constant(const Tensor &x, const std::vector<std::pair<size_t, size_t>> &pad, double value);
// Entry points of the patched ops::max_pool*/avg_pool*/sort/argsort/topk (patches/apply.py; the reference has no device
// branch for them, operations.cpp:2367-2880, :4199-4257). `out_shape` is the shape the reference computed (without the
// batch dim when the input has none). Results live on the GPU.
Tensor gpu_pool(const Tensor &input, bool is_avg, const std::vector<int> &kernel, const std::vector<int> &stride,
                const std::vector<int> &padding, bool count_include_pad, const Shape &out_shape);
Tensor gpu_sort(const Tensor &input, int axis, bool descending, bool want_indices);
std::pair<Tensor, Tensor> gpu_topk(const Tensor &input, int k, int axis, bool largest);

// Weight loading (patched io::safetensors::load / load_tensor on Device::GPU, src/io/io_safetensors.cpp:123-214): the byte
// ranges of `filename` go file -> pinned staging -> device (axb_upload_file), in file order; no host tensor is created.
struct FileTensor {
    std::string name;
    DType dtype;
    Shape shape;
    size_t file_offset;
    size_t bytes;
};
std::map<std::string, Tensor> load_file_tensors(const std::string &filename, const std::vector<FileTensor> &items);

} // namespace cuda
} // namespace backends

namespace graph {
// cuda_graph.cpp: where() with lazy operands lowered into one elementwise program
Tensor fused_where(const Tensor &cond, const Tensor &a, const Tensor &b);
// cuda_graph.cpp: eager softmax / log_softmax of a lazy operand, its elementwise tree as the kernel's prologue (empty: n/a)
Tensor fused_softmax(const Tensor &input, int axis, bool is_log);
Tensor fused_rms_norm(const Tensor &input, const Tensor &weight, int axis, float eps);
} // namespace graph
} // namespace axiom
