// cuda_plan.hpp — graph-level plan cache of the CUDA lowering.
//
// Reference behaviour being matched: the Metal backend keys a compiled MPSGraph on compute_signature(root)
// (src/backends/metal/mpsgraph_graph_compiler.mm:756-765, 896-902; LRU cache src/graph/graph_cache.cpp:12-52) and replays
// it with fresh buffers. That signature deliberately ignores the strides of materialised leaves
// (src/graph/graph_signature.cpp:129-139), which is only sound for a backend that gathers every input first; this
// backend reads strided leaves in place, so the key here is   signature  +  layout class of every leaf   (device, dtype,
// shape, byte strides, offset, 16-byte alignment class, and which leaves alias one storage).
//
// What a plan is: the exact sequence of C-ABI calls (include/axb.h) the lowering issued for that graph — every tensor
// argument rewritten as (leaf k | temporary j) + byte offset + shape + strides — plus the temporaries' sizes and live
// ranges and the description of the result. A MISS runs the normal lowering (cuda_graph.cpp) with the recorder on; a HIT
// skips the tree walk, fusion decisions, Tensor bookkeeping and per-node allocation logic entirely and re-issues the calls
// with the current leaf pointers and freshly allocated temporaries (stream-ordered allocator: freed right after their
// last use). Host scalars captured as CPU leaves (`x + 1.0f`) are re-uploaded from the CURRENT leaf, so plans never
// depend on leaf data. Anything the recorder cannot express (an unwrapped kernel, a device->host copy, a nested
// materialisation, an upload from memory that is not a leaf) marks the signature as not cacheable: it keeps taking the
// normal path. Launch and transfer counters of libaxb cross-check that nothing escaped the recording.
#pragma once

#include <memory>
#include <vector>

#include "cuda_backend.hpp"

namespace axiom {
namespace graph {
struct GraphNode;
}
namespace backends {
namespace cuda {
namespace plan {

struct Stats {
    uint64_t hits = 0, misses = 0, uncacheable = 0, entries = 0, evictions = 0;
};
Stats stats();
void clear();
bool enabled();

// ---- hooks for cuda_storage.cpp -----------------------------------------------------------------------------------
bool recording();
void note_alloc(void *device_ptr, size_t bytes);                         // a CudaStorage was created
void note_upload(void *device_dst, const void *host_src, size_t bytes);  // CudaStorage::copy_from(CPU storage)
void note_foreign(const char *why);                                      // anything a plan cannot express

// ---- entry points for materialize_gpu_graph --------------------------------------------------------------------------
// true: `*out` holds the materialised result of `root` (plan hit, replayed)
bool try_replay(const graph::GraphNode *root, Tensor *out);
// after a miss: starts recording the lowering of `root` on this thread (no-op when the key is known as uncacheable)
void begin_record(const graph::GraphNode *root);
// closes the recording; `result` is what the lowering produced (ignored when `ok` is false, e.g. an exception)
void end_record(const Tensor *result, bool ok);

// ---- recording wrappers of the C-ABI calls the lowering uses ----------------------------------------------------------
axb_status elementwise(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, const axb_tensor *out);
axb_status elementwise_reduce(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t reduce_op,
                              const axb_tensor *out);
axb_status elementwise_reduce_axis(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t reduce_op,
                                   int axis, int keep_dims, const axb_tensor *out);
axb_status elementwise_softmax(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int is_log, int32_t epilogue_op,
                               const axb_tensor *out);
axb_status elementwise_norm(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t kind, const axb_tensor *w,
                            const axb_tensor *b, float eps, const axb_tensor *out);
axb_status matmul(const axb_tensor *a, const axb_tensor *b, int ta, int tb, const axb_matmul_epilogue *epi, const axb_tensor *out);
axb_status softmax(const axb_tensor *in, int axis, int is_log, const axb_tensor *out);
axb_status layer_norm(const axb_tensor *in, const axb_tensor *w, const axb_tensor *b, int axis, float eps, const axb_tensor *out);
axb_status rms_norm(const axb_tensor *in, const axb_tensor *w, int axis, float eps, const axb_tensor *out);
axb_status reduce(int32_t op, const axb_tensor *in, const int32_t *axes, int n_axes, int keep_dims, const axb_tensor *out);
axb_status argreduce(int32_t op, const axb_tensor *in, int axis, int keep_dims, const axb_tensor *out);
axb_status cast(const axb_tensor *in, const axb_tensor *out);
axb_status gather_strided(const axb_tensor *in, const axb_tensor *out);
axb_status im2col_1d(const axb_tensor *in, int k, int stride, int padding, int dilation, const axb_tensor *out);
axb_status binary(int32_t op, const axb_tensor *a, const axb_tensor *b, const axb_tensor *out);
axb_status unary(int32_t op, const axb_tensor *in, const axb_tensor *out, double param);
axb_status where(const axb_tensor *c, const axb_tensor *a, const axb_tensor *b, const axb_tensor *out);
axb_status masked_fill(const axb_tensor *in, const axb_tensor *mask, double value, const axb_tensor *out);
axb_status pad_constant(const axb_tensor *in, const int64_t *lo, const int64_t *hi, double value, const axb_tensor *out);
axb_status conv(const axb_tensor *in, const axb_tensor *w, const axb_tensor *b, int nsp, const int32_t *stride, const int32_t *padding,
                const int32_t *dilation, int groups, int activation, const axb_tensor *out);
axb_status fill(const axb_tensor *out, double value);
axb_status pool(const axb_tensor *in, int is_avg, int nsp, const int32_t *kernel, const int32_t *stride, const int32_t *padding,
                int count_include_pad, const axb_tensor *out);
axb_status sort(const axb_tensor *in, int axis, int descending, const axb_tensor *values, const axb_tensor *indices);

} // namespace plan
} // namespace cuda
} // namespace backends
} // namespace axiom
