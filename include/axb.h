/*
 * axb.h — the thin C ABI of the B200 (sm_100a) CUDA backend for Frikallo/axiom.
 *
 * This is the drop-in boundary for the reference's Device::GPU hot path. Axiom has no
 * FFI of its own: its GPU backend is reached through three in-process C++ seams
 *   (1) axiom::Storage / make_storage            include/axiom/storage.hpp:13-32, src/tensor/storage.cpp:26-48
 *   (2) ops::Operation + OperationRegistry       include/axiom/operations.hpp:210-302
 *   (3) graph lowering hooks                     include/axiom/graph/gpu_fusion.hpp:14-22,
 *                                                src/graph/graph_registry.cpp:16-24,834-850
 * The entry points below are exactly what a `src/backends/cuda` glue layer filling those seams binds
 * (INTEGRATION.md shows that glue). Everything is `extern "C"`, plain pointers and sizes: no C++ types,
 * no torch types, no exceptions across the boundary. Every call returns an axb_status; on failure
 * axb_last_error() holds a message (thread-local) and the glue converts it to the matching
 * axiom::*Error (include/axiom/error.hpp).
 *
 * Conventions
 *  - Tensors are described by axb_tensor: device base pointer + byte offset of element [0,..,0] +
 *    shape + BYTE strides (exactly Tensor's shape_/strides_, include/axiom/tensor.hpp:40-58; negative
 *    and zero strides allowed).  NOTE the reference keeps `offset_` at the LOWEST address of a view and
 *    adds the negative-stride adjustment in Tensor::data() (src/tensor/tensor.cpp:279-291); the glue
 *    passes the adjusted value so that base+byte_offset is element [0,..,0].
 *  - dtype / op codes are numerically identical to axiom::DType (include/axiom/dtype.hpp:170-192) and
 *    axiom::ops::OpType (include/axiom/operations.hpp:55-208), so the glue can static_cast.
 *  - All work is enqueued on a stream (axb_stream_t == cudaStream_t; NULL = the backend's per-device
 *    stream, the analogue of MetalExecutionStream, src/backends/metal/metal_common.hpp:53-123).
 *    Nothing synchronises except axb_stream_sync / axb_memcpy_d2h / the *_host convenience calls.
 *  - No CPU fallback: an unsupported dtype/op returns AXB_ERR_UNSUPPORTED, it never computes on host.
 */
#ifndef AXB_H
#define AXB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define AXB_MAX_DIMS 8
#define AXB_ABI_VERSION 2

typedef enum axb_status {
    AXB_OK = 0,
    AXB_ERR_INVALID = 1,     /* -> axiom::ValueError / ShapeError  */
    AXB_ERR_UNSUPPORTED = 2, /* -> axiom::TypeError / RuntimeError::not_implemented */
    AXB_ERR_CUDA = 3,        /* -> axiom::DeviceError */
    AXB_ERR_OOM = 4,         /* -> axiom::MemoryError */
    AXB_ERR_SHAPE = 5        /* -> axiom::ShapeError */
} axb_status;

/* == axiom::DType */
typedef enum axb_dtype {
    AXB_BOOL = 0, AXB_I8 = 1, AXB_I16 = 2, AXB_I32 = 3, AXB_I64 = 4,
    AXB_U8 = 5, AXB_U16 = 6, AXB_U32 = 7, AXB_U64 = 8,
    AXB_F16 = 9, AXB_BF16 = 10, AXB_F32 = 11, AXB_F64 = 12,
    AXB_C64 = 13, AXB_C128 = 14
} axb_dtype;

/* == axiom::ops::OpType (only the values this backend implements are named) */
typedef enum axb_op {
    AXB_OP_ADD = 0, AXB_OP_SUB = 1, AXB_OP_MUL = 2, AXB_OP_DIV = 3, AXB_OP_POW = 4, AXB_OP_MOD = 5,
    AXB_OP_EQ = 6, AXB_OP_NE = 7, AXB_OP_LT = 8, AXB_OP_LE = 9, AXB_OP_GT = 10, AXB_OP_GE = 11,
    AXB_OP_LOGICAL_AND = 12, AXB_OP_LOGICAL_OR = 13, AXB_OP_LOGICAL_XOR = 14, AXB_OP_LOGICAL_NOT = 15,
    AXB_OP_BIT_AND = 16, AXB_OP_BIT_OR = 17, AXB_OP_BIT_XOR = 18, AXB_OP_SHL = 19, AXB_OP_SHR = 20,
    AXB_OP_MAXIMUM = 21, AXB_OP_MINIMUM = 22, AXB_OP_ATAN2 = 23, AXB_OP_HYPOT = 24,
    AXB_OP_NEG = 25, AXB_OP_ABS = 26, AXB_OP_SQRT = 27, AXB_OP_EXP = 28, AXB_OP_LOG = 29,
    AXB_OP_SIN = 30, AXB_OP_COS = 31, AXB_OP_TAN = 32, AXB_OP_ERF = 33,
    AXB_OP_SIGN = 34, AXB_OP_FLOOR = 35, AXB_OP_CEIL = 36, AXB_OP_TRUNC = 37, AXB_OP_ROUND = 38,
    AXB_OP_RECIPROCAL = 39, AXB_OP_SQUARE = 40, AXB_OP_CBRT = 41,
    AXB_OP_ISNAN = 42, AXB_OP_ISINF = 43, AXB_OP_ISFINITE = 44,
    AXB_OP_RELU = 48, AXB_OP_LEAKY_RELU = 49, AXB_OP_SILU = 50, AXB_OP_SIGMOID = 51, AXB_OP_TANH = 52,
    AXB_OP_GELU = 53, AXB_OP_SOFTMAX = 54, AXB_OP_LOG_SOFTMAX = 55,
    AXB_OP_SUM = 56, AXB_OP_MEAN = 57, AXB_OP_MAX = 58, AXB_OP_MIN = 59,
    AXB_OP_ARGMAX = 60, AXB_OP_ARGMIN = 61, AXB_OP_ANY = 62, AXB_OP_ALL = 63, AXB_OP_PROD = 64,
    AXB_OP_MATMUL = 65, AXB_OP_BATCH_MATMUL = 66,
    AXB_OP_WHERE = 67, AXB_OP_MASKED_FILL = 68,
    AXB_OP_GATHER = 70, AXB_OP_INDEX_SELECT = 72,
    AXB_OP_LAYER_NORM = 75, AXB_OP_RMS_NORM = 76,
    AXB_OP_CAST = 78
} axb_op;

/* Strided device tensor view. base+byte_offset addresses element [0,...,0]. */
typedef struct axb_tensor {
    void *base;
    int64_t byte_offset;
    int32_t ndim;  /* 0..AXB_MAX_DIMS; ndim==0 is a rank-0 scalar (1 element) */
    int32_t dtype; /* axb_dtype */
    int64_t shape[AXB_MAX_DIMS];
    int64_t strides[AXB_MAX_DIMS]; /* BYTES, may be 0 (broadcast) or negative (flip) */
} axb_tensor;

typedef void *axb_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------------------------------------
 * Runtime: device / stream / errors.
 * Replaces: is_metal_available() src/tensor/system.cpp:12-29; MetalExecutionStream
 * src/backends/metal/metal_common.hpp:53-123 (single async stream, sync only when the host reads).
 * ---------------------------------------------------------------------------------------------- */
int axb_abi_version(void);
const char *axb_last_error(void);
axb_status axb_device_count(int *count);
axb_status axb_set_device(int device);           /* binds the calling thread; creates the device context */
axb_status axb_get_device(int *device);
axb_status axb_device_name(int device, char *buf, size_t buflen);
axb_status axb_device_sm_count(int device, int *sms);
axb_status axb_default_stream(axb_stream_t *stream); /* the backend stream of the current device */
axb_status axb_stream_sync(axb_stream_t stream);
/* number of kernels this library has launched since load (bench.py's gpu_launches evidence) */
uint64_t axb_launch_count(void);
/* device-side timing on the stream the kernels run on (cudaEvent_t wrappers; bench harness only) */
typedef void *axb_event_t;
axb_status axb_event_create(axb_event_t *ev);
axb_status axb_event_destroy(axb_event_t ev);
axb_status axb_event_record(axb_event_t ev, axb_stream_t stream);
axb_status axb_event_sync(axb_event_t ev);
axb_status axb_event_elapsed_ms(axb_event_t start, axb_event_t end, float *ms);

/* ------------------------------------------------------------------------------------------------
 * Storage: stream-ordered caching device allocator + transfers.
 * Replaces: make_storage(bytes, Device::GPU) src/tensor/storage.cpp:31-45 (throws on Linux today);
 * MetalStorage copy_to/copy_from src/backends/metal/metal_storage.mm:54-142;
 * Tensor::to/cpu/gpu src/tensor/tensor.cpp:1593-1650.
 * axb_free(ptr, stream): the block becomes reusable by later allocations on `stream` immediately
 * (stream order makes that safe) and by other streams only after an event recorded at free time.
 * size 0 is legal (returns NULL).
 * ---------------------------------------------------------------------------------------------- */
axb_status axb_alloc(size_t bytes, axb_stream_t stream, void **ptr);
axb_status axb_free(void *ptr, axb_stream_t stream);
axb_status axb_empty_cache(void);
typedef struct axb_alloc_stats {
    uint64_t bytes_in_use, bytes_cached, peak_bytes_in_use, n_cuda_malloc, n_cache_hits, n_allocs;
    uint64_t n_splits, n_merges, n_cross_stream; /* block splitting / coalescing / event-ordered hand-overs */
} axb_alloc_stats;
axb_status axb_get_alloc_stats(axb_alloc_stats *out);
/* pinned host staging (H2D of randn()/weights, tensor.cpp:1884-1904) */
axb_status axb_host_alloc(size_t bytes, void **ptr);
axb_status axb_host_free(void *ptr);
/* NUMA node the LAST axb_host_alloc of the calling thread was placed on (the node of the current GPU's PCI function), or
 * -1 when no placement was applied (single-node host, syscall unavailable, AXB_HOST_NUMA=0). */
int axb_host_alloc_numa_node(void);
/* File -> device upload for weight loading: `bytes` bytes of `path` starting at `file_offset` land at `dst` (device memory),
 * read with pread() straight into pinned staging buffers whose DMA overlaps the next read — no pageable host copy of the
 * data ever exists. Replaces "read into a host tensor, then Tensor::gpu()" of the reference's safetensors loader
 * (src/io/io_safetensors.cpp:123-160). The copies are ordered on `stream`; the call returns when the file has been read. */
axb_status axb_upload_file(const char *path, uint64_t file_offset, size_t bytes, void *dst, axb_stream_t stream);
/* async copies on `stream`; d2h additionally synchronises the stream before returning when sync!=0
 * (MetalStorage::copy_to(CPU) synchronises first, metal_storage.mm:66-93) */
axb_status axb_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, axb_stream_t stream);
axb_status axb_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, axb_stream_t stream, int sync);
axb_status axb_memcpy_d2d(void *dst_dev, const void *src_dev, size_t bytes, axb_stream_t stream);
axb_status axb_memset(void *dst_dev, int value, size_t bytes, axb_stream_t stream);
/* Host<->device traffic this library has issued since load (all devices, all entry points incl. the *_host calls).
 * The fallback detector of the C++ glue and the tests: a GPU forward that silently dropped to the reference's CPU
 * path (src/tensor/operations.cpp:913-917, src/graph/graph_executor.cpp:346-349) shows up as D2H bytes between two
 * snapshots, exactly as axb_launch_count shows missing kernels. */
typedef struct axb_transfer_stats {
    uint64_t h2d_bytes, h2d_calls, d2h_bytes, d2h_calls, d2d_bytes, d2d_calls;
} axb_transfer_stats;
axb_status axb_get_transfer_stats(axb_transfer_stats *out);
/* Pre-warms the caching allocator of the current device with one segment of `bytes` (split on demand), so a steady
 * state that fits it never calls cudaMalloc. */
axb_status axb_pool_reserve(size_t bytes);

/* ------------------------------------------------------------------------------------------------
 * Family 1 — fused elementwise / broadcast / where chains (one launch per fusion group).
 * Replaces: CPUBinaryOperation / CPUUnaryOperation / CPUWhereOperation
 *   src/backends/cpu/cpu_operations.cpp:28-88,280-583,824-914,2408-2589 and the fused executors
 *   execute_gpu_fused_chain (include/axiom/graph/gpu_fusion.hpp:14-17),
 *   try_tiled_fused_loop / dispatch_fused_pattern (src/graph/graph_executor.cpp:529-625).
 *
 * A program is a straight-line list of typed instructions over `n_inputs` strided/broadcast tensors.
 * Instruction i produces value v_i of dtype `dtype`; operands name an input (src >= 0) or an earlier
 * value (src = AXB_SRC_VALUE(k), k < i) or an immediate (AXB_SRC_IMM: `imm`, first converted to the
 * operand dtype the op computes in).
 * Semantics mirror the reference op by op: operands are converted to the result dtype (binary ops:
 * `astype(result)`, cpu_operations.cpp:290-312; comparisons/logical: both sides to Float32 first,
 * :115-126,286-296), the op is evaluated in that type, and every intermediate is rounded to its dtype
 * (fused_kernel.cpp chain semantics) — so a fused chain is bit-identical to running the ops one by one.
 * The value of the LAST instruction is written to `out` (contiguous or strided, same shape as the
 * broadcast of all inputs; out->dtype must equal the last instruction's dtype).
 * ---------------------------------------------------------------------------------------------- */
#define AXB_SRC_NONE (INT32_MIN)
#define AXB_SRC_IMM (INT32_MIN + 1)
#define AXB_SRC_VALUE(k) (-1 - (int32_t)(k))
#define AXB_MAX_INSTRS 32
#define AXB_MAX_INPUTS 8

typedef struct axb_instr {
    int32_t op;     /* axb_op: unary, binary, AXB_OP_WHERE (src = cond,a,b), AXB_OP_CAST */
    int32_t dtype;  /* result dtype of this instruction */
    int32_t src[3]; /* operand selectors */
    double imm;     /* immediate used by AXB_SRC_IMM operands; LeakyReLU slope lives here too */
} axb_instr;

axb_status axb_elementwise(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                           const axb_tensor *out, axb_stream_t stream);

/* Compile-only twin of axb_elementwise (needs no GPU): makes sure the kernel specialised for this program
 * SIGNATURE (ops, dtypes, operand access modes — never shapes or immediates) is in the in-tree cubin cache
 * axiom_b200/kernel_cache/. Tensor descriptors only need plausible alignment (base may be any non-null
 * address). build() calls it for the BASELINE configs; unseen signatures are specialised by NVRTC on first
 * use (the analogue of the reference's per-signature MPSGraph cache, mpsgraph_fused.mm:373-377). */
axb_status axb_elementwise_precompile(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                      const axb_tensor *out);
/* softmax (is_log = 0) / log_softmax over the LAST axis of the chain's value, optionally followed by a row reduction of the
 * result over that axis (epilogue_op 0 = none: `out` holds the rows; AXB_OP_SUM / MEAN / MAX: one value per row) — SURVEY §7
 * step 6 "softmax with elementwise prologue / epilogue": `(x * s + m).softmax(-1)` and the README chain's
 * `.softmax(-1).sum({1})` are ONE kernel for fp32 chains with rows of at most 8192 elements (values stay in registers between
 * max, sum and normalisation); otherwise chain, axb_softmax and axb_reduce run back to back. `out` contiguous. */
axb_status axb_elementwise_softmax(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int is_log,
                                   int32_t epilogue_op, const axb_tensor *out, axb_stream_t stream);
/* layer_norm (kind AXB_OP_LAYER_NORM) / rms_norm (AXB_OP_RMS_NORM) over the LAST axis of the chain's value with optional fp32
 * weight / bias rows — `layer_norm(x + sublayer(x))` of a post-LN transformer in ONE kernel (fp32 chains, rows of at most
 * 8192 elements; the reference's op sequence, operations.cpp:1803-1817, :1841-1857); otherwise the chain and the stand-alone
 * kernel run back to back. `out` contiguous with the chain's shape. */
axb_status axb_elementwise_norm(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t kind,
                                const axb_tensor *weight /* may be NULL */, const axb_tensor *bias /* may be NULL */, float eps,
                                const axb_tensor *out, axb_stream_t stream);
axb_status axb_elementwise_reduce_precompile(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                             const axb_tensor *shape_desc, int32_t reduce_op, int reduce_kind);
/* Launch-plan cache of the fused-chain family: program + operand LAYOUTS (dtype, shape, byte strides, alignment class —
 * never addresses) -> specialised kernel + launch parameters. A hit skips validation, iteration-space planning, statement
 * generation and the kernel lookup; only the operand pointers are patched. (Graph-level counterpart in the C++ glue:
 * signature -> lowering plan, reference src/graph/graph_cache.cpp:12-52.) */
typedef struct axb_plan_cache_stats {
    uint64_t entries, hits, misses, evictions;
} axb_plan_cache_stats;
axb_status axb_get_plan_cache_stats(axb_plan_cache_stats *out);
axb_status axb_plan_cache_clear(void);
/* Table-lookup path of the family (csrc/half_lut.cu): a chain over ONE contiguous bf16 / fp16 tensor (>= 2 Mi elements, a
 * transcendental in it, 2-byte result) is evaluated once on all 65 536 input patterns by the chain kernel itself and then
 * streamed through a shared-memory lookup — bit-identical by construction, bandwidth-bound instead of instruction-bound. */
typedef struct axb_half_lut_stats {
    uint64_t tables, builds, hits, launches;
} axb_half_lut_stats;
axb_status axb_get_half_lut_stats(axb_half_lut_stats *out);
typedef struct axb_jit_stats {
    uint64_t mem_hits, disk_hits, compiles;
} axb_jit_stats;
axb_status axb_get_jit_stats(axb_jit_stats *out);

/* Type rules as pure host functions, so the glue infers shapes/dtypes exactly like the reference:
 * promote_types (src/tensor/operations.cpp:346-397); binary result dtype incl. comparison->Bool
 * (cpu_operations.cpp:36-49); unary result dtype (cpu_operations.cpp:836-853). */
int32_t axb_promote_types(int32_t a, int32_t b);
int32_t axb_binary_result_dtype(int32_t op, int32_t a, int32_t b);
int32_t axb_unary_result_dtype(int32_t op, int32_t a);
size_t axb_dtype_size(int32_t dtype);
/* NumPy broadcast of two shapes (ops::compute_broadcast_info, operations.cpp:270-328). */
axb_status axb_broadcast_shape(const int64_t *a, int na, const int64_t *b, int nb, int64_t *out, int *nout);

/* Single-op conveniences == Operation::execute_unary/binary/where/cast/masked_fill
 * (include/axiom/operations.hpp:223-276). `out` is a preallocated contiguous tensor of the result
 * dtype/shape (use the helpers above to size it). */
axb_status axb_unary(int32_t op, const axb_tensor *in, const axb_tensor *out, double param, axb_stream_t stream);
axb_status axb_binary(int32_t op, const axb_tensor *a, const axb_tensor *b, const axb_tensor *out,
                      axb_stream_t stream);
axb_status axb_where(const axb_tensor *cond, const axb_tensor *a, const axb_tensor *b, const axb_tensor *out,
                     axb_stream_t stream);
axb_status axb_masked_fill(const axb_tensor *in, const axb_tensor *mask, double value, const axb_tensor *out,
                           axb_stream_t stream);
axb_status axb_cast(const axb_tensor *in, const axb_tensor *out, axb_stream_t stream);
/* strided -> contiguous gather: gpu_make_contiguous (src/backends/metal/metal_operations.hpp:60-63,
 * kernels.metal:33-148), Tensor::ascontiguousarray (tensor.cpp:503-594). in/out same dtype. */
axb_status axb_gather_strided(const axb_tensor *in, const axb_tensor *out, axb_stream_t stream);
/* device fill / counter-based RNG (gpu_fill_constant, metal_operations.hpp:57-58; Tensor::randn is a
 * serial host fill + H2D in the reference, tensor.cpp:1884-1904) */
axb_status axb_fill(const axb_tensor *out, double value, axb_stream_t stream);
axb_status axb_randn(const axb_tensor *out, uint64_t seed, float mean, float std, axb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Family 2 — axis reductions, argmax/argmin, softmax, layer/RMS norm.
 * Replaces: CPUReductionOperation (cpu_operations.cpp:1307-1712), CPUArgMax/ArgMin (:2157-2350),
 * CPUSoftmaxOperation (:3067-3241), ops::layer_norm / rms_norm (src/tensor/operations.cpp:1775-1858,
 * GPU formula mpsgraph_graph_compiler.mm:303-330), execute_gpu_fused_reduction (gpu_fusion.hpp:19-22).
 * ---------------------------------------------------------------------------------------------- */
/* op in {SUM, MEAN, MAX, MIN, PROD, ANY, ALL}. axes: distinct, normalised (0 <= a < ndim); n_axes==0
 * means all axes (full reduction -> shape {1}, or all-ones with keep_dims). Half inputs accumulate in
 * fp32 and are rounded once at the end (cpu_operations.cpp:1422-1424). out dtype: see
 * axb_reduce_result_dtype. `in` may be strided/broadcast; `out` contiguous. */
int32_t axb_reduce_result_dtype(int32_t op, int32_t in_dtype);
axb_status axb_reduce(int32_t op, const axb_tensor *in, const int32_t *axes, int n_axes, int keep_dims,
                      const axb_tensor *out, axb_stream_t stream);
/* argmax/argmin over ONE axis (or axis<0: flattened), Int64 out; first index wins ties, NaN never wins,
 * all-NaN -> 0 (cpu_operations.cpp:2211-2228). */
axb_status axb_argreduce(int32_t op, const axb_tensor *in, int axis, int keep_dims, const axb_tensor *out,
                         axb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sharded FULL reductions (the one exchange step of the path). Every rank (one process per GPU) owns a
 * shard of the tensor; axb_reduce_sharded reduces the shard and combines the per-rank partials INSIDE the reduction's
 * finish kernel through mailboxes in peer memory (NVLink / NVSwitch stores + polled loads, rank-ordered combine, so
 * every rank gets the identical value) — no collective-library call and no host round trip. sum / mean / max / min of
 * f32 / bf16 / f16 shards; `global_count` = element count over all shards (the divisor of Mean). All ranks must issue
 * the same sequence of sharded reductions. A mailbox is 256 bytes of device memory; its IPC handle is exchanged by the
 * host plumbing (axiom_b200/shard.py over torch.distributed). world == 1 degenerates to axb_reduce.
 * ---------------------------------------------------------------------------------------------- */
typedef struct axb_peer_group axb_peer_group;
axb_status axb_peer_mailbox_create(void **mailbox, unsigned char ipc_handle[64]);
axb_status axb_peer_mailbox_open(const unsigned char ipc_handle[64], void **peer_mailbox);
axb_status axb_peer_mailbox_close(void *peer_mailbox);
axb_status axb_peer_mailbox_destroy(void *mailbox);
axb_status axb_peer_group_create(int rank, int world, void *const *mailboxes /* world pointers, [rank] = own */,
                                 axb_peer_group **group);
axb_status axb_peer_group_destroy(axb_peer_group *group);
/* AXB_ERR_CUDA once any exchange of this group has timed out (a peer never delivered its partial within ~20 s; that call's
 * result is NaN on this rank whatever the op). Also checked at the start of every axb_reduce_sharded. Never blocks. */
axb_status axb_peer_group_status(const axb_peer_group *group);
axb_status axb_reduce_sharded(int32_t op, const axb_tensor *in_shard, int64_t global_count, axb_peer_group *group,
                              const axb_tensor *out, axb_stream_t stream);
/* softmax / log_softmax along `axis` (any axis; fp32/bf16/fp16/f64; half via fp32 round trip). */
axb_status axb_softmax(const axb_tensor *in, int axis, int is_log, const axb_tensor *out, axb_stream_t stream);
/* y = (x-mean)/sqrt(var+eps)*w + b over `axis` (biased variance). weight/bias may be NULL. */
axb_status axb_layer_norm(const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias, int axis,
                          float eps, const axb_tensor *out, axb_stream_t stream);
axb_status axb_rms_norm(const axb_tensor *in, const axb_tensor *weight, int axis, float eps,
                        const axb_tensor *out, axb_stream_t stream);
/* Which kernel the LAST softmax / layer_norm / rms_norm of the calling thread ran on: 0 = generic (register-resident or
 * streaming rows), 1 = wide TMA-staged persistent kernel (rows of 32..128 KiB), 2 = SM-pair cluster kernel (wide rows
 * with affine parameters resident in shared memory). -1 before the first call. */
int axb_rownorm_last_path(void);
/* elementwise prologue fused into a FULL reduction (FusedReductionStep, compiled_graph.hpp:61-66):
 * the program's last value is reduced with `op` instead of being stored. */
axb_status axb_elementwise_reduce(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                  int32_t reduce_op, const axb_tensor *out, axb_stream_t stream);
/* The chain's value under a reduction over ONE axis (the reference fuses elementwise producers into any reduction,
 * src/graph/graph_compiler.cpp:376-478): fp32 chains reduced over their LAST axis by sum / mean / max / min into a
 * contiguous fp32 `out` run as ONE kernel (a CTA or a warp per output row folds the chain's values as it computes them);
 * every other combination runs the chain and axb_reduce as two launches. `out` has the shape axb_reduce would produce. */
axb_status axb_elementwise_reduce_axis(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs,
                                       int32_t reduce_op, int axis, int keep_dims, const axb_tensor *out, axb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Family 3 — matmul / batched matmul (tcgen05 + TMEM + TMA for bf16/fp16; fp32 via 3xTF32 or FFMA).
 * Replaces: CPUMatMulOperation::execute_matmul (cpu_operations.cpp:2089-2150), matmul_2d (:1773-1883),
 * BlasBackend::sgemm (src/backends/cpu/blas/blas_backend.hpp:38); MatMulActivationStep
 * (compiled_graph.hpp:54-58). NumPy batch broadcasting; transposes are folded into strides, never
 * materialised. out dtype == promote(a,b).
 * ---------------------------------------------------------------------------------------------- */
#define AXB_MM_CONST_A 1 /* operand A is a constant (weights): derived operand forms may be cached */
#define AXB_MM_CONST_B 2 /* operand B is a constant */
typedef struct axb_matmul_epilogue {
    const axb_tensor *bias; /* NULL, or broadcastable to [.., M, N] (Linear: {N}) — nn/linear.cpp:22-25 */
    int32_t activation;     /* 0 = none, else a unary axb_op (RELU, GELU, SILU, SIGMOID, TANH). Same function as the
                             * elementwise op (bit-identical for half-precision outputs); GELU of a float32 tensor-core
                             * result evaluates the reference's Abramowitz-Stegun erf with MUFU reciprocal / exponential:
                             * within 3e-7 * max(1, |x|) of the elementwise op, below the matmul's own rounding. */
    /* AXB_MM_CONST_*: the caller promises that the flagged operand's bytes only change through this library (axb_free of
     * its block, axb_memcpy_* / axb_memset into it) or are announced with axb_invalidate_cached. The fp32 tensor-core path
     * then keeps the operand's three-piece bf16 split (nn::Linear / conv weights) instead of re-splitting it per call. */
    int32_t flags;
} axb_matmul_epilogue;
/* Drops every cached operand form derived from device memory overlapping [ptr, ptr + bytes) (in-place writes by kernels). */
axb_status axb_invalidate_cached(const void *ptr, size_t bytes);
typedef struct axb_const_cache_stats {
    uint64_t entries, bytes, hits, misses, invalidations;
} axb_const_cache_stats;
axb_status axb_get_const_cache_stats(axb_const_cache_stats *out);
axb_status axb_matmul(const axb_tensor *a, const axb_tensor *b, int transpose_a, int transpose_b,
                      const axb_matmul_epilogue *epilogue /* may be NULL */, const axb_tensor *out,
                      axb_stream_t stream);
/* Which kernel the LAST axb_matmul of the calling thread ran on (tests assert the path instead of trusting a heuristic):
 * 0 = CUDA-core (SIMT) kernel, 1 = tcgen05 one-CTA tiles, 2 = tcgen05 CTA-pair tiles (cta_group::2, 256x256),
 * 3 = fp32 on the tensor cores (three-way bf16 split). */
int axb_matmul_last_path(void);
/* result shape per matmul rules (graph_registry.cpp:147-220) */
axb_status axb_matmul_shape(const axb_tensor *a, const axb_tensor *b, int transpose_a, int transpose_b,
                            int64_t *out_shape, int *out_ndim);

/* ------------------------------------------------------------------------------------------------
 * Indexing needed by the parity suite / whisper embedding (cpu_operations.cpp gather/index_select).
 * ---------------------------------------------------------------------------------------------- */
axb_status axb_gather(const axb_tensor *in, int dim, const axb_tensor *indices, const axb_tensor *out,
                      axb_stream_t stream);
axb_status axb_index_select(const axb_tensor *in, int dim, const axb_tensor *indices, const axb_tensor *out,
                            axb_stream_t stream);
/* CPUScatterOperation (cpu_operations.cpp:2886-2960): out = copy(in); out[.., indices[c], ..] = src[c] along dim */
axb_status axb_scatter(const axb_tensor *in, int dim, const axb_tensor *indices, const axb_tensor *src,
                       const axb_tensor *out, axb_stream_t stream);

/* ---- scaled dot-product attention ("next" row 1) ---------------------------------------------------------------
 * ops::scaled_dot_product_attention (src/tensor/operations.cpp:3979-4044) / cpu_flash_attention
 * (src/backends/cpu/flash_attention_cpu.cpp:18-131): q {B,H,Lq,D}, k {B,H,Lk,D}, v {B,H,Lk,Dv}, out {B,H,Lq,Dv}, Float32.
 * mask (may be NULL): Bool / 8-bit, broadcastable to {B,H,Lq,Lk}, NON-ZERO = position masked out; is_causal masks kv > q;
 * scale <= 0 means 1/sqrt(D). head_dim <= 64: ONE tcgen05 kernel (flash_attention.cu: Q K^T and P V on the tensor cores, online
 * softmax in registers, P handed to the second MMA through tensor memory, no score matrix in HBM) after a one-launch operand
 * prologue; Float32 (exact three-piece bf16 split), BFloat16 and Float16 (native). Wider heads, Float32 only: three launches
 * (Q K^T, fused scale+mask+softmax writing pre-split probabilities, P V). */
axb_status axb_sdpa(const axb_tensor *q, const axb_tensor *k, const axb_tensor *v, const axb_tensor *mask /* may be NULL */,
                    float scale, int is_causal, const axb_tensor *out, axb_stream_t stream);
/* Which path the LAST axb_sdpa of the calling thread took: 1 = single flash kernel, 0 = three-kernel path, -1 none yet. */
int axb_sdpa_last_path(void);

/* ---- convolution / padding data movement ("next" row: conv1d as GEMM) ------------------------------------------
 * Replaces the reference's naive conv1d loops (src/tensor/operations.cpp:3046-3135) together with axb_matmul:
 * col = im2col(x) with implicit zero padding, then W2d{C_out, C_in*K} x col{N, C_in*K, L_out}.
 * axb_pad_constant is OpType::Pad in constant mode (src/tensor/operations.cpp:2150-2260, lazy node
 * src/graph/graph_registry.cpp:707-722). in may be strided; out must be contiguous. */
axb_status axb_im2col_1d(const axb_tensor *in /* (N, C, L) */, int kernel_size, int stride, int padding, int dilation,
                         const axb_tensor *out /* (N, C*K, L_out) */, axb_stream_t stream);
axb_status axb_pad_constant(const axb_tensor *in, const int64_t *pad_lo, const int64_t *pad_hi, double value,
                            const axb_tensor *out, axb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Spatial operators next to the path. Replace the reference's scalar CPU loops / MPSGraph calls:
 *   axb_conv2d          ops::conv2d (src/tensor/operations.cpp:3141-3330; lazy OpType::Conv2D on the GPU, :3211-3220):
 *                       in (N, C_in, H, W), weight (C_out, C_in/groups, KH, KW), bias {C_out} or NULL, out (N, C_out, OH, OW)
 *   axb_conv_transpose  ops::conv_transpose1d / 2d (:3340-3690, hook backends::metal::gpu_conv_transpose1d/2d :3395, :3563):
 *                       spatial_dims 1: in (N, C_in, L), weight (C_in, C_out/groups, K); 2: the 4-d forms. The output
 *                       shape (which carries output_padding) is the caller's, as in the reference.
 *   axb_interpolate     ops::interpolate (:3694-3975, hook gpu_interpolate :3742): in (N, C, L) or (N, C, H, W);
 *                       mode 0 nearest, 1 linear / bilinear, 2 bicubic (1-d inputs: every non-nearest mode is linear)
 * Inputs and weights may be strided views; outputs are contiguous; floating dtypes; fp32 results of the gather kernels are
 * bit-identical to the CPU loops (same accumulation type and term order). Large axb_conv_transpose problems (groups == 1,
 * >= 2e8 flop, C_out >= 32, C_in >= 8, every output phase reached by at least one filter tap; float32 / bf16 / fp16) run on the
 * tensor cores as one implicit-GEMM convolution per output phase (axb_matmul_last_path() == 4): float32 through the exact
 * three-piece bf16 split — fp32-FMA accuracy, tolerance-based parity like axb_conv.
 * ---------------------------------------------------------------------------------------------- */
axb_status axb_conv2d(const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias /* may be NULL */,
                      const int32_t stride[2], const int32_t padding[2], const int32_t dilation[2], int groups,
                      const axb_tensor *out, axb_stream_t stream);
/* Forward convolution, 1-D ((N, C, L) / (C_out, C_in/groups, K)) or 2-D, optional fused activation (0 or a unary axb_op).
 * groups == 1 and >= 1e8 flop, float32 / bf16 / fp16: IMPLICIT GEMM on the tensor cores — no im2col matrix exists; the K loop
 * of the tcgen05 GEMM walks (filter tap, 64-channel block) steps and the TMA load of each input tile starts at the tap's
 * shifted coordinate, out-of-image rows / columns zero-filled by TMA (= the padding); strides read a phase-de-interleaved
 * copy; bias and activation run in the GEMM epilogue. float32 uses the exact three-piece bf16 split (fp32-FMA accuracy,
 * tolerance-based parity); axb_matmul_last_path() reports 4. Everything else runs the gather kernel of axb_conv2d (fp64
 * accumulation, bit-identical to the CPU loops; an activation runs as a second, in-place launch). Replaces ops::conv1d / conv2d
 * (src/tensor/operations.cpp:2968-3330) and the MPSGraph convolution lowering (mpsgraph_graph_compiler.mm:365-405). */
axb_status axb_conv(const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias /* may be NULL */, int spatial_dims,
                    const int32_t *stride, const int32_t *padding, const int32_t *dilation, int groups, int activation,
                    const axb_tensor *out, axb_stream_t stream);
axb_status axb_conv_transpose(const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias /* may be NULL */,
                              int spatial_dims, const int32_t *stride, const int32_t *padding, const int32_t *dilation,
                              int groups, const axb_tensor *out, axb_stream_t stream);
axb_status axb_interpolate(const axb_tensor *in, int mode, int align_corners, const axb_tensor *out, axb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pooling, sorting, masked_select. The reference has no device branch for these: a GPU tensor is
 * copied to the host, processed by scalar loops and (pooling) uploaded again.
 *   axb_pool           ops::max_pool1d/2d/3d, avg_pool1d/2d/3d (src/tensor/operations.cpp:2367-2880; the adaptive forms
 *                      :2883-2962 call them): in (N, C, spatial...), out contiguous (N, C, out...) with
 *                      out = (in + 2*padding - kernel) / stride + 1 (:2360-2363). fp64 accumulation in the CPU loops' order:
 *                      bit-identical for every float dtype. Input may be a strided view.
 *   axb_sort           ops::sort / argsort / topk (:4053-4257): contiguous keys of float32 / float64 / int32 / int64 (the
 *                      reference converts every other dtype to float32 first), sorted along `axis`. `values` / `indices`
 *                      (either may be NULL) are contiguous with extent k <= n along `axis`: k == n is sort / argsort, k < n
 *                      the first k of the order = topk (largest: descending = 1). Equal keys keep ascending positions (the
 *                      reference's std::sort leaves their order unspecified).
 *   axb_masked_select  ops::masked_select (src/backends/cpu/cpu_operations.cpp:2719-2789): broadcasts input and mask, selects
 *                      where the FIRST BYTE of the mask element is non-zero, row-major order. The result block comes from
 *                      axb_alloc (release it with axb_free); *out_count elements of the input's dtype (NULL when 0). The
 *                      count is read back inside the call (8 bytes device->host, one stream synchronisation).
 * ---------------------------------------------------------------------------------------------- */
axb_status axb_pool(const axb_tensor *in, int is_avg, int spatial_dims, const int32_t *kernel, const int32_t *stride,
                    const int32_t *padding, int count_include_pad, const axb_tensor *out, axb_stream_t stream);
axb_status axb_sort(const axb_tensor *in, int axis, int descending, const axb_tensor *values /* may be NULL */,
                    const axb_tensor *indices /* int64, may be NULL */, axb_stream_t stream);
axb_status axb_masked_select(const axb_tensor *in, const axb_tensor *mask, void **out_data, int64_t *out_count,
                             axb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer entry points (what `Tensor::gpu()` ... `.cpu()` does end to end: H2D, launch, D2H, sync).
 * Inputs/outputs are HOST tensors (base = host pointer, contiguous or strided); used by bench.py's
 * `e2e` leg and by callers that hold no device memory. Staging goes through the caching allocator and
 * the pinned pool.
 * ---------------------------------------------------------------------------------------------- */
axb_status axb_elementwise_host(const axb_instr *prog, int n_instr, const axb_tensor *host_inputs,
                                int n_inputs, const axb_tensor *host_out);
axb_status axb_reduce_host(int32_t op, const axb_tensor *host_in, const int32_t *axes, int n_axes,
                           int keep_dims, const axb_tensor *host_out);
axb_status axb_matmul_host(const axb_tensor *host_a, const axb_tensor *host_b, int transpose_a,
                           int transpose_b, const axb_tensor *host_out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* AXB_H */
