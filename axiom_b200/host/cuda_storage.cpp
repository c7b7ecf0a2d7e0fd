// cuda_storage.cpp — seam 1: Storage on device memory, the make_storage factory, and the stream / availability
// symbols the shared code calls (see cuda_backend.hpp for the map of reference call sites).
#include "cuda_backend.hpp"
#include "cuda_plan.hpp"
#include "axiom/io/io.hpp"

#include <algorithm>
#include <cstring>
#include <mutex>

#include "backends/cpu/cpu_storage.hpp"
#include "backends/metal/metal_common.hpp"
#include "backends/metal/metal_operations.hpp"
#include "backends/metal/unified_storage.hpp"

namespace axiom {
namespace backends {
namespace cuda {

void check(axb_status st, const char *what) {
    if (st == AXB_OK)
        return;
    const std::string msg = std::string(what) + ": " + axb_last_error();
    switch (st) {
    case AXB_ERR_INVALID: throw ValueError(msg);
    case AXB_ERR_UNSUPPORTED: throw TypeError(msg);
    case AXB_ERR_OOM: throw MemoryError(msg);
    case AXB_ERR_SHAPE: throw ShapeError(msg);
    default: throw DeviceError(msg);
    }
}

bool is_cuda_available() {
    static const bool ok = [] {
        int n = 0;
        return axb_device_count(&n) == AXB_OK && n > 0;
    }();
    return ok;
}

void ensure_initialized() {
    static thread_local bool bound = false;
    if (!bound) {
        if (!is_cuda_available())
            throw DeviceError::not_available("CUDA GPU (no sm_100 device visible to libaxb)");
        const char *dev = std::getenv("AXIOM_CUDA_DEVICE"); // one process per GPU: LOCAL_RANK picks the device
        if (!dev)
            dev = std::getenv("LOCAL_RANK");
        check(axb_set_device(dev ? std::atoi(dev) : 0), "axb_set_device");
        bound = true;
    }
    static std::once_flag once;
    std::call_once(once, [] { register_cuda_operations(); });
}

CudaStorage::CudaStorage(size_t size_bytes) : size_bytes_(size_bytes) {
    ensure_initialized();
    check(axb_get_device(&device_index_), "axb_get_device");
    if (size_bytes_ > 0) { // size 0 is legal (tests/test_tensor_basic.cpp: empty tensors)
        check(axb_alloc(size_bytes_, nullptr, &ptr_), "axb_alloc");
        plan::note_alloc(ptr_, size_bytes_); // a temporary of the lowering being recorded (no-op otherwise)
    }
}

CudaStorage::CudaStorage(void *adopted, size_t size_bytes, Adopt) : ptr_(adopted), size_bytes_(size_bytes) {
    check(axb_get_device(&device_index_), "axb_get_device");
}

CudaStorage::~CudaStorage() {
    if (ptr_)
        axb_free(ptr_, nullptr); // stream-ordered: safe against kernels still queued on the backend stream
}

void *CudaStorage::data() { throw DeviceError::cpu_only("direct data access to CUDA storage"); }
const void *CudaStorage::data() const { throw DeviceError::cpu_only("direct data access to CUDA storage"); }

void CudaStorage::copy_to(Storage &other) const {
    const size_t n = std::min(size_bytes_, other.size_bytes());
    if (n == 0)
        return;
    if (other.device() == Device::CPU) {
        // synchronous: the caller reads the host bytes as soon as this returns (metal_storage.mm:69-70)
        plan::note_foreign("device->host copy inside a lowering");
        check(axb_memcpy_d2h(other.data(), ptr_, n, nullptr, /*sync=*/1), "axb_memcpy_d2h");
    } else {
        auto *dst = dynamic_cast<CudaStorage *>(&other);
        if (!dst)
            throw DeviceError("copy_to: destination GPU storage is not CUDA storage");
        check(axb_memcpy_d2d(dst->ptr_, ptr_, n, nullptr), "axb_memcpy_d2d");
    }
}

void CudaStorage::copy_from(const Storage &other) {
    const size_t n = std::min(size_bytes_, other.size_bytes());
    if (n == 0)
        return;
    if (other.device() == Device::CPU) {
        // The reference's host tensors are pageable memory: cudaMemcpyAsync stages such a source through the driver's
        // pinned buffer and returns once the HOST buffer has been consumed, so the caller may free or overwrite it
        // immediately; the device side stays ordered on the backend stream. No stream sync here — scalar operands
        // (`x + 1.0f` uploads a one-element tensor, tensor_operators.hpp) would otherwise serialise host and GPU.
        check(axb_memcpy_h2d(ptr_, other.data(), n, nullptr), "axb_memcpy_h2d");
        plan::note_upload(ptr_, other.data(), n);
    } else {
        auto *src = dynamic_cast<const CudaStorage *>(&other);
        if (!src)
            throw DeviceError("copy_from: source GPU storage is not CUDA storage");
        check(axb_memcpy_d2d(ptr_, src->ptr_, n, nullptr), "axb_memcpy_d2d");
    }
}

std::unique_ptr<Storage> CudaStorage::clone() const {
    auto s = std::make_unique<CudaStorage>(size_bytes_);
    s->copy_from(*this);
    return s;
}

std::map<std::string, Tensor> load_file_tensors(const std::string &filename, const std::vector<FileTensor> &items) {
    ensure_initialized();
    std::vector<const FileTensor *> order;
    for (const auto &it : items)
        order.push_back(&it);
    std::sort(order.begin(), order.end(), [](const FileTensor *a, const FileTensor *b) { return a->file_offset < b->file_offset; });
    std::map<std::string, Tensor> result;
    for (const FileTensor *it : order) {
        Tensor t = empty_gpu(it->shape, it->dtype);
        if (it->bytes != t.nbytes()) // same check and message as io_safetensors.cpp:137-144
            throw io::FileFormatError("SafeTensors: tensor '" + it->name + "' size mismatch: file=" + std::to_string(it->bytes) +
                                  " expected=" + std::to_string(t.nbytes()));
        if (it->bytes != 0) {
            auto *cs = dynamic_cast<CudaStorage *>(t.storage().get());
            const axb_status st = axb_upload_file(filename.c_str(), it->file_offset, it->bytes, cs->device_ptr(), nullptr);
            if (st != AXB_OK) // unreadable / truncated file: the reference's error type (io_safetensors.cpp:148-151)
                throw io::SerializationError("SafeTensors: failed to read tensor '" + it->name + "' data: " + axb_last_error());
        }
        result[it->name] = t;
    }
    return result;
}

axb_tensor describe(const Tensor &t) {
    if (t.device() != Device::GPU)
        throw DeviceError::mismatch("GPU", "CPU");
    auto st = t.storage(); // materialises lazy tensors
    auto *cs = dynamic_cast<CudaStorage *>(st.get());
    if (!cs)
        throw DeviceError("GPU tensor is not backed by CUDA storage");
    const Shape &sh = t.shape();
    const Strides &sd = t.strides();
    if (sh.size() > AXB_MAX_DIMS)
        throw ShapeError("CUDA backend supports at most 8 dimensions");
    axb_tensor d{};
    d.base = cs->device_ptr();
    d.byte_offset = static_cast<int64_t>(t.offset());
    d.ndim = static_cast<int32_t>(sh.size());
    d.dtype = to_axb_dtype(t.dtype());
    for (size_t i = 0; i < sh.size(); ++i) {
        d.shape[i] = static_cast<int64_t>(sh[i]);
        d.strides[i] = static_cast<int64_t>(sd[i]);
    }
    return d;
}

Tensor empty_gpu(const Shape &shape, DType dtype) { return Tensor(shape, dtype, Device::GPU); }

Tensor on_gpu(const Tensor &t) { return t.device() == Device::GPU ? t : t.to(Device::GPU); }

} // namespace cuda

// ---- the names the shared code calls on its GPU paths -------------------------------------------------------------
namespace metal {

bool is_metal_available() { return cuda::is_cuda_available(); }
void register_metal_operations() { cuda::register_cuda_operations(); }
bool is_unified_memory_available() { return false; }

// Tensor::to / Tensor::cpu synchronise "the GPU stream" before host reads (tensor.cpp:1608,1648)
MetalExecutionStream::MetalExecutionStream()
    : command_queue_(nullptr), current_buffer_(nullptr), mps_buffer_(nullptr), execution_descriptor_(nullptr),
      compilation_descriptor_(nullptr), compute_encoder_(nullptr), batch_count_(0) {}
MetalExecutionStream::~MetalExecutionStream() = default;
MetalExecutionStream &MetalExecutionStream::instance() {
    static MetalExecutionStream s;
    return s;
}
void MetalExecutionStream::synchronize() {
    if (cuda::is_cuda_available()) {
        cuda::ensure_initialized();
        cuda::check(axb_stream_sync(nullptr), "axb_stream_sync");
    }
}
void MetalExecutionStream::commit() {}
bool MetalExecutionStream::has_pending_work() const { return false; }

// No unified memory on a discrete GPU: the class only has to exist so that the reference's
// dynamic_cast<UnifiedStorage*> (tensor.cpp:1604,1646) links and yields nullptr. It is never instantiated.
UnifiedStorage::UnifiedStorage(void *, size_t size_bytes, Device tag)
    : device_(nullptr), buffer_(nullptr), cached_contents_(nullptr), size_bytes_(size_bytes), offset_(0), device_tag_(tag) {
    throw DeviceError::not_available("unified memory");
}
UnifiedStorage::UnifiedStorage(void *, void *, void *, size_t size_bytes, size_t offset, Device tag)
    : device_(nullptr), buffer_(nullptr), cached_contents_(nullptr), size_bytes_(size_bytes), offset_(offset), device_tag_(tag) {}
UnifiedStorage::~UnifiedStorage() = default;
void *UnifiedStorage::data() { return cached_contents_; }
const void *UnifiedStorage::data() const { return cached_contents_; }
size_t UnifiedStorage::size_bytes() const { return size_bytes_; }
void UnifiedStorage::copy_to(Storage &) const { throw DeviceError::not_available("unified memory"); }
void UnifiedStorage::copy_from(const Storage &) { throw DeviceError::not_available("unified memory"); }
std::unique_ptr<Storage> UnifiedStorage::clone() const { throw DeviceError::not_available("unified memory"); }
std::unique_ptr<UnifiedStorage> UnifiedStorage::with_device_tag(Device) const {
    throw DeviceError::not_available("unified memory");
}

} // namespace metal
} // namespace backends

// ---- seam 1 factory: replaces src/tensor/storage.cpp in this build (that TU only knows Metal behind __APPLE__) ----
std::unique_ptr<Storage> make_storage(size_t size_bytes, Device device) {
    switch (device) {
    case Device::CPU: return backends::cpu::make_cpu_storage(size_bytes);
    case Device::GPU: return std::make_unique<backends::cuda::CudaStorage>(size_bytes);
    }
    throw DeviceError("Unknown device type");
}

// GPU availability probe. src/tensor/system.cpp answers "false" off Apple platforms at compile time (system.cpp:10-17), so
// that TU is left out of this build too; the reference's own GPU tests gate on should_run_gpu_tests().
namespace system {
bool is_metal_available() { return backends::cuda::is_cuda_available(); }
bool should_run_gpu_tests() {
    const char *off = std::getenv("AXIOM_SKIP_GPU_TESTS");
    return !(off && std::strcmp(off, "1") == 0) && backends::cuda::is_cuda_available();
}
std::string device_to_string(Device device) { return device == Device::CPU ? "CPU" : device == Device::GPU ? "GPU" : "Unknown"; }
} // namespace system

} // namespace axiom
