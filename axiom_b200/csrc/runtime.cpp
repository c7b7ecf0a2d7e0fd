// runtime.cpp — device context, backend stream, error slot, transfers.
//
// B200 counterpart of the reference's MetalExecutionStream singleton (one async stream per device,
// host synchronises only when it reads: src/backends/metal/metal_common.hpp:53-123,
// metal_storage.mm:66-93) and of system::is_metal_available (src/tensor/system.cpp:12-29).
#include <cstring>
#include <sys/syscall.h>
#include <unistd.h>
#include <mutex>
#include <thread>
#include <atomic>
#include <vector>

#include <fcntl.h>
#include <unistd.h>
#include <cerrno>
#include <mutex>

#include "axb_internal.h"
#include "gemm.h"

namespace axb {

std::atomic<uint64_t> g_launches{0};
TransferCounters g_xfer;

std::string &last_error_ref() {
    thread_local std::string e;
    return e;
}

axb_status fail(axb_status st, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return st;
}

namespace {
std::mutex g_ctx_mu;
std::vector<DeviceCtx *> g_ctx; // indexed by device ordinal, created on first use, never destroyed
} // namespace

DeviceCtx *current_ctx() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        fail(AXB_ERR_CUDA, "cudaGetDevice failed: %s (is a CUDA device present?)", cudaGetErrorString(e));
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if ((int)g_ctx.size() <= dev)
        g_ctx.resize(dev + 1, nullptr);
    if (!g_ctx[dev]) {
        auto *c = new DeviceCtx();
        c->device = dev;
        cudaDeviceProp p{};
        e = cudaGetDeviceProperties(&p, dev);
        if (e != cudaSuccess) {
            fail(AXB_ERR_CUDA, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
            delete c;
            return nullptr;
        }
        if (p.major != 10) {
            // sm_100a-only binary: refuse to pretend (no fallback path exists).
            fail(AXB_ERR_UNSUPPORTED, "device %d is sm_%d%d; this backend is built for sm_100a (B200) only", dev,
                 p.major, p.minor);
            delete c;
            return nullptr;
        }
        c->sm_count = p.multiProcessorCount;
        c->smem_optin = p.sharedMemPerBlockOptin;
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            fail(AXB_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
            delete c;
            return nullptr;
        }
        g_ctx[dev] = c;
    }
    return g_ctx[dev];
}

DeviceCtx *ctx_of_device(int dev) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    return (dev >= 0 && dev < (int)g_ctx.size()) ? g_ctx[dev] : nullptr;
}

axb_status resolve_stream(axb_stream_t s, cudaStream_t *out, DeviceCtx **ctx) {
    DeviceCtx *c = current_ctx();
    if (!c)
        return AXB_ERR_CUDA;
    if (ctx)
        *ctx = c;
    *out = s ? static_cast<cudaStream_t>(s) : c->stream;
    return AXB_OK;
}

} // namespace axb

using namespace axb;

extern "C" {

int axb_abi_version(void) { return AXB_ABI_VERSION; }
const char *axb_last_error(void) { return last_error_ref().c_str(); }
uint64_t axb_launch_count(void) { return g_launches.load(); }

axb_status axb_device_count(int *count) {
    AXB_REQUIRE(count, AXB_ERR_INVALID, "axb_device_count: null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(AXB_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *count = n;
    return AXB_OK;
}
axb_status axb_set_device(int device) {
    AXB_CUDA(cudaSetDevice(device));
    return current_ctx() ? AXB_OK : AXB_ERR_CUDA;
}
axb_status axb_get_device(int *device) {
    AXB_REQUIRE(device, AXB_ERR_INVALID, "axb_get_device: null out");
    AXB_CUDA(cudaGetDevice(device));
    return AXB_OK;
}
axb_status axb_device_name(int device, char *buf, size_t buflen) {
    AXB_REQUIRE(buf && buflen, AXB_ERR_INVALID, "axb_device_name: null buffer");
    cudaDeviceProp p{};
    AXB_CUDA(cudaGetDeviceProperties(&p, device));
    snprintf(buf, buflen, "%s", p.name);
    return AXB_OK;
}
axb_status axb_device_sm_count(int device, int *sms) {
    AXB_REQUIRE(sms, AXB_ERR_INVALID, "axb_device_sm_count: null out");
    AXB_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, device));
    return AXB_OK;
}
axb_status axb_default_stream(axb_stream_t *stream) {
    AXB_REQUIRE(stream, AXB_ERR_INVALID, "axb_default_stream: null out");
    DeviceCtx *c = current_ctx();
    if (!c)
        return AXB_ERR_CUDA;
    *stream = c->stream;
    return AXB_OK;
}
axb_status axb_stream_sync(axb_stream_t stream) {
    cudaStream_t s;
    AXB_TRY(resolve_stream(stream, &s));
    AXB_CUDA(cudaStreamSynchronize(s));
    return AXB_OK;
}

axb_status axb_event_create(axb_event_t *ev) {
    AXB_REQUIRE(ev, AXB_ERR_INVALID, "axb_event_create: null out");
    if (!current_ctx())
        return AXB_ERR_CUDA;
    cudaEvent_t e;
    AXB_CUDA(cudaEventCreate(&e));
    *ev = e;
    return AXB_OK;
}
axb_status axb_event_destroy(axb_event_t ev) {
    if (ev)
        AXB_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
    return AXB_OK;
}
axb_status axb_event_record(axb_event_t ev, axb_stream_t stream) {
    cudaStream_t s;
    AXB_TRY(resolve_stream(stream, &s));
    AXB_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev), s));
    return AXB_OK;
}
axb_status axb_event_sync(axb_event_t ev) {
    AXB_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(ev)));
    return AXB_OK;
}
axb_status axb_event_elapsed_ms(axb_event_t start, axb_event_t end, float *ms) {
    AXB_REQUIRE(ms, AXB_ERR_INVALID, "axb_event_elapsed_ms: null out");
    AXB_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(end)));
    return AXB_OK;
}

// Pinned staging memory is placed on the NUMA node the current GPU hangs off (sysfs numa_node of its PCI function):
// cudaHostAlloc faults the pages in while pinning them, so a PREFERRED memory policy around the call decides the node.
// Eight ranks that all stage through whichever node their Python process happened to start on push every transfer of
// the remote GPUs across the socket interconnect. Best effort: a container without the syscall, or a single-node host,
// changes nothing. AXB_HOST_NUMA=0 switches it off (A/B for the e2e ceiling probe).
static int gpu_numa_node(int dev) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bus; *c; ++c)
        if (*c >= 'A' && *c <= 'F')
            *c = (char)(*c - 'A' + 'a');
    char path[96];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f)
        return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1)
        node = -1;
    fclose(f);
    return node;
}
static int &last_host_numa_node() {
    static thread_local int v = -1;
    return v;
}
int axb_host_alloc_numa_node(void) { return last_host_numa_node(); }

axb_status axb_host_alloc(size_t bytes, void **ptr) {
    AXB_REQUIRE(ptr, AXB_ERR_INVALID, "axb_host_alloc: null out");
    *ptr = nullptr;
    if (bytes == 0)
        return AXB_OK;
    last_host_numa_node() = -1;
    bool policy_set = false;
    static const bool numa_on = [] {
        const char *e = getenv("AXB_HOST_NUMA");
        return !(e && e[0] == '0');
    }();
    int dev = -1;
    if (numa_on && cudaGetDevice(&dev) == cudaSuccess) {
        const int node = gpu_numa_node(dev);
        if (node >= 0 && node < 64) {
            unsigned long mask = 1ul << node;
            if (syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 65ul) == 0) {
                policy_set = true;
                last_host_numa_node() = node;
            }
        }
    }
    const cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (policy_set)
        syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
    AXB_CUDA(e);
    return AXB_OK;
}
axb_status axb_host_free(void *ptr) {
    if (ptr)
        AXB_CUDA(cudaFreeHost(ptr));
    return AXB_OK;
}
// ---- large transfers from / to PAGEABLE host memory --------------------------------------------------------------------------
// cudaMemcpyAsync on pageable memory stages through ONE driver-internal pinned buffer with a single-threaded copy (~12 GB/s
// up, ~8 GB/s down on this host against 55 GB/s per direction for pinned memory). Tensor::gpu() / .cpu() of a large host tensor
// (weights, randn() inputs: src/tensor/tensor.cpp:1593-1650, :1884-1904) therefore go through a small pool of pinned staging
// buffers filled / drained by several host threads: thread t owns two 8 MiB buffers and the chunks t, t + T, ...; it waits for
// the buffer's previous DMA, copies its chunk and enqueues the DMA on the caller's stream (H2D), or the other way round (D2H).
namespace {
constexpr size_t kStageChunk = 8u << 20;
constexpr int kStageThreads = 6;
constexpr size_t kStagedMin = 16u << 20;
struct CopyLane {
    void *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
};
struct CopyPool {
    std::mutex mu; // one staged transfer at a time per process (the lanes are shared)
    CopyLane lane[kStageThreads];
    int device = -1;
    bool ok = false;
};
CopyPool g_copy_pool;

bool pageable(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
bool staged_enabled() {
    static const bool off = getenv("AXB_NO_STAGED_COPY") != nullptr; // A/B switch
    return !off;
}
axb_status ensure_copy_pool(int device) {
    CopyPool &P = g_copy_pool;
    if (P.ok && P.device == device)
        return AXB_OK;
    for (auto &l : P.lane)
        for (int i = 0; i < 2; ++i) {
            if (l.buf[i])
                cudaFreeHost(l.buf[i]);
            if (l.ev[i])
                cudaEventDestroy(l.ev[i]);
            l.buf[i] = nullptr;
            l.ev[i] = nullptr;
        }
    P.ok = false;
    for (auto &l : P.lane)
        for (int i = 0; i < 2; ++i) {
            AXB_TRY(axb_host_alloc(kStageChunk, &l.buf[i]));
            AXB_CUDA(cudaEventCreateWithFlags(&l.ev[i], cudaEventDisableTiming));
        }
    P.device = device;
    P.ok = true;
    return AXB_OK;
}
// to_device: host `src` -> device `dst`; otherwise device `src` -> host `dst`. Returns after the host buffer has been fully
// read (H2D; the DMAs may still be in flight on `s`) or fully written (D2H).
axb_status staged_copy(void *dst, const void *src, size_t bytes, bool to_device, cudaStream_t s, int device) {
    std::lock_guard<std::mutex> lk(g_copy_pool.mu);
    AXB_TRY(ensure_copy_pool(device));
    const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
    std::atomic<int> failed{0};
    cudaEvent_t start = nullptr; // D2H: the lanes' DMAs run on side streams? no — on `s`, after everything queued before
    (void)start;
    auto worker = [&](int t) {
        if (cudaSetDevice(device) != cudaSuccess) {
            failed.store(1);
            return;
        }
        CopyLane &l = g_copy_pool.lane[t];
        int slot = 0;
        size_t pending_chunk[2] = {SIZE_MAX, SIZE_MAX};
        for (size_t c = (size_t)t; c < n_chunks && !failed.load(); c += kStageThreads) {
            const size_t off = c * kStageChunk, n = std::min(kStageChunk, bytes - off);
            if (to_device) {
                if (pending_chunk[slot] != SIZE_MAX && cudaEventSynchronize(l.ev[slot]) != cudaSuccess)
                    failed.store(1);
                memcpy(l.buf[slot], static_cast<const char *>(src) + off, n);
                if (cudaMemcpyAsync(static_cast<char *>(dst) + off, l.buf[slot], n, cudaMemcpyHostToDevice, s) != cudaSuccess ||
                    cudaEventRecord(l.ev[slot], s) != cudaSuccess)
                    failed.store(1);
                pending_chunk[slot] = c;
            } else {
                // drain the buffer's previous chunk before reusing it, then start this chunk's DMA
                if (pending_chunk[slot] != SIZE_MAX) {
                    if (cudaEventSynchronize(l.ev[slot]) != cudaSuccess)
                        failed.store(1);
                    const size_t po = pending_chunk[slot] * kStageChunk;
                    memcpy(static_cast<char *>(dst) + po, l.buf[slot], std::min(kStageChunk, bytes - po));
                }
                if (cudaMemcpyAsync(l.buf[slot], static_cast<const char *>(src) + off, n, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                    cudaEventRecord(l.ev[slot], s) != cudaSuccess)
                    failed.store(1);
                pending_chunk[slot] = c;
            }
            slot ^= 1;
        }
        for (int k = 0; k < 2; ++k) { // tail: H2D buffers may be reused by the next call only after their DMA; D2H still has to drain
            if (pending_chunk[k] == SIZE_MAX)
                continue;
            if (cudaEventSynchronize(l.ev[k]) != cudaSuccess)
                failed.store(1);
            if (!to_device) {
                const size_t po = pending_chunk[k] * kStageChunk;
                memcpy(static_cast<char *>(dst) + po, l.buf[k], std::min(kStageChunk, bytes - po));
            }
        }
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>(kStageThreads, n_chunks);
    for (int t = 1; t < nt; ++t)
        th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th)
        x.join();
    if (failed.load()) {
        cudaGetLastError();
        return fail(AXB_ERR_CUDA, "staged %s copy of %zu bytes failed", to_device ? "host->device" : "device->host", bytes);
    }
    return AXB_OK;
}
} // namespace

axb_status axb_memcpy_h2d(void *dst, const void *src, size_t bytes, axb_stream_t stream) {
    if (!bytes)
        return AXB_OK;
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    gemm::invalidate_const_cache(dst, bytes); // cached forms of overwritten constants
    if (bytes >= kStagedMin && staged_enabled() && pageable(src)) {
        AXB_TRY(staged_copy(dst, src, bytes, true, s, ctx->device));
        count_h2d(bytes);
        return AXB_OK;
    }
    AXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    count_h2d(bytes);
    return AXB_OK;
}
// File -> device without a pageable stop: pread() fills pinned staging buffers (two, reused for the life of the process)
// while the previous chunk's DMA runs. Used for weight loading (safetensors), where the reference reads every tensor into a
// pageable host tensor and then copies it to the device (src/io/io_safetensors.cpp:123-160).
namespace {
struct FileStage {
    void *buf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool busy[2] = {false, false};
    size_t chunk = 0;
    int device = -1;
};
std::mutex g_stage_mu;
FileStage g_stage;
} // namespace

axb_status axb_upload_file(const char *path, uint64_t file_offset, size_t bytes, void *dst, axb_stream_t stream) {
    AXB_REQUIRE(path, AXB_ERR_INVALID, "axb_upload_file: null path");
    if (!bytes)
        return AXB_OK;
    AXB_REQUIRE(dst, AXB_ERR_INVALID, "axb_upload_file: null destination");
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const int fd = ::open(path, O_RDONLY);
    AXB_REQUIRE(fd >= 0, AXB_ERR_INVALID, "axb_upload_file: cannot open %s", path);
    std::lock_guard<std::mutex> lk(g_stage_mu);
    constexpr size_t kChunk = 16u << 20;
    if (!g_stage.buf[0] || g_stage.device != ctx->device) {
        for (int i = 0; i < 2; ++i) {
            if (g_stage.buf[i])
                cudaFreeHost(g_stage.buf[i]);
            if (g_stage.done[i])
                cudaEventDestroy(g_stage.done[i]);
            g_stage.buf[i] = nullptr;
            g_stage.done[i] = nullptr;
            g_stage.busy[i] = false;
        }
        for (int i = 0; i < 2; ++i) {
            const axb_status st = axb_host_alloc(kChunk, &g_stage.buf[i]);
            if (st != AXB_OK) {
                ::close(fd);
                return st;
            }
            if (cudaEventCreateWithFlags(&g_stage.done[i], cudaEventDisableTiming) != cudaSuccess) {
                ::close(fd);
                return fail(AXB_ERR_CUDA, "axb_upload_file: cudaEventCreate failed");
            }
        }
        g_stage.chunk = kChunk;
        g_stage.device = ctx->device;
    }
    gemm::invalidate_const_cache(dst, bytes);
    axb_status st = AXB_OK;
    size_t done = 0;
    int slot = 0;
    while (done < bytes && st == AXB_OK) {
        const size_t n = std::min(g_stage.chunk, bytes - done);
        if (g_stage.busy[slot]) { // the DMA that last read this buffer must have finished
            if (cudaEventSynchronize(g_stage.done[slot]) != cudaSuccess)
                st = fail(AXB_ERR_CUDA, "axb_upload_file: staging event failed");
            g_stage.busy[slot] = false;
        }
        size_t got = 0;
        while (got < n && st == AXB_OK) {
            const ssize_t r = ::pread(fd, static_cast<char *>(g_stage.buf[slot]) + got, n - got, (off_t)(file_offset + done + got));
            if (r < 0 && errno == EINTR)
                continue;
            if (r <= 0)
                st = fail(AXB_ERR_INVALID, "axb_upload_file: short read from %s at byte %llu", path, (unsigned long long)(file_offset + done + got));
            else
                got += (size_t)r;
        }
        if (st != AXB_OK)
            break;
        if (cudaMemcpyAsync(static_cast<char *>(dst) + done, g_stage.buf[slot], n, cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaEventRecord(g_stage.done[slot], s) != cudaSuccess) {
            st = fail(AXB_ERR_CUDA, "axb_upload_file: cudaMemcpyAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        g_stage.busy[slot] = true;
        count_h2d(n);
        done += n;
        slot ^= 1;
    }
    ::close(fd);
    return st;
}

axb_status axb_memcpy_d2h(void *dst, const void *src, size_t bytes, axb_stream_t stream, int sync) {
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    if (bytes >= kStagedMin && sync && staged_enabled() && pageable(dst)) {
        // a synchronous read into pageable memory (Tensor::cpu()): drained through the pinned lanes by several threads; every
        // chunk has landed in `dst` when staged_copy returns
        AXB_TRY(staged_copy(dst, src, bytes, false, s, ctx->device));
        count_d2h(bytes);
        return AXB_OK;
    }
    if (bytes) {
        AXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        count_d2h(bytes);
    }
    if (sync)
        AXB_CUDA(cudaStreamSynchronize(s));
    return AXB_OK;
}
axb_status axb_memcpy_d2d(void *dst, const void *src, size_t bytes, axb_stream_t stream) {
    if (!bytes)
        return AXB_OK;
    cudaStream_t s;
    AXB_TRY(resolve_stream(stream, &s));
    gemm::invalidate_const_cache(dst, bytes);
    AXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
    count_d2d(bytes);
    return AXB_OK;
}
axb_status axb_memset(void *dst, int value, size_t bytes, axb_stream_t stream) {
    if (!bytes)
        return AXB_OK;
    cudaStream_t s;
    AXB_TRY(resolve_stream(stream, &s));
    gemm::invalidate_const_cache(dst, bytes);
    AXB_CUDA(cudaMemsetAsync(dst, value, bytes, s));
    return AXB_OK;
}

axb_status axb_get_transfer_stats(axb_transfer_stats *out) {
    AXB_REQUIRE(out, AXB_ERR_INVALID, "axb_get_transfer_stats: null out");
    out->h2d_bytes = g_xfer.h2d_bytes.load();
    out->h2d_calls = g_xfer.h2d_calls.load();
    out->d2h_bytes = g_xfer.d2h_bytes.load();
    out->d2h_calls = g_xfer.d2h_calls.load();
    out->d2d_bytes = g_xfer.d2d_bytes.load();
    out->d2d_calls = g_xfer.d2d_calls.load();
    return AXB_OK;
}

size_t axb_dtype_size(int32_t dtype) { return dtype_size(dtype); }

} // extern "C"
