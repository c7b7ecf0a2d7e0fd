// allocator.cpp — stream-ordered caching device allocator.
//
// Backs make_storage(bytes, Device::GPU) (reference: src/tensor/storage.cpp:31-45, which throws on
// Linux today) and the graph executor's arena slots, which are shared_ptr<Storage> released from
// arbitrary host threads while kernels that read them may still be queued
// (src/graph/graph_executor.cpp:254-303, compiled_graph.hpp:121-141).
//
// Design (B200, 180 GB HBM3e):
//  * sizes are rounded to classes (512 B granules below 1 MiB, 2 MiB granules above) so freed blocks
//    are re-usable; cudaMalloc is only hit on a class miss.
//  * free(ptr, stream) returns the block to the free list of THAT stream at once — later work on the
//    same stream is ordered after everything that used the block, so no synchronisation is needed.
//  * a block is handed to a different stream only through a cudaStreamWaitEvent on an event recorded
//    at free time (device-side ordering, the host never blocks).
//  * on cudaErrorMemoryAllocation the cache is drained (after a device sync) and the malloc retried.
//  * one mutex per device; all entry points are thread-safe.
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "axb_internal.h"

namespace axb {
namespace {

constexpr size_t kSmallGranule = 512;
constexpr size_t kLargeGranule = 2u << 20;
constexpr size_t kSmallLimit = 1u << 20;

size_t round_size(size_t n) {
    if (n < kSmallLimit)
        return (n + kSmallGranule - 1) / kSmallGranule * kSmallGranule;
    return (n + kLargeGranule - 1) / kLargeGranule * kLargeGranule;
}

struct Block {
    void *ptr = nullptr;
    size_t size = 0;           // rounded size
    cudaStream_t stream = nullptr; // stream that owns it while cached
    cudaEvent_t freed = nullptr;   // recorded on `stream` when it was freed
};

struct DevicePool {
    std::mutex mu;
    // live allocations: ptr -> block
    std::unordered_map<void *, Block> live;
    // cached blocks: (stream, size) ordered so lower_bound finds best fit within a stream
    std::multimap<std::pair<cudaStream_t, size_t>, Block> cached;
    std::vector<cudaEvent_t> event_pool;
    axb_alloc_stats st{};

    cudaEvent_t get_event() {
        if (!event_pool.empty()) {
            cudaEvent_t e = event_pool.back();
            event_pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        return e;
    }
    void put_event(cudaEvent_t e) {
        if (e)
            event_pool.push_back(e);
    }

    // returns true and fills b if a cached block fits
    bool take_cached(size_t size, cudaStream_t stream, Block *b) {
        // 1. same stream, best fit (allow up to 25% slack for large blocks, exact class for small ones)
        auto it = cached.lower_bound({stream, size});
        if (it != cached.end() && it->first.first == stream) {
            size_t have = it->first.second;
            if (have == size || (size >= kSmallLimit && have <= size + size / 4)) {
                *b = it->second;
                cached.erase(it);
                put_event(b->freed);
                b->freed = nullptr;
                return true;
            }
        }
        // 2. any other stream, exact class: order the new stream after the free point
        for (auto jt = cached.begin(); jt != cached.end(); ++jt) {
            if (jt->first.second == size && jt->first.first != stream) {
                *b = jt->second;
                cached.erase(jt);
                if (b->freed) {
                    cudaStreamWaitEvent(stream, b->freed, 0);
                    put_event(b->freed);
                    b->freed = nullptr;
                }
                b->stream = stream;
                return true;
            }
        }
        return false;
    }

    void drain_locked() {
        cudaDeviceSynchronize();
        for (auto &kv : cached) {
            cudaFree(kv.second.ptr);
            st.bytes_cached -= kv.second.size;
            put_event(kv.second.freed);
        }
        cached.clear();
    }
};

std::mutex g_pools_mu;
std::vector<DevicePool *> g_pools;

DevicePool *pool_for(int dev) {
    std::lock_guard<std::mutex> lk(g_pools_mu);
    if ((int)g_pools.size() <= dev)
        g_pools.resize(dev + 1, nullptr);
    if (!g_pools[dev])
        g_pools[dev] = new DevicePool();
    return g_pools[dev];
}

} // namespace
} // namespace axb

using namespace axb;

extern "C" {

axb_status axb_alloc(size_t bytes, axb_stream_t stream, void **ptr) {
    AXB_REQUIRE(ptr, AXB_ERR_INVALID, "axb_alloc: null out");
    *ptr = nullptr;
    if (bytes == 0)
        return AXB_OK; // size_bytes == 0 is legal
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    DevicePool *p = pool_for(ctx->device);
    size_t size = round_size(bytes);
    std::lock_guard<std::mutex> lk(p->mu);
    Block b;
    p->st.n_allocs++;
    if (p->take_cached(size, s, &b)) {
        p->st.n_cache_hits++;
        p->st.bytes_cached -= b.size;
    } else {
        void *d = nullptr;
        cudaError_t e = cudaMalloc(&d, size);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError(); // clear
            p->drain_locked();
            e = cudaMalloc(&d, size);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(e == cudaErrorMemoryAllocation ? AXB_ERR_OOM : AXB_ERR_CUDA,
                        "axb_alloc: cudaMalloc(%zu) failed: %s", size, cudaGetErrorString(e));
        }
        p->st.n_cuda_malloc++;
        b.ptr = d;
        b.size = size;
        b.stream = s;
    }
    b.stream = s;
    p->live[b.ptr] = b;
    p->st.bytes_in_use += b.size;
    if (p->st.bytes_in_use > p->st.peak_bytes_in_use)
        p->st.peak_bytes_in_use = p->st.bytes_in_use;
    *ptr = b.ptr;
    return AXB_OK;
}

axb_status axb_free(void *ptr, axb_stream_t stream) {
    if (!ptr)
        return AXB_OK;
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    // the block may belong to another device's pool (multi-GPU host thread): search the owner
    cudaPointerAttributes attr{};
    int dev = ctx->device;
    if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice)
        dev = attr.device;
    else
        cudaGetLastError();
    DevicePool *p = pool_for(dev);
    std::lock_guard<std::mutex> lk(p->mu);
    auto it = p->live.find(ptr);
    if (it == p->live.end())
        return fail(AXB_ERR_INVALID, "axb_free: pointer %p was not allocated by axb_alloc", ptr);
    Block b = it->second;
    p->live.erase(it);
    p->st.bytes_in_use -= b.size;
    // If the caller frees on a different stream than it allocated on, the free stream is the one whose
    // queued work may still touch the block: cache it under that stream.
    b.stream = s;
    b.freed = p->get_event();
    if (b.freed)
        cudaEventRecord(b.freed, s);
    p->cached.emplace(std::make_pair(b.stream, b.size), b);
    p->st.bytes_cached += b.size;
    return AXB_OK;
}

axb_status axb_empty_cache(void) {
    DeviceCtx *ctx = current_ctx();
    if (!ctx)
        return AXB_ERR_CUDA;
    DevicePool *p = pool_for(ctx->device);
    std::lock_guard<std::mutex> lk(p->mu);
    p->drain_locked();
    return AXB_OK;
}

axb_status axb_get_alloc_stats(axb_alloc_stats *out) {
    AXB_REQUIRE(out, AXB_ERR_INVALID, "axb_get_alloc_stats: null out");
    DeviceCtx *ctx = current_ctx();
    if (!ctx)
        return AXB_ERR_CUDA;
    DevicePool *p = pool_for(ctx->device);
    std::lock_guard<std::mutex> lk(p->mu);
    *out = p->st;
    return AXB_OK;
}

} // extern "C"
