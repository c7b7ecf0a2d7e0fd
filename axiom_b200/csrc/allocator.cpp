// allocator.cpp — stream-ordered caching device allocator.
//
// Backs make_storage(bytes, Device::GPU) (reference: src/tensor/storage.cpp:31-45, which throws on
// Linux today) and the graph executor's arena slots, which are shared_ptr<Storage> released from
// arbitrary host threads while kernels that read them may still be queued
// (src/graph/graph_executor.cpp:254-303, compiled_graph.hpp:121-141).
//
// Design (B200, 180 GB HBM3e):
//  * sizes are rounded to classes (512 B granules below 1 MiB, 2 MiB granules above). Small blocks are cached per exact
//    class. Large blocks live inside SEGMENTS (one cudaMalloc each): a cached block that is larger than the request is
//    SPLIT (the remainder stays cached), and a freed block is COALESCED with free neighbours of the same segment that
//    were freed on the same stream — so a freed 2 GiB block can serve four 512 MiB requests and grow back.
//  * free(ptr, stream) returns the block to the cache at once, owned by that stream: later work on the same stream is
//    ordered after everything that used the block, so no synchronisation is needed. A block is handed to a different
//    stream only through a cudaStreamWaitEvent on the event recorded at free time (device-side ordering, the host
//    never blocks). The cache is ONE ordered index by (size, address); a lookup walks at most kScan candidates from the
//    best fit, preferring the caller's stream — no scan of the whole cache.
//  * every block remembers its DEVICE. A free with stream == NULL (the Storage destructor, called from arbitrary host
//    threads whose current device may be another GPU) goes to the backend stream of the block's own device, and the
//    event is recorded with that device current.
//  * stream capture: while the stream is capturing, freed blocks are parked in a capture list that only later
//    allocations of the SAME capture reuse; they never return to the general cache, because replays of the captured
//    graph keep writing to them. Allocations made during capture never wait on events from outside the capture.
//  * on cudaErrorMemoryAllocation the cache is drained (after a device sync) and the malloc retried.
//  * one mutex per device; all entry points are thread-safe. axb_pool_reserve pre-warms a segment.
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>
#include <vector>

#include "axb_internal.h"
#include "gemm.h"

namespace axb {
namespace {

constexpr size_t kSmallGranule = 512;
constexpr size_t kLargeGranule = 2u << 20;
constexpr size_t kSmallLimit = 1u << 20;
constexpr int kScan = 32;

size_t round_size(size_t n) {
    if (n < kSmallLimit)
        return (n + kSmallGranule - 1) / kSmallGranule * kSmallGranule;
    return (n + kLargeGranule - 1) / kLargeGranule * kLargeGranule;
}

// free-time event shared by the pieces of a block that was split while cached
struct Ev {
    cudaEvent_t e = nullptr;
    int refs = 0;
};

struct Block {
    void *ptr = nullptr;
    size_t size = 0; // rounded size
    int device = 0;
    cudaStream_t stream = nullptr; // allocation stream while live, free stream while cached
    Ev *freed = nullptr;           // recorded on `stream` when it was freed (nullptr: nothing to wait for)
    Block *prev = nullptr, *next = nullptr; // neighbours inside the segment (large blocks only)
    bool cached = false;
    bool capture = false; // parked by a free during stream capture: never returns to the general cache
};

struct BySizeAddr {
    bool operator()(const Block *a, const Block *b) const {
        return a->size != b->size ? a->size < b->size : a->ptr < b->ptr;
    }
};

bool is_capturing(cudaStream_t s) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

struct DevicePool {
    std::mutex mu;
    int device = 0;
    std::unordered_map<void *, Block *> live;
    std::set<Block *, BySizeAddr> cached;                         // general cache
    std::multimap<std::pair<cudaStream_t, size_t>, Block *> park; // freed during capture, per capturing stream
    std::vector<cudaEvent_t> event_pool;
    axb_alloc_stats st{};

    Ev *new_event(cudaStream_t s) {
        cudaEvent_t e = nullptr;
        if (!event_pool.empty()) {
            e = event_pool.back();
            event_pool.pop_back();
        } else if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        if (cudaEventRecord(e, s) != cudaSuccess) {
            cudaGetLastError();
            event_pool.push_back(e);
            return nullptr;
        }
        Ev *ev = new Ev();
        ev->e = e;
        ev->refs = 1;
        return ev;
    }
    void release(Ev *&ev) {
        if (ev && --ev->refs == 0) {
            event_pool.push_back(ev->e);
            delete ev;
        }
        ev = nullptr;
    }

    // hands a cached block to `stream`; false if the free point cannot be ordered (no event could be recorded)
    void adopt(Block *b, cudaStream_t stream) {
        if (b->stream != stream && b->freed) {
            cudaStreamWaitEvent(stream, b->freed->e, 0);
            st.n_cross_stream++;
        }
        release(b->freed);
        b->stream = stream;
    }

    Block *take_cached(size_t size, cudaStream_t stream, bool same_stream_only) {
        Block key;
        key.size = size;
        key.ptr = nullptr;
        auto it = cached.lower_bound(&key);
        Block *other = nullptr, *pick = nullptr;
        const bool small = size < kSmallLimit;
        for (int n = 0; it != cached.end() && n < kScan; ++it, ++n) {
            Block *b = *it;
            if (small && b->size != size)
                break; // small blocks: exact class only (never split)
            if (!small && b->size < kSmallLimit)
                continue;
            if (b->stream == stream) {
                pick = b;
                break;
            }
            // a block freed on another stream is ordered through its free event (none = the free point was synchronised)
            if (!other && !same_stream_only)
                other = b;
            if (other && b->size > other->size)
                break; // a larger same-stream block is not worth the extra split
        }
        if (!pick)
            pick = other;
        if (!pick)
            return nullptr;
        cached.erase(pick);
        pick->cached = false;
        st.bytes_cached -= pick->size;
        // split: keep the tail cached under the stream (and free point) it was freed on
        if (!small && pick->size >= size + kLargeGranule) {
            Block *rem = new Block();
            rem->ptr = static_cast<char *>(pick->ptr) + size;
            rem->size = pick->size - size;
            rem->device = pick->device;
            rem->stream = pick->stream;
            rem->freed = pick->freed;
            if (rem->freed)
                rem->freed->refs++;
            rem->prev = pick;
            rem->next = pick->next;
            if (pick->next)
                pick->next->prev = rem;
            pick->next = rem;
            pick->size = size;
            rem->cached = true;
            cached.insert(rem);
            st.bytes_cached += rem->size;
            st.n_splits++;
        }
        adopt(pick, stream);
        return pick;
    }

    Block *take_parked(size_t size, cudaStream_t stream) {
        auto it = park.lower_bound({stream, size});
        if (it == park.end() || it->first.first != stream)
            return nullptr;
        if (it->first.second != size && (size < kSmallLimit || it->first.second > size + size / 4))
            return nullptr;
        Block *b = it->second;
        park.erase(it);
        return b;
    }

    // merges b with a cached neighbour freed on the same stream; `b` is not in the index yet
    void coalesce(Block *b) {
        for (int side = 0; side < 2; ++side) {
            Block *n = side == 0 ? b->next : b->prev;
            if (!n || !n->cached || n->capture || n->stream != b->stream)
                continue;
            cached.erase(n);
            st.bytes_cached -= n->size;
            release(n->freed); // b's event is later on the same stream: it covers n's free point
            if (side == 0) {
                b->size += n->size;
                b->next = n->next;
                if (n->next)
                    n->next->prev = b;
            } else {
                b->ptr = n->ptr;
                b->size += n->size;
                b->prev = n->prev;
                if (n->prev)
                    n->prev->next = b;
            }
            delete n;
            st.n_merges++;
        }
    }

    void drain_locked() {
        cudaDeviceSynchronize();
        // everything queued has finished: free points no longer matter, neighbours merge across streams
        std::vector<Block *> blocks(cached.begin(), cached.end());
        cached.clear();
        std::set<Block *> dead;
        for (Block *b : blocks) {
            if (dead.count(b))
                continue;
            release(b->freed);
            // forward merges on every surviving block cover every mergeable pair, whatever the visiting order
            while (b->next && b->next->cached && !b->next->capture) {
                Block *n = b->next;
                release(n->freed);
                b->size += n->size;
                b->next = n->next;
                if (n->next)
                    n->next->prev = b;
                dead.insert(n);
            }
        }
        for (Block *b : blocks) {
            if (dead.count(b))
                continue;
            if (!b->prev && !b->next) { // a whole segment
                cudaFree(b->ptr);
                st.bytes_cached -= b->size;
                dead.insert(b);
            } else {
                cached.insert(b); // part of a segment that still has live blocks
            }
        }
        for (Block *b : dead)
            delete b;
    }
};

std::mutex g_pools_mu;
std::vector<DevicePool *> g_pools;
// which pool owns a live pointer (a free may come from a thread bound to another device)
std::mutex g_owner_mu;
std::unordered_map<void *, int> g_owner;

DevicePool *pool_for(int dev) {
    std::lock_guard<std::mutex> lk(g_pools_mu);
    if ((int)g_pools.size() <= dev)
        g_pools.resize(dev + 1, nullptr);
    if (!g_pools[dev]) {
        g_pools[dev] = new DevicePool();
        g_pools[dev]->device = dev;
    }
    return g_pools[dev];
}

struct DeviceGuard { // makes `dev` current for the scope (events must be recorded with their device current)
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev)
            switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched)
            cudaSetDevice(prev);
    }
};

axb_status new_segment(DevicePool *p, size_t size, Block **out) {
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, size);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError(); // clear
        p->drain_locked();
        e = cudaMalloc(&d, size);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? AXB_ERR_OOM : AXB_ERR_CUDA, "axb_alloc: cudaMalloc(%zu) failed: %s", size,
                    cudaGetErrorString(e));
    }
    p->st.n_cuda_malloc++;
    Block *b = new Block();
    b->ptr = d;
    b->size = size;
    b->device = p->device;
    *out = b;
    return AXB_OK;
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
    const size_t size = round_size(bytes);
    const bool capturing = is_capturing(s);
    std::lock_guard<std::mutex> lk(p->mu);
    p->st.n_allocs++;
    Block *b = capturing ? p->take_parked(size, s) : nullptr;
    if (!b)
        b = p->take_cached(size, s, /*same_stream_only=*/capturing);
    if (b) {
        p->st.n_cache_hits++;
    } else {
        AXB_TRY(new_segment(p, size, &b));
    }
    b->stream = s;
    p->live[b->ptr] = b;
    {
        std::lock_guard<std::mutex> ol(g_owner_mu);
        g_owner[b->ptr] = p->device;
    }
    p->st.bytes_in_use += b->size;
    if (p->st.bytes_in_use > p->st.peak_bytes_in_use)
        p->st.peak_bytes_in_use = p->st.bytes_in_use;
    *ptr = b->ptr;
    return AXB_OK;
}

axb_status axb_free(void *ptr, axb_stream_t stream) {
    if (!ptr)
        return AXB_OK;
    int dev = -1;
    {
        std::lock_guard<std::mutex> ol(g_owner_mu);
        auto it = g_owner.find(ptr);
        if (it != g_owner.end()) {
            dev = it->second;
            g_owner.erase(it);
        }
    }
    if (dev < 0)
        return fail(AXB_ERR_INVALID, "axb_free: pointer %p was not allocated by axb_alloc", ptr);
    DevicePool *p = pool_for(dev);
    // stream == NULL means "the backend stream of the block's OWN device" — the calling thread may be bound to another
    // GPU or to none (Storage destructors run on arbitrary host threads)
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!s) {
        DeviceCtx *octx = ctx_of_device(dev);
        if (!octx)
            return fail(AXB_ERR_CUDA, "axb_free: device %d has no context", dev);
        s = octx->stream;
    }
    // operand forms derived from this block (cached weight splits) die with it — before the block can be handed out again
    if (gemm::const_cache_active()) {
        size_t blk = 0;
        {
            std::lock_guard<std::mutex> lk(p->mu);
            auto it = p->live.find(ptr);
            if (it != p->live.end())
                blk = it->second->size;
        }
        if (blk)
            gemm::invalidate_const_cache(ptr, blk);
    }
    DeviceGuard guard(dev);
    const bool capturing = is_capturing(s);
    std::lock_guard<std::mutex> lk(p->mu);
    auto it = p->live.find(ptr);
    if (it == p->live.end())
        return fail(AXB_ERR_INVALID, "axb_free: pointer %p is not live on device %d", ptr, dev);
    Block *b = it->second;
    p->live.erase(it);
    p->st.bytes_in_use -= b->size;
    // If the caller frees on a different stream than it allocated on, the free stream is the one whose queued work may
    // still touch the block: cache it under that stream.
    b->stream = s;
    if (capturing) {
        b->capture = true;
        p->park.emplace(std::make_pair(s, b->size), b);
        return AXB_OK;
    }
    if (b->capture) { // allocated from the capture list, freed after the capture ended: stays out of the general cache
        p->park.emplace(std::make_pair(s, b->size), b);
        return AXB_OK;
    }
    b->freed = p->new_event(s);
    if (!b->freed) {
        // no event could be recorded: make the free point explicit on the host instead of risking an unordered reuse
        cudaStreamSynchronize(s);
    }
    if (b->size >= kSmallLimit)
        p->coalesce(b);
    b->cached = true;
    p->cached.insert(b);
    p->st.bytes_cached += b->size;
    return AXB_OK;
}

axb_status axb_pool_reserve(size_t bytes) {
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(nullptr, &s, &ctx));
    if (bytes < kSmallLimit)
        return AXB_OK;
    DevicePool *p = pool_for(ctx->device);
    const size_t size = round_size(bytes);
    std::lock_guard<std::mutex> lk(p->mu);
    Block *b = nullptr;
    AXB_TRY(new_segment(p, size, &b));
    b->stream = s;
    b->cached = true;
    p->cached.insert(b);
    p->st.bytes_cached += b->size;
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
