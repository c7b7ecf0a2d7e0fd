// jit.cpp — see jit.h. NVRTC is loaded lazily (dlopen) so the library itself loads on machines without it; a
// program whose specialised kernel is neither pre-built nor compilable fails loudly (no interpreter, no CPU path).
#include "jit.h"

#include "driver.h"

#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <tuple>
#include <mutex>
#include <sstream>
#include <vector>

namespace axb {
namespace jit {

namespace {

const char kPrelude[] =
#include "ew_prelude_embed.inc"
    ;

const char kCacheTag[] = "axb-jit-v3 nvrtc sm_100a -std=c++17 --fmad=false";

uint64_t fnv1a(const void *data, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

std::string cache_dir() {
    static std::string dir = [] {
        const char *env = getenv("AXB_KERNEL_CACHE");
        if (env && *env)
            return std::string(env);
        Dl_info info{};
        std::string d = ".";
        if (dladdr(reinterpret_cast<const void *>(&fnv1a), &info) && info.dli_fname) {
            d = info.dli_fname;
            size_t slash = d.find_last_of('/');
            d = slash == std::string::npos ? "." : d.substr(0, slash);
        }
        return d + "/kernel_cache";
    }();
    return dir;
}

struct Nvrtc {
    void *handle = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcGetErrorString) err = nullptr;
    std::string why;
};

Nvrtc &nvrtc() {
    static Nvrtc n = [] {
        Nvrtc r;
        const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char *nm : names) {
            r.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (r.handle)
                break;
        }
        if (!r.handle) {
            r.why = std::string("libnvrtc not found: ") + (dlerror() ? dlerror() : "?");
            return r;
        }
#define AXB_SYM(field, name)                                                                            \
    r.field = reinterpret_cast<decltype(r.field)>(dlsym(r.handle, #name));                              \
    if (!r.field)                                                                                       \
        r.why = "libnvrtc lacks " #name;
        AXB_SYM(create, nvrtcCreateProgram)
        AXB_SYM(compile, nvrtcCompileProgram)
        AXB_SYM(cubin_size, nvrtcGetCUBINSize)
        AXB_SYM(cubin, nvrtcGetCUBIN)
        AXB_SYM(log_size, nvrtcGetProgramLogSize)
        AXB_SYM(log, nvrtcGetProgramLog)
        AXB_SYM(destroy, nvrtcDestroyProgram)
        AXB_SYM(err, nvrtcGetErrorString)
#undef AXB_SYM
        return r;
    }();
    return n;
}

axb_status compile_cubin(const std::string &source, std::vector<char> *out) {
    Nvrtc &n = nvrtc();
    if (!n.why.empty() || !n.handle)
        return fail(AXB_ERR_UNSUPPORTED, "kernel specialisation needs NVRTC and it is unavailable (%s); no fallback path exists",
                    n.why.c_str());
    nvrtcProgram prog;
    nvrtcResult r = n.create(&prog, source.c_str(), "axb_kernel.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS)
        return fail(AXB_ERR_CUDA, "nvrtcCreateProgram: %s", n.err(r));
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "-lineinfo"};
    r = n.compile(prog, 4, opts);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        n.log_size(prog, &ls);
        std::string log(ls, '\0');
        if (ls)
            n.log(prog, &log[0]);
        n.destroy(&prog);
        if (log.size() > 900)
            log.resize(900);
        return fail(AXB_ERR_CUDA, "nvrtcCompileProgram: %s\n%s", n.err(r), log.c_str());
    }
    size_t sz = 0;
    n.cubin_size(prog, &sz);
    out->resize(sz);
    n.cubin(prog, out->data());
    n.destroy(&prog);
    return AXB_OK;
}

bool read_file(const std::string &path, std::vector<char> *out) {
    std::ifstream f(path, std::ios::binary);
    if (!f)
        return false;
    f.seekg(0, std::ios::end);
    std::streamoff n = f.tellg();
    if (n <= 0)
        return false;
    f.seekg(0);
    out->resize((size_t)n);
    f.read(out->data(), n);
    return (bool)f;
}

void write_file_atomic(const std::string &path, const std::vector<char> &data) {
    mkdir(cache_dir().c_str(), 0755); // best effort; an unwritable cache only costs a recompile
    char tmp[512];
    snprintf(tmp, sizeof tmp, "%s.tmp.%d", path.c_str(), (int)getpid());
    {
        std::ofstream f(tmp, std::ios::binary);
        if (!f)
            return;
        f.write(data.data(), (std::streamsize)data.size());
        if (!f)
            return;
    }
    rename(tmp, path.c_str());
}

std::mutex g_mu;
// Kernel identity = (FNV-1a of the source, a second independent hash, source length): a 64-bit collision alone can no longer
// alias two programs, neither in memory nor on disk.
struct SourceId {
    uint64_t h1, h2, len;
    bool operator<(const SourceId &o) const { return std::tie(h1, h2, len) < std::tie(o.h1, o.h2, o.len); }
};
std::map<std::pair<SourceId, int>, Kernel> g_kernels; // (source id, device) -> function
Stats g_stats{0, 0, 0};

// `body` is the program-specific tail of the translation unit; the shared prelude (ew_prelude.cuh + hwy_math.cuh, ~45 KB)
// is hashed ONCE per process and only seeds the per-call hashes, so a cache lookup costs a hash of ~1 KB, not of the
// whole source (which dominated the host time of a fused-chain launch: two passes over 45 KB per call).
SourceId source_id(const std::string &body) {
    static const SourceId prelude = [] {
        const std::string &p = ew_prelude_text();
        SourceId id;
        id.h1 = fnv1a(p.data(), p.size(), fnv1a(kCacheTag, sizeof kCacheTag));
        id.h2 = fnv1a(p.data(), p.size(), 0x9E3779B97F4A7C15ull ^ p.size());
        id.len = p.size();
        return id;
    }();
    SourceId id;
    id.h1 = fnv1a(body.data(), body.size(), prelude.h1);
    id.h2 = fnv1a(body.data(), body.size(), prelude.h2 ^ body.size());
    id.len = prelude.len + body.size();
    return id;
}
std::string cache_path(const SourceId &id) {
    char name[64];
    snprintf(name, sizeof name, "/%016llx.axbk", (unsigned long long)id.h1);
    return cache_dir() + name;
}

// Cache file = header + cubin. The header names the source it was compiled from (length + second hash) and the payload
// (length + hash), so a truncated, corrupt or colliding file is detected, deleted and recompiled instead of being loaded.
struct CacheHeader {
    char magic[8]; // "AXBKRN01"
    uint64_t source_len, source_h2, cubin_len, cubin_hash;
};
const char kMagic[8] = {'A', 'X', 'B', 'K', 'R', 'N', '0', '1'};

bool read_cached(const std::string &path, const SourceId &id, std::vector<char> *bin) {
    std::vector<char> raw;
    if (!read_file(path, &raw))
        return false;
    CacheHeader h;
    bool ok = raw.size() > sizeof h;
    if (ok) {
        memcpy(&h, raw.data(), sizeof h);
        ok = memcmp(h.magic, kMagic, 8) == 0 && h.source_len == id.len && h.source_h2 == id.h2 &&
             h.cubin_len == raw.size() - sizeof h && h.cubin_hash == fnv1a(raw.data() + sizeof h, h.cubin_len, 0x1234567ull);
    }
    if (!ok) {
        remove(path.c_str()); // evict: the next line recompiles and rewrites it
        return false;
    }
    bin->assign(raw.begin() + sizeof h, raw.end());
    return true;
}

void write_cached(const std::string &path, const SourceId &id, const std::vector<char> &bin) {
    CacheHeader h;
    memcpy(h.magic, kMagic, 8);
    h.source_len = id.len;
    h.source_h2 = id.h2;
    h.cubin_len = bin.size();
    h.cubin_hash = fnv1a(bin.data(), bin.size(), 0x1234567ull);
    std::vector<char> raw(sizeof h + bin.size());
    memcpy(raw.data(), &h, sizeof h);
    memcpy(raw.data() + sizeof h, bin.data(), bin.size());
    write_file_atomic(path, raw);
}

// cubin bytes for `source`: disk cache, else NVRTC (+ store)
axb_status cubin_for(const std::string &body, const SourceId &id, std::vector<char> *bin, bool force_compile = false) {
    const std::string path = cache_path(id);
    if (!force_compile && read_cached(path, id, bin)) {
        g_stats.disk_hits++;
        return AXB_OK;
    }
    AXB_TRY(compile_cubin(ew_prelude_text() + body, bin));
    g_stats.compiles++;
    write_cached(path, id, *bin);
    return AXB_OK;
}

} // namespace

const std::string &ew_prelude_text() {
    static const std::string s(kPrelude);
    return s;
}

axb_status precompile(const std::string &source) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<char> bin;
    return cubin_for(source, source_id(source), &bin);
}

axb_status get_kernel(const std::string &source, const char *name, int device, Kernel *out) {
    const SourceId id = source_id(source);
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_kernels.find({id, device});
    if (it != g_kernels.end()) {
        g_stats.mem_hits++;
        *out = it->second;
        return AXB_OK;
    }
    std::vector<char> bin;
    AXB_TRY(cubin_for(source, id, &bin));
    const drv::Api *d = drv::api();
    if (!d)
        return AXB_ERR_CUDA;
    CUmodule mod;
    CUresult r = d->ModuleLoadData(&mod, bin.data());
    if (r != CUDA_SUCCESS) {
        // a cache entry that passes its checksums but does not load (other driver / toolkit): compile afresh, once
        AXB_TRY(cubin_for(source, id, &bin, /*force_compile=*/true));
        r = d->ModuleLoadData(&mod, bin.data());
    }
    if (r != CUDA_SUCCESS)
        return fail(AXB_ERR_CUDA, "cuModuleLoadData failed: %s", drv::err_str(r));
    Kernel k;
    r = d->ModuleGetFunction(&k.fn, mod, name);
    if (r != CUDA_SUCCESS)
        return fail(AXB_ERR_CUDA, "specialised kernel '%s' not found in module", name);
    if (d->OccupancyMaxActiveBlocksPerMultiprocessor)
        d->OccupancyMaxActiveBlocksPerMultiprocessor(&k.blocks_per_sm, k.fn, 256, 0);
    g_kernels[{id, device}] = k;
    *out = k;
    return AXB_OK;
}

axb_status launch(const Kernel &k, unsigned grid, unsigned block, unsigned smem, cudaStream_t s, void **args) {
    const drv::Api *d = drv::api();
    if (!d)
        return AXB_ERR_CUDA;
    CUresult r = d->LaunchKernel(k.fn, grid, 1, 1, block, 1, 1, smem, (CUstream)s, args, nullptr);
    if (r != CUDA_SUCCESS)
        return fail(AXB_ERR_CUDA, "cuLaunchKernel failed: %s", drv::err_str(r));
    return AXB_OK;
}

Stats stats() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_stats;
}

} // namespace jit
} // namespace axb
