// driver.cpp — see driver.h.
#include "driver.h"

#include <mutex>

namespace axb {
namespace drv {

namespace {
Api g_api;
std::once_flag g_once;
std::string g_why;

template <typename F> bool resolve(const char *name, F *out) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        g_why = std::string("driver entry point ") + name + " unavailable";
        return false;
    }
    *out = reinterpret_cast<F>(p);
    return true;
}
} // namespace

const Api *api() {
    std::call_once(g_once, [] {
        bool ok = resolve("cuModuleLoadData", &g_api.ModuleLoadData);
        ok = ok && resolve("cuModuleGetFunction", &g_api.ModuleGetFunction);
        ok = ok && resolve("cuLaunchKernel", &g_api.LaunchKernel);
        ok = ok && resolve("cuGetErrorString", &g_api.GetErrorString);
        ok = ok && resolve("cuFuncSetAttribute", &g_api.FuncSetAttribute);
        ok = ok && resolve("cuTensorMapEncodeTiled", &g_api.TensorMapEncodeTiled);
        ok = ok && resolve("cuOccupancyMaxActiveBlocksPerMultiprocessor", &g_api.OccupancyMaxActiveBlocksPerMultiprocessor);
        g_api.ok = ok;
    });
    if (!g_api.ok) {
        fail(AXB_ERR_CUDA, "CUDA driver API unavailable (%s)", g_why.c_str());
        return nullptr;
    }
    return &g_api;
}

const char *err_str(CUresult r) {
    const char *s = nullptr;
    if (g_api.GetErrorString)
        g_api.GetErrorString(r, &s);
    return s ? s : "unknown driver error";
}

} // namespace drv
} // namespace axb
