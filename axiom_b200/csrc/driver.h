// driver.h — the few CUDA driver-API entry points the backend needs (module loading for specialised kernels,
// TMA descriptor encoding), resolved through the runtime (cudaGetDriverEntryPoint) so libaxb.so does not link
// libcuda directly and still loads on a machine without a driver (the CPU-side build / export checks).
#pragma once
#include <cuda.h>

#include "axb_internal.h"

namespace axb {
namespace drv {

struct Api {
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void **, void **) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    bool ok = false;
};

// nullptr + error set when the driver is unavailable
const Api *api();
const char *err_str(CUresult r);

} // namespace drv
} // namespace axb
