// jit.h — kernel specialisation service: CUDA C++ text -> sm_100a cubin (NVRTC) -> CUfunction, with an
// in-memory cache per device and an in-tree cubin cache on disk (axiom_b200/kernel_cache/<hash>.axbk: header + cubin).
//
// B200 counterpart of the reference's per-signature MPSGraph caches (src/backends/metal/mpsgraph_fused.mm:373-377,
// 256 entries; mpsgraph_graph_compiler.mm:756-765, 128 entries): there the Metal backend builds and compiles one
// MPSGraph per fused-chain signature; here one specialised CUDA kernel.
#pragma once
#include <cuda.h>

#include <string>

#include "axb_internal.h"

namespace axb {
namespace jit {

struct Kernel {
    CUfunction fn = nullptr;
    int blocks_per_sm = 0; // resident 256-thread CTAs per SM (occupancy query at load time)
};

// text of ew_prelude.cuh + hwy_math.cuh embedded at build time
const std::string &ew_prelude_text();

// Returns the kernel `name` of the translation unit  ew_prelude_text() + body  for `device` (compiling / loading it on
// first use). The cache key is a hash of `body` seeded with the prelude's hash: a lookup never touches the prelude text.
axb_status get_kernel(const std::string &body, const char *name, int device, Kernel *out);
// Compile-only: make sure the cubin of  ew_prelude_text() + body  is in the on-disk cache. Needs no GPU.
axb_status precompile(const std::string &body);
axb_status launch(const Kernel &k, unsigned grid, unsigned block, unsigned smem, cudaStream_t s, void **args);

struct Stats {
    uint64_t mem_hits, disk_hits, compiles;
};
Stats stats();

} // namespace jit
} // namespace axb
