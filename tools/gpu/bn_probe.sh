python tools/perf_probe.py bmm8
AXB_GEMM_BN=128 python tools/perf_probe.py bmm8
AXB_GEMM_BN=64 python tools/perf_probe.py bmm8
