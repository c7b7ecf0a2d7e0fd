# CTA-pair GEMM: parity first (bounded by timeout), then the C4 shape with the pair kernel off / on
timeout 600 python -m pytest tests/test_matmul_gpu.py -m gpu -q 2>&1 | tail -15
AXB_GEMM_PAIR=0 timeout 120 python tools/perf_probe.py bmm_nt mm:bf16:8192:8192:8192
AXB_GEMM_PAIR=1 timeout 120 python tools/perf_probe.py bmm bmm_nt mm:bf16:8192:8192:8192
