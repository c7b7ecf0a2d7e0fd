# CTA-pair GEMM: parity first (bounded by timeout), then the C4 shape with the pair kernel off / on
timeout 600 python -m pytest tests/test_matmul_gpu.py -m gpu -q 2>&1 | tail -5
timeout 120 python tools/perf_probe.py bmm bmm_nt mm:bf16:8192:8192:8192 mm:bf16:16384:16384:4096 mm:bf16:4096:4096:4096
