set -x
timeout 1200 python -m pytest tests/test_elementwise_gpu.py tests/test_matmul_gpu.py -q -x --timeout 600 2>&1 | tail -3
timeout 300 python tools/perf_probe.py gelu sigmoid exp 2>&1 | tail -3
