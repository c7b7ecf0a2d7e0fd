set -x
timeout 900 python -m pytest tests/test_elementwise_gpu.py -q -x --timeout 600 -k "half_table" 2>&1 | tail -8
timeout 300 python tools/perf_probe.py bf16:relu bf16:exp bf16:tanh bf16:gelu bf16:sigmoid 2>&1 | tail -5
AXB_NO_HALF_LUT=1 timeout 300 python tools/perf_probe.py bf16:exp bf16:tanh bf16:gelu 2>&1 | tail -3
