set -x
timeout 1200 python -m pytest tests/test_spatial_gpu.py -q -x --timeout 600 2>&1 | tail -3
timeout 300 python tools/perf_probe.py conv:bf16:32,256,56,56:256,256,3,3:1:1 conv:bf16:32,256,56,56:512,256,1,1:1:0 conv:bf16:32,512,14,14:512,512,3,3:1:1 2>&1 | tail -3
AXB_CONV_NO_DIRECT=1 timeout 300 python tools/perf_probe.py conv:bf16:32,256,56,56:512,256,1,1:1:0 2>&1 | tail -1
