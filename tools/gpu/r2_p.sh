set -x
timeout 1200 python -m pytest tests/test_spatial_gpu.py tests/test_matmul_gpu.py tests/test_attention_gpu.py -q -x --timeout 600 2>&1 | tail -5
timeout 300 python tools/perf_probe.py conv:f32:32,64,56,56:64,64,3,3:1:1 conv:bf16:32,64,56,56:64,64,3,3:1:1 conv:bf16:32,256,56,56:256,256,3,3:1:1 conv:bf16:32,128,28,28:256,128,3,3:2:1 conv:bf16:32,512,14,14:512,512,3,3:1:1 conv:f16:8,64,224,224:64,64,3,3:1:1 2>&1 | tail -6
