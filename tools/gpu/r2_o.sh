set -x
timeout 900 python -m pytest tests/test_reduce_gpu.py -q -x --timeout 600 -k "last_axis or unfused_cases" 2>&1 | tail -12
timeout 300 axiom_b200/host/_build/api_parity 2>&1 | tail -1 | cut -c1-1500
timeout 900 python -m pytest tests/test_axiom_cpp.py -q -x --timeout 600 -k "fusion or reductions or lazy or graph_compiler or plan_cache" 2>&1 | tail -4
timeout 300 python tools/perf_probe.py conv:f32:32,64,56,56:64,64,3,3:1:1 conv:bf16:32,64,56,56:64,64,3,3:1:1 conv:bf16:32,256,56,56:256,256,3,3:1:1 conv:f32:8,80,3000:384,80,3:1:1:gelu conv:f32:8,384,3000:384,384,3:2:1:gelu conv:bf16:32,128,28,28:256,128,3,3:2:1 2>&1 | tail -6
AXB_CONV_NO_TC=1 timeout 300 python tools/perf_probe.py conv:f32:32,64,56,56:64,64,3,3:1:1 2>&1 | tail -1
