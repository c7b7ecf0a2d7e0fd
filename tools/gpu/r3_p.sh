mkdir -p gpurun_out
tools/_bin/erf_equiv | tee gpurun_out/r3_erf_equiv.json
timeout 900 python -m pytest tests/test_elementwise_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/perf_probe.py gelu erf exp 2>&1 | tail -4 | cut -c1-600
