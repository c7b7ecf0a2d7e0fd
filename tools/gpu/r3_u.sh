mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spatial_gpu.py tests/test_matmul_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/convt_probe.py | tail -1 | tee -a gpurun_out/r3_convt_probe.json
