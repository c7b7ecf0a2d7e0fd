mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spatial_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; tail -2 gpurun_out/r3_bench.err | cut -c1-300; cut -c1-200 gpurun_out/r3_bench.json
