mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; tail -2 gpurun_out/r3_bench.err | cut -c1-300; cut -c1-300 gpurun_out/r3_bench.json
