set -x
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -3 gpurun_out/r2_bench_c.err; wc -c gpurun_out/r2_bench_c.json
