set -x
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; tail -2 gpurun_out/r3_bench.err | cut -c1-300; cut -c1-600 gpurun_out/r3_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r3_bench_ref.json 2>> gpurun_out/r3_bench.err; cut -c1-300 gpurun_out/r3_bench_ref.json
