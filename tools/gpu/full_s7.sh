set -x
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_s7.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_bench_s7.csv
