set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err; cut -c1-1200 gpurun_out/r2_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench.err; cut -c1-300 gpurun_out/r2_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_launches_bench.csv
cd axiom_b200/host/_build
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /root/repo/gpurun_out/r2_whisper_launches_b8.csv ./whisper_encoder 8 1 0 > /root/repo/gpurun_out/whisper_ncu.log 2>&1
tail -2 /root/repo/gpurun_out/whisper_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /root/repo/gpurun_out/r2_whisper_launches_b1.csv ./whisper_encoder 1 1 0 > /root/repo/gpurun_out/whisper_ncu1.log 2>&1
