set -x
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
REPS=1 timeout 600 ncu --set full --clock-control none -k regex:"gemm_tc_pair" -o /tmp/r1s7_pair python tools/prof_targets.py bmm > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log
python tools/ncu_summary.py /tmp/r1s7_pair.ncu-rep > gpurun_out/ncu_summary_s7_pair.txt 2>&1
cp /tmp/r1s7_pair.ncu-rep gpurun_out/ 2>/dev/null
