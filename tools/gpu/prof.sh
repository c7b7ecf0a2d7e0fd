set -x
REPS=1 ncu --set full --clock-control none -k regex:"axb_ew|rowwide|rowcluster|argreduce_rows|reduce_rows|gemm_tc|split_f32" -o /tmp/r1s6_all python tools/prof_targets.py c2 sum argmax softmax layer_norm bf16 bmm mm32 > gpurun_out/ncu_all.log 2>&1
tail -2 gpurun_out/ncu_all.log
python tools/ncu_summary.py /tmp/r1s6_all.ncu-rep > gpurun_out/ncu_summary_s6.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/ncu_summary_s6.txt gpurun_out/launches_bench.csv
