set -x
REPS=1 ncu --set full --clock-control none -k regex:"rowwide|argreduce_rows" -o gpurun_out/r1s6_rows python tools/prof_targets.py argmax softmax layer_norm bf16 > gpurun_out/ncu_rows.log 2>&1
tail -3 gpurun_out/ncu_rows.log
ls -la gpurun_out
