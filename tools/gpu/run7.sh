REPS=1 ncu --set full --clock-control none --import-source on -k regex:"rowcluster" -c 2 -o gpurun_out/r1s6_cluster python tools/prof_targets.py layer_norm bf16 > gpurun_out/ncu_cluster.log 2>&1
tail -3 gpurun_out/ncu_cluster.log
ls -la gpurun_out
