set -x
python -m pytest tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -5
python tools/perf_probe.py c3:argmax:f32 c3:argmax:bf16 c3:sum:f32 c3:sum:bf16 c3:argmax:f32:65536:4096 c3:argmax:f32:64:4194304
ncu --set full --clock-control none --import-source on -k regex:"rowwide|argreduce_rows" -o gpurun_out/r1s6_rows python tools/prof_targets.py argmax softmax layer_norm bf16 > gpurun_out/ncu_rows.log 2>&1
tail -3 gpurun_out/ncu_rows.log
