mkdir -p gpurun_out
T="timeout 300 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q -k test_layer_and_rms_norm_wide_and_cluster_f32"
for cfg in "X=1" "AXB_CLUSTER_NO_PIECES=1" "AXB_NO_STAGED_COPY=1"; do
  for i in 1 2 3 4 5 6; do echo "$cfg run $i: $(env $cfg $T 2>&1 | tail -1)"; done
done
