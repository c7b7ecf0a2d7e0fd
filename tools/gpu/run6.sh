set -x
timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -5
P="c3:layer_norm:f32 c3:layer_norm:bf16 c3:rms_norm:f32 c3:rms_norm:bf16 c3:layer_norm:f32:16384:8192 c3:layer_norm:bf16:16384:16384 c3:layer_norm:f32:8191:32768"
timeout 300 python tools/perf_probe.py $P
AXB_CLUSTER_ROWS_PER_PASS=1 timeout 300 python tools/perf_probe.py c3:layer_norm:bf16 c3:rms_norm:bf16
