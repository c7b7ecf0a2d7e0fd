timeout 1500 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -4
P="c3:softmax:f32:16384:8192 c3:softmax:f32:32768:4096 c3:softmax:bf16:16384:16384 c3:softmax:bf16:32768:8192 c3:layer_norm_noaffine:f32:16384:8192 c3:layer_norm:f32:32768:4096 c3:softmax:f32 c3:softmax:f32:16384:16384"
python tools/perf_probe.py $P
AXB_WIDE_ONE_ROW_PER_PASS=1 python tools/perf_probe.py c3:softmax:f32:16384:8192 c3:softmax:f32:32768:4096 c3:softmax:bf16:16384:16384 c3:softmax:f32:16384:16384
