set -x
python -m pytest tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -5
P="c3:softmax:f32 c3:softmax:bf16 c3:log_softmax:f32 c3:layer_norm:f32 c3:layer_norm:bf16 c3:layer_norm_noaffine:f32 c3:layer_norm_noaffine:bf16 c3:rms_norm:f32 c3:rms_norm:bf16 c3:softmax:f32:16384:8192 c3:layer_norm:f32:16384:8192 c3:softmax:f32:65536:4096"
python tools/perf_probe.py $P
AXB_WB_NO_ALTERNATE=1 python tools/perf_probe.py c3:layer_norm:f32 c3:layer_norm:bf16
