timeout 1500 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -4
P="c3:softmax:f32:16384:8192 c3:softmax:f32:32768:4096 c3:softmax:bf16:32768:8192 c3:softmax:bf16:16384:16384 c3:layer_norm:f32:32768:4096 c3:layer_norm:f32:16384:8192 c3:softmax:f32:72000:1500 c3:layer_norm:f32:100000:384 c3:softmax:f32 c3:softmax:bf16 c3:layer_norm:f32 c3:layer_norm:bf16"
python tools/perf_probe.py $P
cd axiom_b200/host/_build; ./whisper_encoder 8 5 1 | tail -1
