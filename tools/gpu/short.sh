timeout 1500 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/perf_probe.py c3:layer_norm:f32:100000:384 c3:softmax:f32:100000:384 c3:layer_norm:f32:100000:512 c3:layer_norm:f32:100000:768 c3:layer_norm:f32:100000:1024
