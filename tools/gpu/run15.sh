timeout 1500 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -4
python tools/perf_probe.py c3:softmax:f32 c3:softmax:bf16 c3:argmax:f32 c3:argmax:bf16 c3:softmax:f32:16384:8192
