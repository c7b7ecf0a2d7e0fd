python tools/perf_probe.py c3:softmax:f32:16384:8192 c3:softmax:f32:16384:8192 c3:softmax:f32 c3:softmax:f32:32768:4096 c3:softmax:bf16:16384:16384
