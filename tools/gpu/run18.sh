P="c3:softmax:f32:16384:8192 c3:softmax:f32:32768:4096 c3:softmax:bf16:32768:8192 c3:layer_norm:f32:32768:4096 c3:layer_norm:f32:16384:8192 c3:softmax:f32:65536:2048 c3:softmax:f32:72000:1500"
AXB_NO_WIDE_ROWS=1 AXB_NO_CLUSTER_ROWS=1 python tools/perf_probe.py $P
