set -x
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q -k "test_layer_and_rms_norm_wide_and_cluster_f32" 2>&1 | grep -E "passed|failed|Error|assert|Mismatch|Max abs|where" | head -12; done
timeout 200 python tools/softmax_probe.py 8192 32768 bf16 | tail -1
timeout 200 python tools/softmax_probe.py 8192 32768 bf16 | tail -1
timeout 600 python -m pytest tests/test_matmul_gpu.py -m gpu -x -q 2>&1 | tail -3
cd axiom_b200/host/_build
timeout 300 ./whisper_encoder 8 10 1 2>&1 | tail -1 | cut -c1-600
