set -x
timeout 900 python -m pytest tests/test_reduce_gpu.py -q -x --timeout 600 -k "norm" 2>&1 | tail -6
cd axiom_b200/host/_build
timeout 300 ./api_parity 2>&1 | tail -1 | cut -c1-1200
timeout 300 ./whisper_encoder 2 1 1 2>&1 | tail -1 | cut -c1-200
for t in test_tensor_normalization test_nn_layers test_transformer_ops test_fusion test_fusion_extended test_cpu_gpu_parity; do ./reftest_$t 2>&1 | grep MINIGTEST_SUMMARY; done
