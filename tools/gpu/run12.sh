timeout 1200 python -m pytest tests/test_matmul_gpu.py tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -4
cd axiom_b200/host/_build; ./whisper_encoder 8 5 1 | tail -1; ./whisper_encoder 32 3 0 | tail -1
./reftest_test_flash_attention | grep SUMMARY; ./reftest_test_nn_attention | grep SUMMARY; ./reftest_test_cpu_gpu_parity | grep SUMMARY;  ./reftest_test_tensor_normalization | grep SUMMARY
