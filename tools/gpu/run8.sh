timeout 900 python -m pytest tests/test_matmul_gpu.py -m gpu -x -q 2>&1 | tail -6
cd axiom_b200/host/_build
echo "== chains"; timeout 600 ./axiom_chains 4096 16384 10 2>&1 | tail -2
echo "== whisper B=1"; timeout 900 ./whisper_encoder 1 3 1 2>&1 | tail -2
echo "== whisper B=8"; timeout 900 ./whisper_encoder 8 5 0 2>&1 | tail -2
echo "== whisper B=32"; timeout 900 ./whisper_encoder 32 3 0 2>&1 | tail -2
./reftest_test_tensor_matmul | grep SUMMARY; ./reftest_test_cpu_gpu_parity | grep SUMMARY; ./reftest_test_flash_attention | grep SUMMARY
