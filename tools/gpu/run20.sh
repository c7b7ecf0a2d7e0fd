timeout 900 python -m pytest tests/test_matmul_gpu.py -m gpu -x -q 2>&1 | tail -4
python tools/perf_probe.py mm:f32:12000:384:384:1 mm:f32:12000:1536:384:1:gelu mm:f32:12000:384:1536:1 mm:bf16:12000:384:384:1 mm:f32:4096:4096:4096 mm:bf16:4096:4096:4096 bmm
AXB_GEMM_BN=256 python tools/perf_probe.py mm:f32:12000:384:384:1 mm:bf16:12000:384:384:1
cd axiom_b200/host/_build; ./whisper_encoder 8 5 1 | tail -1; ./reftest_test_tensor_matmul | grep SUMMARY; ./reftest_test_flash_attention | grep SUMMARY
