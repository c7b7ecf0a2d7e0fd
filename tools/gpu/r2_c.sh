set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -25
timeout 300 python -m pytest tests/test_matmul_gpu.py -m gpu -q -x -k "constant_operand or split" --timeout 120 2>&1 | tail -8
timeout 300 axiom_b200/host/_build/whisper_encoder 8 5 1 > gpurun_out/r2_whisper_b8.json 2> gpurun_out/r2_whisper_b8.err; cat gpurun_out/r2_whisper_b8.json; tail -3 gpurun_out/r2_whisper_b8.err
timeout 300 axiom_b200/host/_build/reftest_test_flash_attention 2>&1 | tail -5
timeout 300 axiom_b200/host/_build/reftest_test_nn_attention 2>&1 | tail -3
