set -x
cd axiom_b200/host/_build
timeout 300 ./op_surface_parity 2>&1 | tail -1 | cut -c1-600
for t in test_tensor_basic test_nn_layers test_nn_module test_lazy_evaluation test_tensor_operations; do ./reftest_$t 2>&1 | grep MINIGTEST_SUMMARY; done
timeout 600 ./whisper_encoder 2 1 1 2>&1 | tail -1 | cut -c1-100
