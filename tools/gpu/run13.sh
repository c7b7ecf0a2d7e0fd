cd axiom_b200/host/_build; ./axiom_chains 4096 16384 20 | tail -1; ./whisper_encoder 8 5 1 | tail -1
for t in reftest_test_cpu_gpu_parity reftest_test_fusion_extended reftest_test_tensor_operations reftest_test_lazy_evaluation reftest_test_nn_layers; do ./$t | grep SUMMARY; done
