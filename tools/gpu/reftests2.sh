cd axiom_b200/host/_build
for t in reftest_test_conv_transpose reftest_test_custom_functors reftest_test_interpolate reftest_test_nn_activation_modules reftest_test_nn_ergonomics reftest_test_nn_groupnorm reftest_test_nn_positional reftest_test_tensor_numpy_ops reftest_test_tensor_random reftest_test_nn_module reftest_test_tensor_indexing reftest_test_shape_ops_extended; do
  timeout 300 ./$t > /root/repo/gpurun_out/$t.log 2>&1
  echo "$t rc=$? $(grep MINIGTEST_SUMMARY /root/repo/gpurun_out/$t.log | tail -1)"
  grep -E "^\[  FAILED  \] [A-Za-z]" /root/repo/gpurun_out/$t.log | sort -u | head -20
done
