cd axiom_b200/host/_build
echo "== chains"; timeout 600 ./axiom_chains 4096 16384 10 2>&1 | tail -5
echo "== whisper B=1"; timeout 900 ./whisper_encoder 1 3 1 2>&1 | tail -5
echo "== whisper B=8"; timeout 900 ./whisper_encoder 8 5 0 2>&1 | tail -5
echo "== reftests scatter/ops"; ./reftest_test_tensor_gather_scatter 2>&1 | grep -E "SUMMARY|FAILED  \] T"; ./reftest_test_tensor_where 2>&1 | grep -E "SUMMARY|FAILED  \] T"
