cd axiom_b200/host/_build
timeout 300 ./spatial_parity
for t in test_conv_transpose test_interpolate test_conv1d; do timeout 300 ./reftest_$t 2>&1 | grep -E "FAILED|MINIGTEST_SUMMARY"; done
