# runs every reference-test runner on the GPU box; prints one summary line per file
cd axiom_b200/host/_build
for t in reftest_test_*; do
  [ -x "$t" ] || continue
  timeout 300 ./$t > /root/repo/gpurun_out/$t.log 2>&1
  rc=$?
  echo "$t rc=$rc $(grep MINIGTEST_SUMMARY /root/repo/gpurun_out/$t.log | tail -1)"
  grep -E "^\[  FAILED  \] [A-Za-z]" /root/repo/gpurun_out/$t.log | sort -u | head -12
done
