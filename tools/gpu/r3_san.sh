mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do echo "== $tool"; timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_targets_r2c.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize targets ok.*|=========.*(Invalid|Race|hazard|Barrier error|Error)|Traceback|Error" | head -14; done 2>&1 | tee gpurun_out/r3_sanitizer.txt
