for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_targets.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize targets ok|=========.*(Invalid|Race|hazard|Barrier error)" | head -12
done
