set -x
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err | cut -c1-200; cut -c1-400 gpurun_out/r2_bench_final.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref_final.json 2>> gpurun_out/r2_bench_final.err; cut -c1-200 gpurun_out/r2_bench_ref_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_launches_bench_final.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|conv_prep|axb_ew_rowred|axb_ew_softmax|pool_kernel|sort_|sel_" -c 60 -o /tmp/r2_prof_b python tools/prof_targets_r2b.py > gpurun_out/r2_ncu_b.log 2>&1; tail -2 gpurun_out/r2_ncu_b.log
python tools/ncu_summary.py /tmp/r2_prof_b.ncu-rep > gpurun_out/r2_ncu_summary_b.txt 2>&1; wc -l gpurun_out/r2_ncu_summary_b.txt
cd axiom_b200/host/_build
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /root/repo/gpurun_out/r2_whisper_launches_b8_final.csv ./whisper_encoder 8 1 0 > /root/repo/gpurun_out/whisper_ncu.log 2>&1
cd /root/repo
for tool in memcheck racecheck synccheck; do echo "== $tool"; timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_targets_r2.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize targets ok|=========.*(Invalid|Race|hazard|Barrier error|Error)|Traceback|Error" | head -14; done 2>&1 | tee gpurun_out/r2_sanitizer_final.txt
