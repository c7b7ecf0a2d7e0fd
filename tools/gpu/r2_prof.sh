set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"axb_ew|lut_kernel|fa_kernel|gemm_tc_kernel|split_f32" -c 40 -o /tmp/r2_prof python tools/prof_targets_r2.py > gpurun_out/r2_ncu.log 2>&1; tail -3 gpurun_out/r2_ncu.log
python tools/ncu_summary.py /tmp/r2_prof.ncu-rep > gpurun_out/r2_ncu_summary.txt 2>&1; wc -l gpurun_out/r2_ncu_summary.txt
cp /tmp/r2_prof.ncu-rep gpurun_out/ 2>/dev/null; ls -la gpurun_out/r2_prof.ncu-rep
