set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_matmul_gpu.py tests/test_attention_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 300 axiom_b200/host/_build/whisper_encoder 8 5 1 2>&1 | tail -1
timeout 300 axiom_b200/host/_build/whisper_encoder 32 3 0 2>&1 | tail -1
timeout 300 python tools/power_probe.py 4 ours > gpurun_out/r2_power_probe_c.json 2> gpurun_out/r2_power_probe_c.err; python -c "
import json; j=json.load(open('gpurun_out/r2_power_probe_c.json')); print({k:(round(v['TFLOPs_mean'],1), v['TFLOPs_over_time'], v['nvml_t_sm_mhz_power_w_temp_c'][-3:]) for k,v in j.items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_whisper_b8.csv axiom_b200/host/_build/whisper_encoder 8 3 0 > gpurun_out/r2_whisper_under_ncu.log 2>&1; wc -l gpurun_out/r2_launches_whisper_b8.csv
