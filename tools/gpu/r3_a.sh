set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for n in 8192 16384 24576 32768; do rows=$((268435456 / n)); timeout 200 python tools/rows_probe.py $rows $n 2>&1 | tail -1 | tee -a gpurun_out/r3_rows_probe_n.json; done
