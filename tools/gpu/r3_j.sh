timeout 300 python tools/ln_race_probe.py 2>&1 | tail -40
