set -x
mkdir -p gpurun_out
timeout 300 axiom_b200/host/_build/plan_cache 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_axiom_cpp.py -m gpu -q -x --timeout 900 2>&1 | tail -5
