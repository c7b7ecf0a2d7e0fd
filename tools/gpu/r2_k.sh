set -x
timeout 900 python -m pytest tests/test_pool_sort_gpu.py -q -x --timeout 600 2>&1 | tail -15
timeout 600 axiom_b200/host/_build/op_surface_parity 2>&1 | tail -2 | cut -c1-3000
