set -x
cd axiom_b200/host/_build
timeout 300 ./api_parity 2>&1 | tail -1 | cut -c1-1200
cd /root/repo
timeout 1500 python -m pytest tests/test_axiom_cpp.py tests/test_reduce_gpu.py -q -x --timeout 900 2>&1 | tail -3
