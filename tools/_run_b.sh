timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "non_finite" 2>&1 | grep -v "^$" | tail -30
