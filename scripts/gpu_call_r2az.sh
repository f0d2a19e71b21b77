timeout 600 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -k "hoisted" 2>&1 | tail -4
