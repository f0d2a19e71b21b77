mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -k "fused_chain and c64" -x 2>&1 | grep -E "Error|error|assert|kw|^E " | head -30
