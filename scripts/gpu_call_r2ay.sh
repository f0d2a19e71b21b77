mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowprog.py tests/test_gpu_zfiles.py tests/test_gpu_zopen.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2ay_pytest.log
