set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowprog.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2h_pytest.log 2>&1; tail -5 gpurun_out/r2h_pytest.log
timeout 600 python scripts/probe_batch.py > gpurun_out/r2h_probe_batch.log 2>&1; cat gpurun_out/r2h_probe_batch.log
PROBE_WORKLOAD=rqc_6x6_d16_c32_s64 timeout 600 python scripts/probe_batch.py > gpurun_out/r2h_probe_batch_6x6.log 2>&1; cat gpurun_out/r2h_probe_batch_6x6.log
