set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2s_pytest_all.log 2>&1; tail -12 gpurun_out/r2s_pytest_all.log
