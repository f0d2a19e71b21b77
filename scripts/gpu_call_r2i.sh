set -x
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -x -k "ring and 6x6" > gpurun_out/r2i_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/r2i_sanitizer.log
timeout 900 python -m pytest tests/test_gpu_rowprog.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2i_pytest.log 2>&1; tail -5 gpurun_out/r2i_pytest.log
timeout 600 python scripts/probe_ring.py > gpurun_out/r2i_probe_ring.log 2>&1; cat gpurun_out/r2i_probe_ring.log
