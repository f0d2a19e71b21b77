set -x
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -x -k "fused_chain and c64" > gpurun_out/r2v_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/r2v_sanitizer.log
timeout 1200 python -m pytest tests/test_gpu_rowprog.py tests/test_gpu_fullsize.py -q -m gpu -x > gpurun_out/r2v_pytest.log 2>&1; tail -6 gpurun_out/r2v_pytest.log
timeout 600 python scripts/probe_ring.py > gpurun_out/r2v_probe_dmma.log 2>&1; grep -E "^\[|chain:|ROWPROG" gpurun_out/r2v_probe_dmma.log
