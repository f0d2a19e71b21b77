set -x
mkdir -p gpurun_out
# 1. tcgen05 GEMM in the library: parity first (sanitizer on the random GEMM node), then timing against mma.sync / SIMT
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm_shaped_node_random and c32 and 2-0" > gpurun_out/r2k_sanitizer_tc5.log 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/r2k_sanitizer_tc5.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x -k "gemm or tensor_core" > gpurun_out/r2k_pytest_gemm.log 2>&1; tail -6 gpurun_out/r2k_pytest_gemm.log
PROBE_DTYPES=c32,c64 PROBE_MODES=1,4,2 timeout 600 python scripts/probe_gemm.py > gpurun_out/r2k_probe_gemm.log 2>&1; cat gpurun_out/r2k_probe_gemm.log
# 2. ring kernel policy (big rows only) + TMA kernel default
timeout 600 python scripts/probe_ring.py > gpurun_out/r2k_probe_ring.log 2>&1; cat gpurun_out/r2k_probe_ring.log
