set -x
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -x -k "fused_chain and c64" > gpurun_out/r2o_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/r2o_sanitizer.log
timeout 900 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -x > gpurun_out/r2o_pytest.log 2>&1; tail -6 gpurun_out/r2o_pytest.log
timeout 600 python scripts/probe_ring.py > gpurun_out/r2o_probe_chain.log 2>&1; cat gpurun_out/r2o_probe_chain.log
PROBE_SLICES=2 PROBE_NO_C64=1 timeout 900 python scripts/probe_syc12.py > gpurun_out/r2o_probe_syc12_s16.log 2>&1; tail -30 gpurun_out/r2o_probe_syc12_s16.log
cp gpurun_out/probe_syc12.json gpurun_out/r2o_probe_syc12_s16.json 2>/dev/null
