# Round 2, second GPU call (r2b): first execution of the row-program kernel.
set -x
mkdir -p gpurun_out
# 1. parity of the new path under compute-sanitizer on a small case first (memcheck: out-of-bounds shared/global accesses)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -x -k "kat0 or (vs_oracle and 3-3-8-2 and c64)" > gpurun_out/r2b_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/r2b_sanitizer.log
timeout 600 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -x > gpurun_out/r2b_pytest_rows.log 2>&1; tail -15 gpurun_out/r2b_pytest_rows.log
# 2. headline workload: per-op vs rows under several knob settings
timeout 600 python scripts/probe_rows.py > gpurun_out/r2b_probe_rows.log 2>&1; tail -30 gpurun_out/r2b_probe_rows.log
ls -la gpurun_out | tail -5
