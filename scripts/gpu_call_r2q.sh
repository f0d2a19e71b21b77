set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "split_k or gemm" > gpurun_out/r2q_pytest_kred.log 2>&1; tail -6 gpurun_out/r2q_pytest_kred.log
PROBE_SLICES=2 timeout 900 python scripts/probe_syc12.py > gpurun_out/r2q_probe_syc12_s16.log 2>&1; tail -22 gpurun_out/r2q_probe_syc12_s16.log
cp gpurun_out/probe_syc12.json gpurun_out/r2q_probe_syc12_s16.json 2>/dev/null
