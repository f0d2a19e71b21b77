set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "big_times_small or split_k" 2>&1 | tail -3
PROBE_NO_C64=1 PROBE_SLICES=4 timeout 900 python scripts/probe_syc12.py > gpurun_out/r2ag_syc12.log 2>&1; tail -16 gpurun_out/r2ag_syc12.log; cp gpurun_out/probe_syc12.json gpurun_out/r2ag_probe_syc12.json; cp gpurun_out/op_profile_syc12.json gpurun_out/r2ag_op_profile_syc12.json
