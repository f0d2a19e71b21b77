set -x
mkdir -p gpurun_out
PROBE_SLICES=2 timeout 900 python scripts/probe_syc12.py > gpurun_out/r2n_probe_syc12_s16.log 2>&1; tail -30 gpurun_out/r2n_probe_syc12_s16.log
cp gpurun_out/probe_syc12.json gpurun_out/r2n_probe_syc12_s16.json 2>/dev/null
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
