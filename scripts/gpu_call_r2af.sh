set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
PROBE_MULTI_ONLY=syc PROBE_MULTI_REPS=3 timeout 1200 python scripts/probe_multi.py > gpurun_out/r2af_probe_multi.log 2>&1; tail -8 gpurun_out/r2af_probe_multi.log; cp gpurun_out/probe_multi.json gpurun_out/r2af_probe_multi.json
