set -x
mkdir -p gpurun_out
PROBE_ONLY=default,chain_dmma,chain_tt8 timeout 600 python scripts/probe_ring.py > gpurun_out/r2w_probe_levels.log 2>&1; grep -E "^\[|chain:|cycles per" gpurun_out/r2w_probe_levels.log
