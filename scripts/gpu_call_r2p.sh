set -x
mkdir -p gpurun_out
timeout 600 python scripts/probe_ring.py > gpurun_out/r2p_probe_chain.log 2>&1; cat gpurun_out/r2p_probe_chain.log
