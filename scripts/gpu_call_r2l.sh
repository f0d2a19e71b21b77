set -x
mkdir -p gpurun_out
timeout 300 python scripts/probe_gemm_node.py > gpurun_out/r2l_gemm_node_k9.log 2>&1; cat gpurun_out/r2l_gemm_node_k9.log
PROBE_PK=6 PROBE_AMPS=64 timeout 300 python scripts/probe_gemm_node.py > gpurun_out/r2l_gemm_node_k6.log 2>&1; cat gpurun_out/r2l_gemm_node_k6.log
