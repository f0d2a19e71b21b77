# Round 2, first GPU call (r2a): the artefacts round 1 left unrun + the baseline of this round.
#   gpurun --timeout 900 -- 'bash scripts/gpu_call_r2a.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.csv
# 1. tcgen05 bring-up kernel (self-check against fp64, then 4096 x 4096 x 512 timed)
timeout 60 scripts/microbench/bin/tc5_cgemm > gpurun_out/r2a_tc5.log 2>&1; echo "tc5 rc=$?" >> gpurun_out/r2a_tc5.log; tail -15 gpurun_out/r2a_tc5.log
# 2. headline step, autotuned (persists nothing yet: baseline for this round)
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench.jsonl 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.jsonl
cp gpurun_out/op_profile_rqc_7x7_d20_c64_s4096.json gpurun_out/r2a_op_profile.json 2>/dev/null
# 3. TMA-staged contraction kernel (QXB_SMEM_TMA=1): bounded waits trap instead of hanging
PROBE_CONFIGS=tma timeout 300 python scripts/probe_variants.py > gpurun_out/r2a_probe_tma.log 2>&1; tail -12 gpurun_out/r2a_probe_tma.log
ls -la gpurun_out | tail -20
