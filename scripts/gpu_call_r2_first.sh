# First GPU call of round 2: everything the CPU-only r1q session could not measure (profiles/r1q_summary.md).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_call_r2_first.sh'
# Each step is bounded by its own timeout and writes into gpurun_out/; a failing step does not stop the others.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r2a_gpu.csv
# 1. parity: the new GPU test files first (never run on hardware), then the whole suite
timeout 600 python -m pytest tests/test_gpu_zfiles.py tests/test_gpu_zopen.py -q -m gpu -x > gpurun_out/r2a_pytest_new.log 2>&1; tail -5 gpurun_out/r2a_pytest_new.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest_all.log 2>&1; tail -5 gpurun_out/r2a_pytest_all.log
# 2. headline step: the bench's own measured choice (config.autotune lists all twelve candidates), then the two plans untuned
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench_autotune.jsonl 2> gpurun_out/r2a_bench_autotune.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-autotune > gpurun_out/r2a_bench_l1model.jsonl 2> gpurun_out/r2a_bench_l1model.err
QXB_PLAN_L1_BW=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-autotune > gpurun_out/r2a_bench_r1pmodel.jsonl 2> gpurun_out/r2a_bench_r1pmodel.err
cp gpurun_out/op_profile_rqc_7x7_d20_c64_s4096.json gpurun_out/r2a_op_profile_r1pmodel.json 2>/dev/null
# 3. register tiles for small nodes (QXB_MIN_LOB), per-op times and equality of the amplitudes
PROBE_CONFIGS=lob timeout 500 python scripts/probe_variants.py > gpurun_out/r2a_probe_lob.log 2>&1; tail -12 gpurun_out/r2a_probe_lob.log
# 3b. TMA-staged contraction kernel (QXB_SMEM_TMA=1): bounded waits trap instead of hanging; own timeout as well
PROBE_CONFIGS=tma timeout 400 python scripts/probe_variants.py > gpurun_out/r2a_probe_tma.log 2>&1; tail -12 gpurun_out/r2a_probe_tma.log
# 4. tcgen05 bring-up kernel: self-check against fp64, then the timed 4096 x 4096 x 512 case
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tc5_cgemm scripts/microbench/tc5_cgemm.cu \
  && timeout 60 /tmp/tc5_cgemm > gpurun_out/r2a_tc5.log 2>&1; cat gpurun_out/r2a_tc5.log
# 5. launch list of the new headline step (share of each kernel; absolute times are cold-cache)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2a_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
du -sm gpurun_out; ls -la gpurun_out | tail -20
