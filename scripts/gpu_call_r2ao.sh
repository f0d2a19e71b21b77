mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rowprog.py tests/test_gpu_parity.py -q -m gpu -k "row_options or fused_chain or big_times_small" 2>&1 | tail -3
NCU_CALLS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2ao_launches.csv python scripts/ncu_default_step.py > /dev/null 2>&1
wc -l gpurun_out/r2ao_launches.csv
