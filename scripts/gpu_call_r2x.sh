set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "big_times_small" 2>&1 | tail -5
PROBE_SLICES=4 timeout 900 python scripts/probe_syc12.py > gpurun_out/r2x_syc12.log 2>&1; tail -20 gpurun_out/r2x_syc12.log; cp gpurun_out/probe_syc12.json gpurun_out/r2x_probe_syc12.json
QXB_BIGSMALL=0 PROBE_NO_C64=1 PROBE_SLICES=4 timeout 900 python scripts/probe_syc12.py > gpurun_out/r2x_syc12_off.log 2>&1; tail -16 gpurun_out/r2x_syc12_off.log
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2x_bench.jsonl 2> gpurun_out/r2x_bench.err; cat gpurun_out/r2x_bench.jsonl
