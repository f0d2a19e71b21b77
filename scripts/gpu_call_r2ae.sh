set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2ae_pytest.log
PROBE_SLICES=4 timeout 900 python scripts/probe_syc12.py > gpurun_out/r2ae_syc12.log 2>&1; tail -16 gpurun_out/r2ae_syc12.log; cp gpurun_out/probe_syc12.json gpurun_out/r2ae_probe_syc12.json; cp gpurun_out/op_profile_syc12.json gpurun_out/r2ae_op_profile_syc12.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ae_bench.jsonl 2> gpurun_out/r2ae_bench.err; cut -c1-600 gpurun_out/r2ae_bench.jsonl
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ae_bench_ref.jsonl 2> gpurun_out/r2ae_bench_ref.err; cut -c1-400 gpurun_out/r2ae_bench_ref.jsonl
