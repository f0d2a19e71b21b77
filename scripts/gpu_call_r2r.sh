set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2r_pytest_all.log 2>&1; tail -12 gpurun_out/r2r_pytest_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench.jsonl 2> gpurun_out/r2r_bench.err; tail -c 3000 gpurun_out/r2r_bench.jsonl; tail -3 gpurun_out/r2r_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2r_bench_ref.jsonl 2> gpurun_out/r2r_bench_ref.err; tail -c 600 gpurun_out/r2r_bench_ref.jsonl
