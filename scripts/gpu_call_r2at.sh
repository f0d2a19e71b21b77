set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2at_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2at_bench.jsonl 2> gpurun_out/r2at_bench.err; cut -c1-300 gpurun_out/r2at_bench.jsonl
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2at_bench_ref.jsonl 2> gpurun_out/r2at_bench_ref.err; cut -c1-200 gpurun_out/r2at_bench_ref.jsonl
python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
