mkdir -p gpurun_out
timeout 900 python bench.py --workload sycamore53_d12_c32_s16 --amps 1 --no-replan --steps 3 --warmup 3 --no-cpu --no-as-given > gpurun_out/r2au_bench_syc_1.jsonl 2> gpurun_out/r2au_bench_syc_1.err
tail -3 gpurun_out/r2au_bench_syc_1.err; cut -c1-1500 gpurun_out/r2au_bench_syc_1.jsonl
