mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --workload sycamore53_d12_c32_s16 --amps 1 --no-replan --steps 3 --warmup 3 --no-cpu --no-as-given 2> gpurun_out/r2aw_bench_syc_8.err | tail -1 > gpurun_out/r2aw_bench_syc_8.jsonl
tail -3 gpurun_out/r2aw_bench_syc_8.err; cut -c1-700 gpurun_out/r2aw_bench_syc_8.jsonl
