mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload sycamore53_d12_c32_s16 --amps 1 --no-replan --steps 3 --warmup 3 --no-cpu --no-as-given 2> gpurun_out/r2av_bench_syc_2.err | tail -1 > gpurun_out/r2av_bench_syc_2.jsonl
tail -3 gpurun_out/r2av_bench_syc_2.err; cut -c1-900 gpurun_out/r2av_bench_syc_2.jsonl
