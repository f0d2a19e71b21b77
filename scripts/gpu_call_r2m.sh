set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2m_pytest_multi.log 2>&1; tail -8 gpurun_out/r2m_pytest_multi.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2m_bench_n1.jsonl 2> gpurun_out/r2m_bench_n1.err; tail -c 2500 gpurun_out/r2m_bench_n1.jsonl; tail -3 gpurun_out/r2m_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2m_bench_n2.jsonl 2> gpurun_out/r2m_bench_n2.err; tail -c 1200 gpurun_out/r2m_bench_n2.jsonl; tail -3 gpurun_out/r2m_bench_n2.err
