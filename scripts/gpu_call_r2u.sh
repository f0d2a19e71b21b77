set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python scripts/probe_multi.py > gpurun_out/r2u_probe_multi.log 2>&1; cat gpurun_out/r2u_probe_multi.log | tail -12
for N in 1 4 8; do
  if [ $N -eq 1 ]; then timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-as-given > gpurun_out/r2u_bench_n$N.jsonl 2> gpurun_out/r2u_bench_n$N.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r2u_bench_n$N.jsonl 2> gpurun_out/r2u_bench_n$N.err; fi
  tail -2 gpurun_out/r2u_bench_n$N.err
done
python - <<'PY'
import json
for N in (1, 4, 8):
    try:
        l = json.loads(open(f"gpurun_out/r2u_bench_n{N}.jsonl").read().strip().splitlines()[-1])
        print(N, f"{l['value']:.4g} amp/s", f"{l['ms_per_step']:.3f} ms", "e2e", f"{l['e2e']['value']:.4g}", l["roofline"]["bound"], round(l["roofline"]["frac"], 3), l["config"]["partition"][:60])
    except Exception as e:
        print(N, "failed", e)
PY
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2u_pytest_multi.log 2>&1; tail -3 gpurun_out/r2u_pytest_multi.log
