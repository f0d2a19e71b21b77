mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for G in 2 4 8; do
  if [ $G -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 20 --warmup 5 --no-cpu --no-as-given 2> gpurun_out/r2an_scale_$G.err | tail -1 > gpurun_out/r2an_scale_$G.json
    python -c "
import json
l=json.loads(open('gpurun_out/r2an_scale_$G.json').read()); print('N=$G', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'])"
  fi
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-as-given 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('N=1', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'])"
