mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rowprog.py -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-as-given 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('half-warp model', l['value'], l['ms_per_step'], l['roofline']['ms_per_launch'], l['e2e']['value'])"
for A in 10 1024; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-as-given --amps $A 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('amps $A', l['value'], l['ms_per_step'])"; done
