mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "edge_cases or kat or rqc" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-as-given 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('bench', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], 131072/l['e2e']['value']*1e3)"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-as-given --amps 16384 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('amps 16384', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'])"
