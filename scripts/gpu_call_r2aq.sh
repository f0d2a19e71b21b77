mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rowprog.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-as-given 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('bench', l['value'], l['ms_per_step'], l['roofline']['ms_per_launch'], l['e2e']['value'], l['config']['l2'])"
