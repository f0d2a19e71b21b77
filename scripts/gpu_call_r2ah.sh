set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowprog.py -q -m gpu 2>&1 | tail -3
for B in 1 0; do QXB_ROW_BANK_OPT=$B timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-as-given 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('bank_opt $B', l['value'], l['ms_per_step'], l['roofline']['ms_per_launch'], l['e2e']['value'])"; done
for B in 1 0; do QXB_ROW_BANK_OPT=$B timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-as-given --amps 1024 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('bank_opt $B amps 1024', l['value'], l['ms_per_step'])"; done
