mkdir -p gpurun_out
for T in 8 6; do QXB_CHAIN_MIN_TT=$T timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-as-given 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('chain_min_tt $T', l['value'], l['ms_per_step'], l['roofline']['ms_per_launch'], l['e2e']['value'])"; done
QXB_ROW_DMMA=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-as-given 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('dmma', l['value'], l['ms_per_step'], l['roofline']['ms_per_launch'], l['e2e']['value'])"
