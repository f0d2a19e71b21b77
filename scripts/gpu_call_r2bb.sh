timeout 600 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -k "hoisted or rows_vs_oracle or kat0 or many_rows" 2>&1 | tail -3
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-as-given --amps 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('amps 10', l['value'], l['ms_per_step'], l['gpu_launches'], l['algorithmic'])"
