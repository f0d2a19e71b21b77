mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r2ax_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ax_bench.jsonl 2> gpurun_out/r2ax_bench.err; cut -c1-260 gpurun_out/r2ax_bench.jsonl
for A in 1 10 1024 16384; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-as-given --amps $A 2>/dev/null | tee -a gpurun_out/r2ax_breadth.jsonl | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('amps $A', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'])"; done
timeout 300 python bench.py --workload rqc_6x6_d16_c32_s64 --amps 10 --steps 20 --warmup 3 --no-cpu --no-as-given 2>/dev/null | tee -a gpurun_out/r2ax_breadth.jsonl | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('6x6 amps 10', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'])"
