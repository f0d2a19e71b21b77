set -x
mkdir -p gpurun_out
# 1. ncu --set full of the fused-chain launch of the default plan (4th rowprog launch: block, chain, block, CHAIN)
timeout 600 ncu --set full --clock-control none -k regex:rowprog_kernel -s 3 -c 1 -o /tmp/prof_chain python scripts/ncu_default_step.py > gpurun_out/r2al_ncu_chain.log 2>&1; tail -4 gpurun_out/r2al_ncu_chain.log
ncu -i /tmp/prof_chain.ncu-rep --page raw --csv > gpurun_out/r2al_chain.raw.csv 2>/dev/null
ncu -i /tmp/prof_chain.ncu-rep --page details --csv > gpurun_out/r2al_chain.details.csv 2>/dev/null
SHA=$(grep plan_sha1 gpurun_out/r2al_ncu_chain.log | awk '{print $2}')
python scripts/ncu_extract_traffic.py gpurun_out/r2al_chain.raw.csv gpurun_out/r2al_ncu_traffic.json rqc_7x7_d20_c64_s4096 "$SHA" 32768 "profiles/r2_ncu_chain.raw.csv (ncu --set full --clock-control none, the fused-chain launch of the default plan, 32768 bitstrings)"
# 2. launch list of one default step (serial launches; cold-cache, serialised: shares, not absolutes)
NCU_CALLS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2al_launches.csv python scripts/ncu_default_step.py > /dev/null 2>&1
wc -l gpurun_out/r2al_launches.csv
# 3. the whole GPU suite, then the bench lines
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2al_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2al_bench.jsonl 2> gpurun_out/r2al_bench.err; cut -c1-300 gpurun_out/r2al_bench.jsonl
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2al_bench_ref.jsonl 2> gpurun_out/r2al_bench_ref.err; cut -c1-200 gpurun_out/r2al_bench_ref.jsonl
for A in 1 10 1024; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-as-given --amps $A >> gpurun_out/r2al_bench_breadth.jsonl 2>> gpurun_out/r2al_bench_breadth.err; done
timeout 300 python bench.py --workload qft_20_unsliced --amps 1024 --steps 20 --warmup 3 --no-cpu --no-as-given >> gpurun_out/r2al_bench_breadth.jsonl 2>> gpurun_out/r2al_bench_breadth.err
for A in 10 1024 131072; do timeout 300 python bench.py --workload rqc_6x6_d16_c32_s64 --amps $A --steps 20 --warmup 3 --no-cpu --no-as-given >> gpurun_out/r2al_bench_breadth.jsonl 2>> gpurun_out/r2al_bench_breadth.err; done
python - <<'PY'
import json
for ln in open("gpurun_out/r2al_bench_breadth.jsonl"):
    l = json.loads(ln)
    print(l["config"]["workload"], l["config"]["n_amp_per_step"], f"{l['value']:.4g} amp/s", f"{l['ms_per_step']:.4f} ms", "e2e", f"{l['e2e']['value']:.4g}", l["roofline"]["bound"], l["roofline"]["frac"], "launches/step", l["algorithmic"]["launches_per_step"])
PY
