# Round 2, GPU call r2c: row-program kernel with whole-descriptor prefetch; per-level cycle profile.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rowprog.py -q -m gpu -x > gpurun_out/r2c_pytest_rows.log 2>&1; tail -5 gpurun_out/r2c_pytest_rows.log
PROBE_ONLY=${PROBE_ONLY:-per_op,rows,rows_tt8,rows_tt6,rows_cta1} timeout 600 python scripts/probe_rows.py > gpurun_out/r2c_probe_rows.log 2>&1; tail -30 gpurun_out/r2c_probe_rows.log
python - <<'PY'
import json
p = json.load(open("gpurun_out/op_profile_rows_rows.json"))
for v in p["variants"]:
    for o in v["ops"]:
        if o["name"] == "ROWPROG_CHUNK":
            print("level cycles (CTA 0, summed over its rows):", o["level_cycles_cta0"])
PY
