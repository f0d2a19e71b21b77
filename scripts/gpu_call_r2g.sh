# ncu of the chunk row-program kernel: details + per-instruction (SASS) sampling
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:rowprog_kernel -s 1 -c 1 -o /tmp/prof_rows python scripts/ncu_rows.py 2>&1 | tail -3
ncu -i /tmp/prof_rows.ncu-rep --page details --csv > gpurun_out/r2g_rows.details.csv 2>/dev/null
ncu -i /tmp/prof_rows.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2g_rows.source.csv.gz
ls -la gpurun_out/r2g*
