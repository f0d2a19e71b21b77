# Round 2, GPU call r2d: ncu --set full of the chunk row-program kernel (why is it 10x off the FP64 pipe bound?)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:rowprog_kernel -s 1 -c 1 -o /tmp/prof_rows python scripts/ncu_rows.py 2>&1 | tail -3
ncu -i /tmp/prof_rows.ncu-rep --page raw --csv > gpurun_out/r2d_rows.raw.csv 2>/dev/null
ncu -i /tmp/prof_rows.ncu-rep --page details --csv > gpurun_out/r2d_rows.details.csv 2>/dev/null
ls -la /tmp/prof_rows.ncu-rep gpurun_out/r2d*
