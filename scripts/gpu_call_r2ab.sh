set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bigsmall|kreduce_grid" -c 40 -o /tmp/syc python scripts/ncu_syc.py 2>&1 | tail -3
ncu -i /tmp/syc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_syc.raw.csv 2>/dev/null
ncu -i /tmp/syc.ncu-rep --page details --csv > gpurun_out/r2_ncu_syc.details.csv 2>/dev/null
ls -la /tmp/syc.ncu-rep gpurun_out/r2_ncu_syc.*
