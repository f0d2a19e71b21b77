# Targeted ncu captures of the dominant contractions of the headline workload (see scripts/ncu_ops.py).
# Only CSV extracts are kept: the .ncu-rep files (40-50 MB each with the whole cubin) exceed what gpurun copies back.
set -x
mkdir -p gpurun_out
NCU_CASE=rqc QXB_NCU_OPS=R474,R260,R473,R571,R779,R472 timeout 100 ncu --profile-from-start off --set full --clock-control none \
    -o /tmp/prof_rqc python scripts/ncu_ops.py 2>&1 | tail -3
ncu -i /tmp/prof_rqc.ncu-rep --page raw --csv > gpurun_out/prof_r1p_rqc_dominant.raw.csv 2>/dev/null
ncu -i /tmp/prof_rqc.ncu-rep --page details --csv > gpurun_out/prof_r1p_rqc_dominant.details.csv 2>/dev/null
NCU_CASE=rqc timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1p.csv \
    python scripts/ncu_ops.py 2>&1 | tail -2
du -sm gpurun_out; ls -la gpurun_out
