mkdir -p gpurun_out
SMCB200_FORCE_ISLAND=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_step" -s 2 -c 4 --csv --log-file gpurun_out/s3q.csv python scripts/prof_target.py 24 8 > gpurun_out/s3q.log 2>&1
grep -v "^==" gpurun_out/s3q.csv | cut -d, -f5,13- | head -40
