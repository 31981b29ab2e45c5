mkdir -p gpurun_out
timeout 200 python scripts/modes_time.py > gpurun_out/s3p_modes.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s3p_launches_m0.csv python scripts/prof_target.py 24 8 0 > gpurun_out/s3p_ncu.log 2>&1
cat gpurun_out/s3p_modes.log
