timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s3w_launches_m0.csv python scripts/prof_target.py 24 8 0 > gpurun_out/s3w_ncu.log 2>&1
