mkdir -p gpurun_out
SMCB200_FORCE_ISLAND=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_rs_" -s 8 -c 4 -f -o gpurun_out/s3f_island python scripts/prof_target.py 24 6 > gpurun_out/s3f_ncu_island.log 2>&1
tail -n 2 gpurun_out/s3f_ncu_island.log
