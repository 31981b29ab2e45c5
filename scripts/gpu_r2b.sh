mkdir -p gpurun_out
timeout 120 python scripts/prof_target.py 24 6 > gpurun_out/r2b_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 2 -c 1 -f -o gpurun_out/r2d_step python scripts/prof_target.py 24 6 > gpurun_out/r2b_ncu.log 2>&1
tail -n 2 gpurun_out/r2b_ncu.log
