mkdir -p gpurun_out
timeout 200 python scripts/ab_time.py 24 100 > gpurun_out/r2f_ab.log 2>&1
SMCB200_LIB=build/libsmcb200_massasc.so timeout 200 python scripts/ab_time.py 24 100 >> gpurun_out/r2f_ab.log 2>&1
timeout 200 python scripts/ab_time.py 26 50 >> gpurun_out/r2f_ab.log 2>&1
timeout 200 python scripts/ab_time.py 22 100 >> gpurun_out/r2f_ab.log 2>&1
cat gpurun_out/r2f_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rs_" -s 3 -c 3 -f -o gpurun_out/r2f_passes python scripts/prof_target.py 24 6 > gpurun_out/r2f_ncu.log 2>&1
tail -n 2 gpurun_out/r2f_ncu.log
