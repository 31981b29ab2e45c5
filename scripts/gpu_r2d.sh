mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pfnonlinbs.py tests/test_gpu_parity_full.py tests/test_gpu_pflineart.py tests/test_gpu_pmmh.py -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
grep -E "FAILED|passed|failed|Error|assert|rc=" gpurun_out/r2d_tests.log | head -20
timeout 200 python scripts/ab_time.py 24 100 > gpurun_out/r2d_ab.log 2>&1
SMCB200_LIB=build/libsmcb200_minb4.so timeout 200 python scripts/ab_time.py 24 100 >> gpurun_out/r2d_ab.log 2>&1
SMCB200_LIB=build/libsmcb200_q2.so timeout 200 python scripts/ab_time.py 24 100 >> gpurun_out/r2d_ab.log 2>&1
timeout 200 python scripts/ab_time.py 26 50 >> gpurun_out/r2d_ab.log 2>&1
cat gpurun_out/r2d_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_step" -s 2 -c 1 -f -o gpurun_out/r2d_step python scripts/prof_target.py 24 6 > gpurun_out/r2d_ncu.log 2>&1
tail -n 2 gpurun_out/r2d_ncu.log
