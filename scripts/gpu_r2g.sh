mkdir -p gpurun_out
timeout 200 python scripts/ab_time.py 24 100 > gpurun_out/r2g_ab.log 2>&1
SMCB200_LIB=build/libsmcb200_map4.so timeout 200 python scripts/ab_time.py 24 100 >> gpurun_out/r2g_ab.log 2>&1
cat gpurun_out/r2g_ab.log
timeout 300 python -m pytest tests/test_gpu_units.py tests/test_gpu_pfnonlinbs.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_user_model.py -m gpu -q -x 2>&1 | tail -15
./tests/cpp/user_model 16384 2>&1 | tail -12
