mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
grep -E "FAILED|passed|failed|Error|assert|rc=" gpurun_out/r2e_tests.log | head -20
timeout 200 python scripts/ab_time.py 24 100 > gpurun_out/r2e_ab.log 2>&1
SMCB200_NO_PDL=1 timeout 200 python scripts/ab_time.py 24 100 >> gpurun_out/r2e_ab.log 2>&1
for r in 0.3 0.5 0.7 1.0; do SMCB200_L2_VERBOSE=1 SMCB200_L2_PERSIST=$r timeout 200 python scripts/ab_time.py 24 100 >> gpurun_out/r2e_ab.log 2>&1; done
timeout 200 python scripts/ab_time.py 26 50 >> gpurun_out/r2e_ab.log 2>&1
timeout 200 python scripts/ab_time.py 22 100 >> gpurun_out/r2e_ab.log 2>&1
timeout 200 python scripts/ab_time.py 20 100 >> gpurun_out/r2e_ab.log 2>&1
cat gpurun_out/r2e_ab.log
./build/mix > gpurun_out/r2e_mix.log 2>&1; cat gpurun_out/r2e_mix.log
./build/dfma > gpurun_out/r2e_dfma.log 2>&1; cat gpurun_out/r2e_dfma.log
