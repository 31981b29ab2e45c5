mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
grep -E "FAILED|passed|failed|Error" gpurun_out/r2a_tests.log | head -20
timeout 120 python scripts/ab_time.py 24 100 3 2>&1 | tee gpurun_out/r2a_ab.log
timeout 120 python scripts/ab_time.py 24 100 2 2>&1 | tee -a gpurun_out/r2a_ab.log
for n in 16 20 22 26; do timeout 120 python scripts/ab_time.py $n 100 3; done 2>&1 | tee -a gpurun_out/r2a_ab.log
