mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
grep -E "FAILED|passed|failed|Error|assert|rc=" gpurun_out/r2c_tests.log | head -20
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 2500 gpurun_out/r2c_bench.json
timeout 200 python scripts/ab_time.py 24 100 > gpurun_out/r2c_ab.log 2>&1; tail -n 3 gpurun_out/r2c_ab.log
timeout 200 python scripts/ab_time.py 26 50 >> gpurun_out/r2c_ab.log 2>&1; tail -n 1 gpurun_out/r2c_ab.log
timeout 120 python scripts/prof_target.py 24 6 > gpurun_out/r2c_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches.csv python scripts/prof_target.py 24 12 > gpurun_out/r2c_ncu0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_rs_" -s 8 -c 4 -f -o gpurun_out/r2c_full python scripts/prof_target.py 24 6 > gpurun_out/r2c_ncu.log 2>&1
tail -n 2 gpurun_out/r2c_ncu.log
