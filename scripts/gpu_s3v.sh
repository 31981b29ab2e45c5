mkdir -p gpurun_out
L=gpurun_out/s3v.log
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_both_paths.py tests/test_gpu_parity_full.py tests/test_gpu_pflineart.py -m gpu -x -q 2>&1 | tail -8 > $L
timeout 300 python scripts/check_big_draws.py >> $L 2>&1
timeout 200 python scripts/modes_time.py >> $L 2>&1
timeout 100 python scripts/lineart_time.py >> $L 2>&1
cat $L
