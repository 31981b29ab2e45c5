mkdir -p gpurun_out
L=gpurun_out/s3t.log
timeout 900 python -m pytest tests/test_gpu_both_paths.py tests/test_gpu_parity_full.py tests/test_gpu_pflineart.py tests/test_gpu_linreg.py -m gpu -x -q 2>&1 | tail -8 > $L
timeout 200 python scripts/modes_time.py >> $L 2>&1
timeout 100 python scripts/lineart_time.py >> $L 2>&1
cat $L
