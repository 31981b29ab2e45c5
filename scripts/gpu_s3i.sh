mkdir -p gpurun_out
L=gpurun_out/s3i.log
timeout 600 python -m pytest tests/test_gpu_units.py tests/test_gpu_pflineart.py tests/test_gpu_pfnonlinbs.py tests/test_gpu_parity_full.py -m gpu -x -q 2>&1 | tail -4 > $L
timeout 100 python scripts/lineart_time.py >> $L 2>&1
timeout 100 python scripts/ab_time.py 24 100 >> $L 2>&1
cat $L
