mkdir -p gpurun_out
L=gpurun_out/s3n.log
timeout 300 python -m pytest tests/test_gpu_small.py tests/test_gpu_pfnonlinbs.py tests/test_gpu_csmc.py -m gpu -x -q 2>&1 | tail -4 > $L
timeout 100 python scripts/quick_time.py 9,10 500 >> $L 2>&1
timeout 100 python scripts/nc_small_time.py >> $L 2>&1
cat $L
