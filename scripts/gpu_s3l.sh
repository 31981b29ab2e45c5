mkdir -p gpurun_out
L=gpurun_out/s3l.log
timeout 300 python -m pytest tests/test_gpu_small.py tests/test_gpu_pmmh.py -m gpu -x -q 2>&1 | tail -5 > $L
timeout 100 python scripts/quick_time.py 10,14,16,18 500 >> $L 2>&1
timeout 200 python scripts/pmmh_time.py 100 >> $L 2>&1
cat $L
