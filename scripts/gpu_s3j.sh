mkdir -p gpurun_out
L=gpurun_out/s3j.log
timeout 300 python -m pytest tests/test_gpu_small.py -m gpu -x -q 2>&1 | tail -25 > $L
timeout 100 python scripts/quick_time.py 10,16,18 500 >> $L 2>&1
SMCB200_SMALL_MAX=0 timeout 100 python scripts/quick_time.py 10,16,18 500 >> $L 2>&1
cat $L
