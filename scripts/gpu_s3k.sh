mkdir -p gpurun_out
L=gpurun_out/s3k.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $L
timeout 200 python scripts/pmmh_time.py 100 >> $L 2>&1
SMCB200_SMALL_MAX=0 timeout 200 python scripts/pmmh_time.py 100 >> $L 2>&1
cat $L
