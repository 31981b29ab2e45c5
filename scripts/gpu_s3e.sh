mkdir -p gpurun_out
L=gpurun_out/s3e.log
SMCB200_FORCE_ISLAND=1 timeout 120 python scripts/ab_time.py 24 100 > $L 2>&1
timeout 100 python scripts/ab_time.py 24 100 >> $L 2>&1
cat $L
