mkdir -p gpurun_out
L=gpurun_out/s3g.log
timeout 100 python scripts/ab_time.py 24 100 > $L 2>&1
for v in p12 p8rev p12rev p6; do SMCB200_LIB=build/libsmcb200_$v.so timeout 100 python scripts/ab_time.py 24 100 >> $L 2>&1; done
timeout 100 python scripts/ab_time.py 24 100 >> $L 2>&1
cat $L
