mkdir -p gpurun_out
L=gpurun_out/s3d.log
SMCB200_FORCE_ISLAND=1 timeout 200 python scripts/ab_time.py 24 100 > $L 2>&1
for v in nofence nosleep; do SMCB200_FORCE_ISLAND=1 SMCB200_LIB=build/libsmcb200_$v.so timeout 200 python scripts/ab_time.py 24 100 >> $L 2>&1; done
cat $L
