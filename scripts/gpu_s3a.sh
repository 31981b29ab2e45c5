mkdir -p gpurun_out
L=gpurun_out/s3a_ab.log
timeout 200 python scripts/ab_time.py 24 100 > $L 2>&1
SMCB200_FORCE_ISLAND=1 timeout 200 python scripts/ab_time.py 24 100 >> $L 2>&1
for v in mass4 mass12 cnt4map6; do SMCB200_LIB=build/libsmcb200_$v.so timeout 200 python scripts/ab_time.py 24 100 >> $L 2>&1; done
cat $L
