mkdir -p gpurun_out
for v in nolw nolwwt; do SMCB200_LIB=build/libsmcb200_$v.so timeout 120 python scripts/ab_time.py 24 100 3; done 2>&1 | tee gpurun_out/r2b_ab.log
