# final build of round 2: full ncu capture of the four kernels (plain) and of the island instantiation of k_step on one GPU
mkdir -p gpurun_out
timeout 120 python scripts/prof_target.py 24 6 > gpurun_out/r2j_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_rs_" -s 8 -c 4 -f -o gpurun_out/r2j_full python scripts/prof_target.py 24 6 > gpurun_out/r2j_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2j_full.ncu-rep gpurun_out/r2j_kernels > /dev/null 2>&1; rm -f gpurun_out/r2j_full.ncu-rep
SMCB200_FORCE_ISLAND=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step" -s 4 -c 1 -f -o gpurun_out/r2j_isl python scripts/prof_target.py 24 6 > gpurun_out/r2j_ncu_isl.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2j_isl.ncu-rep gpurun_out/r2j_island_kstep > /dev/null 2>&1; rm -f gpurun_out/r2j_isl.ncu-rep
grep -E "^==|time_duration|inst_executed.sum|registers" gpurun_out/r2j_kernels.txt gpurun_out/r2j_island_kstep.txt | head -30
