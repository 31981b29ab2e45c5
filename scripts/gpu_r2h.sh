# round-2 evidence: bench line, launch list of the same command, full ncu capture of the four kernels (plain and island code path)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err || exit 1
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2h_bench_short.json 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2h_ncu0.log 2>&1
timeout 120 python scripts/prof_target.py 24 6 > gpurun_out/r2h_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_rs_" -s 8 -c 4 -f -o gpurun_out/r2h_full python scripts/prof_target.py 24 6 > gpurun_out/r2h_ncu.log 2>&1
tail -n 1 gpurun_out/r2h_ncu.log
timeout 100 python scripts/sweep_time.py > gpurun_out/r2h_sweep.log 2>&1
tail -3 gpurun_out/r2h_bench.json | cut -c1-300
