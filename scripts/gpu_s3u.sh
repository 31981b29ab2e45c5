mkdir -p gpurun_out
L=gpurun_out/s3u.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $L
timeout 800 python scripts/bench_rows.py gpurun_out/r2h_rows.json > gpurun_out/s3u_rows.log 2>&1; tail -3 gpurun_out/s3u_rows.log | cut -c1-300 >> $L
cat $L
