mkdir -p gpurun_out
cat > /tmp/lin.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rcppsmc_b200 as smc
import helpers as H
T, N = 60, 1 << 20
data = H.sim_lineart(T, 1)
mode = int(sys.argv[1])
smc.pfLineartBS(data, N, resample_mode=mode, threshold=0.5, seed=3)
PY
timeout 100 python scripts/lineart_time.py > gpurun_out/s3h_lin.log 2>&1
for m in 2 1; do timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3h_lin_launches_$m.csv python /tmp/lin.py $m > gpurun_out/s3h_ncu_$m.log 2>&1; done
cat gpurun_out/s3h_lin.log
