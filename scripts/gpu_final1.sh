mkdir -p gpurun_out
L=gpurun_out/final1.log
python -c "import __graft_entry__ as g; g.smoke()" > $L 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
timeout 300 python bench.py > gpurun_out/final1_bench.json 2>> $L
python - >> $L <<'PY'
import json
d=json.loads(open("gpurun_out/final1_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline"]["whole_path"], d["e2e"]["value"], d["clocks"])
PY
cat $L
