mkdir -p gpurun_out
L=gpurun_out/final.log
python -c "import __graft_entry__ as g; g.smoke()" > $L 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $L
timeout 300 python bench.py > gpurun_out/final_bench1.json 2>> $L
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_benchref.json 2>> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/final_bench2.json 2>> $L
python - >> $L <<'PY'
import json
for f in ("final_bench1","final_benchref","final_bench2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("impl","n_gpus","value","ms_per_step","gpu_launches")}, (d.get("roofline") or {}).get("whole_path"), (d.get("e2e") or {}).get("value"), (d.get("check") or {}).get("island_vs_single_gpu"))
    except Exception as e:
        print(f, "ERR", e)
PY
cat $L
