mkdir -p gpurun_out
L=gpurun_out/island2.log
SMCB200_FORCE_ISLAND=1 timeout 200 python scripts/ab_time.py 24 100 > $L 2>&1
timeout 600 python -m pytest tests/test_gpu_island.py -m gpu -x -q 2>&1 | tail -5 >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/island_time.py 24 40 2>&1 | grep -E "^rank" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-400 >> $L
cat $L
