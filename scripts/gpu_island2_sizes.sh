mkdir -p gpurun_out
L=gpurun_out/island2_sizes.log
timeout 600 python -m pytest tests/test_gpu_island.py -m gpu -x -q 2>&1 | tail -3 > $L
for ln in 24 21; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --log2n $ln 2>&1 | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('2 gpus log2n $ln ms_per_step',d['ms_per_step'],'value',d['value'])" >> $L 2>&1; done
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu --log2n 21 2>&1 | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('1 gpu log2n 21 ms_per_step',d['ms_per_step'])" >> $L 2>&1
cat $L
