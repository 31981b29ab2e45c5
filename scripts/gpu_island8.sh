mkdir -p gpurun_out
L=gpurun_out/island8.log
timeout 900 python -m pytest tests/test_gpu_island.py -m gpu -x -q 2>&1 | tail -5 > $L
for g in 8 4; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $g --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/island8_bench$g.json; python -c "
import json;d=json.load(open('gpurun_out/island8_bench$g.json'));print($g,'gpus ms_per_step',d['ms_per_step'],'value',d['value'],'check',d['check']['island_vs_single_gpu'])" >> $L 2>&1; done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --log2n 21 2>&1 | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('strong-scaling point: 8 gpus x 2^21 ms_per_step',d['ms_per_step'],'value',d['value'])" >> $L 2>&1
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('1 gpu ms_per_step',d['ms_per_step'],'value',d['value'], d['roofline']['whole_path'])" >> $L 2>&1
cat $L
