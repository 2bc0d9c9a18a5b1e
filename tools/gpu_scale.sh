for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_lane_n$n.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_lane_n$n.json')); print($n, round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['clocks'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --workload iterative 2>/dev/null | tail -1 > gpurun_out/bench_lane_iter_n8.json; python -c "
import json; d=json.load(open('gpurun_out/bench_lane_iter_n8.json')); print('iter 8', round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"
