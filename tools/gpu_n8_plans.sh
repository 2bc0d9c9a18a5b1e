# 8-GPU run of the headline bench (no configs) under different explicit chunk plans of the per-rank shard (TXP_PLAN, experiments only)
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-configs 2>/dev/null | tail -1 > gpurun_out/n8_$name.json
  python -c "
import json; d=json.load(open('gpurun_out/n8_$name.json')); print('$name', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['paths_agree'])"
}
run auto A=1
run p55x4 TXP_PLAN="8,28,55x4;L=2"
run auto2 A=1
run p4_55x4 TXP_PLAN="4,32,55x4;L=2"
run p27x2_55x4 TXP_PLAN="8,14,14,55x4;L=3"
