# usage: bash tools/gpu_scale_n.sh N   -- headline bench at N ranks on an N-GPU box (JSON line -> gpurun_out/bench_scale_nN.json)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_scale_n$N.err | tail -1 > gpurun_out/bench_scale_n$N.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_scale_n$N.json'))
print($N, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'], 'e2e ms', d['e2e'].get('ms_per_step'), d['clocks'], 'multi', d.get('multi_matches_single'), 'parity', {k:v.get('pct') for k,v in d['parity'].items() if isinstance(v,dict)})
c=d.get('configs',{})
for k,v in c.items():
    print(' ', k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','mpix_s','e2e_ms','e2e_mpix_s','textures_per_s','matches_lone_call')})
PY
tail -c 300 gpurun_out/bench_scale_n$N.err
