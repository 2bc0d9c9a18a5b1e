# 8-GPU evidence: headline bench at N=8 and BASELINE config 5 (scaled: 2048 textures 1024^2 + mips, BC3 ClusterFit) over 8 GPUs
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_lane_n8.json
python -c "
import json; d=json.load(open('gpurun_out/bench_lane_n8.json')); print(8, round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['clocks'])"
python tools/bench_extra.py --mips 2048 8 2>&1 | tail -1 | tee gpurun_out/mips_n8.json      # ~100 s: most of it generating 2048 textures on the host
