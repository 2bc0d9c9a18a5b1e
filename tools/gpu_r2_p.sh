mkdir -p gpurun_out
python tools/bench_wave_plan.py 2>&1 | tee gpurun_out/wave_plan_r02.jsonl
for n in 8 4; do python bench.py --shard-of $n --no-configs --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_p_shard$n.json; python -c "
import json; d=json.load(open('gpurun_out/bench_p_shard$n.json')); print('shard-of',$n,'dev ms/step',d['ms_per_step'],'e2e ms/step',d['e2e']['ms_per_step'],d['paths_agree'])"; done
