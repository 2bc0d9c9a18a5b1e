# round 2, session 3 sanity call: smoke, GPU suite, default bench, and rank 0's shard of an 8-rank run on one GPU (pipeline tuning baseline)
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_o_n1.json 2> gpurun_out/bench_o_n1.err; tail -c 300 gpurun_out/bench_o_n1.err; cut -c1-300 gpurun_out/bench_o_n1.json
for n in 8 4 2; do python bench.py --shard-of $n --no-configs --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_o_shard$n.json; python -c "
import json; d=json.load(open('gpurun_out/bench_o_shard$n.json')); print('shard-of',$n,'dev ms/step',d['ms_per_step'],'e2e ms/step',d['e2e']['ms_per_step'],d['per_format'])"; done
