python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python tools/bench_extra.py --cases cfg2,range,smooth,iter --reps 5 2>&1 | cut -c1-200 | tee gpurun_out/extra_r02e.jsonl
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_r02e.json 2>gpurun_out/bench_r02e.err; cut -c1-400 gpurun_out/bench_r02e.json; tail -3 gpurun_out/bench_r02e.err
