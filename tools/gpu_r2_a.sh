# round 2, call A: smoke, full GPU suite, headline bench with the new sub-records, hybrid-launch size sweep
python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02a_n1.json 2> gpurun_out/bench_r02a_n1.err; tail -c 600 gpurun_out/bench_r02a_n1.err; cut -c1-300 gpurun_out/bench_r02a_n1.json
python tools/size_sweep.py > gpurun_out/size_sweep_r02.jsonl 2>&1; tail -3 gpurun_out/size_sweep_r02.jsonl
