# One GPU call that re-creates the evidence of a round: smoke, full -m gpu suite, headline bench (N=1), the other configurations,
# the ncu launch list of the bench command.  Run as:  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n1.json; cut -c1-160 gpurun_out/bench_n1.json
python bench.py --steps 3 --warmup 3 --workload iterative 2>&1 | tail -1 > gpurun_out/bench_iterative_n1.json; cut -c1-200 gpurun_out/bench_iterative_n1.json
python tools/bench_extra.py > gpurun_out/extra.jsonl 2>&1; cut -c1-150 gpurun_out/extra.jsonl | tail -30
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
