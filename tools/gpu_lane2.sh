python tools/size_sweep.py > gpurun_out/size_sweep.jsonl 2>&1; tail -40 gpurun_out/size_sweep.jsonl | cut -c1-150
ncu --set full --clock-control none --import-source on -k regex:cluster_lane -s 1 -c 1 -o gpurun_out/prof_lane_bc3 -f python tools/prof_lane.py > gpurun_out/ncu_lane.log 2>&1; tail -2 gpurun_out/ncu_lane.log
ncu --set full --clock-control none --import-source on -k regex:cluster_setup_sorted -s 1 -c 1 -o gpurun_out/prof_setup_bc3 -f python tools/prof_lane.py > gpurun_out/ncu_setup.log 2>&1; tail -2 gpurun_out/ncu_setup.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lane.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
