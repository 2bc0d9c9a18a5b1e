python -m pytest tests/test_gpu_cluster_lane.py -x -q 2>&1 | grep -v "^$" | cut -c1-1500 | tail -8
python tools/size_sweep.py > gpurun_out/size_sweep2.jsonl 2>&1; tail -40 gpurun_out/size_sweep2.jsonl | cut -c1-170
