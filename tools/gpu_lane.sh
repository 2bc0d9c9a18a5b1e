# one GPU call: lane-kernel parity tests, interleaved A/B of the ClusterFit kernel structures, full GPU suite, bench
python -m pytest tests/test_gpu_cluster_lane.py -x -q 2>&1 | grep -v "^$" | cut -c1-1200 | tail -8
python tools/ab_test.py --cases=bc3,bc1,bc3_smooth,bc1_smooth warp=texpresso_b200/libtexpresso_b200.so:warp lane=texpresso_b200/libtexpresso_b200.so:lane 2>&1 | tail -3 | tee gpurun_out/ab_lane.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_lane_n1.json
