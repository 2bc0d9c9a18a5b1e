python -m pytest tests/test_gpu_cluster_lane.py -x -q 2>&1 | grep -v "^$" | cut -c1-1500 | tail -12
python tools/ab_test.py --cases=bc3,bc1,bc1_iter,bc3_smooth,bc3_smooth_iter warp=texpresso_b200/libtexpresso_b200.so:warp lane=texpresso_b200/libtexpresso_b200.so:lane 2>&1 | tail -3 | tee gpurun_out/ab_lane3.txt
