# compute-sanitizer over the lane-per-block ClusterFit kernels and the pixel expansion (memcheck, racecheck, synccheck) on small GPU tests
echo "compute-sanitizer $(compute-sanitizer --version | tail -1), $(nvidia-smi --query-gpu=name --format=csv,noheader), round 1 session 3: cluster_setup_sorted / cluster_lane / cluster_lane_iter / expand_pixels kernels"
for tool in memcheck racecheck synccheck; do
  echo "$tool: tests/test_gpu_cluster_lane.py -k 'images or mipchain' + tests/test_gpu_pixels.py"
  TXP_ALPHA_TMA=$([ $tool = synccheck ] && echo 0 || echo 1) compute-sanitizer --tool $tool python -m pytest tests/test_gpu_cluster_lane.py tests/test_gpu_pixels.py -q -x -k "images or mipchain or pixels" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|error" | head -5
done
