# compute-sanitizer over the BC4/BC5 lattice kernels (memcheck, racecheck, synccheck) on the small GPU tests
echo "compute-sanitizer $(compute-sanitizer --version | tail -1), $(nvidia-smi --query-gpu=name --format=csv,noheader), round 1 session 2: alpha lattice kernels"
for tool in memcheck racecheck synccheck; do
  echo "$tool: tests/test_gpu_alpha_lattice.py -k 'images or exhaustive'"
  compute-sanitizer --tool $tool python -m pytest tests/test_gpu_alpha_lattice.py -q -x -k "images or exhaustive" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|error" | head -5
done
