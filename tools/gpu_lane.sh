timeout 900 bash tools/gpu_sanitize.sh > gpurun_out/sanitizer_lane.txt 2>&1; cat gpurun_out/sanitizer_lane.txt
python -m pytest tests/test_gpu_parity.py -x -q -k "multi_gpu or fullsize_8192 or thread_safe" 2>&1 | tail -2
