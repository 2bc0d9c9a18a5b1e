python -m pytest tests -x -q -m gpu 2>&1 | tail -4
bash tools/gpu_r2_g.sh > /dev/null 2>&1; cat gpurun_out/ab_rolled_r02.txt
