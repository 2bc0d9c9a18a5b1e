set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests/test_gpu_alpha_lattice.py -x -q 2>&1 | tail -15
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for v in 0 1 6 7 8; do echo "variant $v"; TXP_ALPHA_VARIANT=$v python tools/bench_extra.py --cases bc4,bc5 --reps 5 2>&1 | tail -2; done
