python -m pytest tests/test_gpu_alpha_lattice.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
./tools/micro/alpha_ab r02b 0 2>&1 | tee gpurun_out/alpha_ab_r02b.txt
python tools/bench_iter_e2e.py quick 2>&1 | tee gpurun_out/iter_e2e_r02b.txt
