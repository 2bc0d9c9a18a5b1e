# round 2, call B: TMA-staged BC4/BC5 kernel (tests + A/B against the cp.async kernel), chunk-plan experiments
python -m pytest tests/test_gpu_alpha_lattice.py tests/test_gpu_parity.py tests/test_gpu_cluster_lane.py -x -q 2>&1 | tail -8
./tools/micro/alpha_ab r02 0 2>&1 | tee gpurun_out/alpha_ab_r02.txt
./tools/micro/alpha_ab r02smooth 1 2>&1 | tee -a gpurun_out/alpha_ab_r02.txt
for t in 1 0 1 0; do TXP_ALPHA_TMA=$t python tools/bench_extra.py --cases bc4,bc5 --reps 7 2>&1 | grep r_rg\" | cut -c1-110 | sed "s/^/tma=$t /"; done | tee gpurun_out/extra_tma_r02.txt
python tools/bench_iter_e2e.py 2>&1 | tee gpurun_out/iter_e2e_r02.txt
