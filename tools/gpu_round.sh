# one GPU call: lattice tests, full -m gpu suite, BC4/BC5 timings on three inputs, ncu of the BC5 kernel
python -m pytest tests/test_gpu_alpha_lattice.py -x -q 2>&1 | tail -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python tools/bench_extra.py --cases bc4,bc5 --reps 5 > gpurun_out/extra_alpha.jsonl 2>&1; cut -c1-130 gpurun_out/extra_alpha.jsonl
ncu --set full --clock-control none --import-source on -k regex:alpha_lattice -s 5 -c 1 -o gpurun_out/prof_lattice_bc5_r01d -f python tools/prof_alpha.py > gpurun_out/ncu_lat.log 2>&1; tail -1 gpurun_out/ncu_lat.log
