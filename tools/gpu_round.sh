# one GPU call: full -m gpu test suite, BC4/BC5 timings, ncu captures of the lattice kernels
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/bench_extra.py --cases bc4,bc5 --reps 5 > gpurun_out/extra_alpha.jsonl 2>&1; cut -c1-200 gpurun_out/extra_alpha.jsonl
ncu --set full --clock-control none --import-source on -k regex:alpha_lattice -s 2 -c 1 -o gpurun_out/prof_lattice_bc4_r01c -f python tools/prof_alpha.py > gpurun_out/ncu_lat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:alpha_lattice -s 5 -c 1 -o gpurun_out/prof_lattice_bc5_r01c -f python tools/prof_alpha.py >> gpurun_out/ncu_lat.log 2>&1
tail -2 gpurun_out/ncu_lat.log
