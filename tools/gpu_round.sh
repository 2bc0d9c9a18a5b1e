# one GPU call: lattice tests, full -m gpu suite, BC4/BC5 timings on three inputs, stand-alone harness on noise
python -m pytest tests/test_gpu_alpha_lattice.py -x -q 2>&1 | tail -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python tools/bench_extra.py --cases bc4,bc5 --reps 5 > gpurun_out/extra_alpha.jsonl 2>&1; cut -c1-130 gpurun_out/extra_alpha.jsonl
./tools/micro/ab2/narrow narrow
