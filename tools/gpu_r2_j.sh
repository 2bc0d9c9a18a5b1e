python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
TEXPRESSO_B200_LIB=$PWD/texpresso_b200/libtexpresso_b200.so python tools/bench_extra.py --cases range --reps 7 2>&1 | cut -c1-140
