python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -6
python tools/bench_extra.py --mips 1024 1 2>&1 | tail -1
python tools/bench_extra.py --mips 2048 1 2>&1 | tail -1
python tools/bench_batch.py 2>&1 | tail -5 | cut -c1-200
