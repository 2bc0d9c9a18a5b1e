for lib in texpresso_b200/libtexpresso_b200.so tools/micro/ab_lane/lib_s8.so; do echo "$lib (auto)"; for r in 1 2; do TEXPRESSO_B200_LIB=$lib python tools/bench_extra.py --mips 256 1 2>&1 | tail -1 | cut -c60-200; done; done
python -m pytest tests/test_gpu_parity.py -x -q -k "mip or batch or image_encode or thread" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120; python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
