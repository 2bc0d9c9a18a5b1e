python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/bench_e2e.py 2>&1 | tail -3 | tee gpurun_out/e2e_pixels.jsonl
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_final_n1.json | cut -c1-400
python bench.py --steps 3 --warmup 3 --workload iterative 2>&1 | tail -1 | tee gpurun_out/bench_iter_n1.json | cut -c1-300
python tools/bench_extra.py > gpurun_out/extra_final.jsonl 2>&1; cut -c1-200 gpurun_out/extra_final.jsonl | tail -30
