python tools/bench_midsize.py > gpurun_out/mid_default.jsonl; TEXPRESSO_B200_LIB=tools/micro/ab_lane/lib_hc.so python tools/bench_midsize.py > gpurun_out/mid_hc.jsonl
paste -d'|' gpurun_out/mid_default.jsonl gpurun_out/mid_hc.jsonl | cut -c1-300
