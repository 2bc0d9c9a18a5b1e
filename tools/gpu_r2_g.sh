for rep in 1 2; do for lib in gpurun_tmp/old.so texpresso_b200/libtexpresso_b200.so; do
  echo "== $lib"; TEXPRESSO_B200_LIB=$PWD/$lib python tools/bench_extra.py --cases cfg2,iter,smooth,range --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('  %-34s %-10s %9.4f ms' % (d['case'], d['size'], d['ms']))"
done; done 2>&1 | tee gpurun_out/ab_rolled_r02.txt
