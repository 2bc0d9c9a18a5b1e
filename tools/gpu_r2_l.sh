ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_batch_r02.csv python tools/bench_extra.py --mips 64 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_batch_r02.csv')) if len(r)>10]
hdr=rows[0]; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    d=dict(zip(hdr,r)); k=d['Kernel Name'][:60]; agg[k][0]+=1; agg[k][1]+=float(d['Metric Value'])
for k,(n,t) in agg.items(): print(f"{k:62s} n={n:4d} total={t/1e3:9.1f} us  avg={t/n/1e3:8.1f} us")
PY
