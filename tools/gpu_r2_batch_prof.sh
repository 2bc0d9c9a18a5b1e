# launch list of a texture-batch run (BASELINE config 5 scaled to 64 textures): which kernels take the time of a texture group
mkdir -p gpurun_out
bash tools/gpu_sanitize.sh 2>&1 | tee gpurun_out/sanitizer_r02_lane.txt | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_batch_r02.csv python tools/bench_extra.py --mips 64 1 > gpurun_out/batch_under_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_batch_r02.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name'); ui=h.index('Metric Unit')
t=collections.Counter(); n=collections.Counter()
for r in rows[hdr+1:]:
    if len(r)<=vi or r[mi]!='gpu__time_duration.sum': continue
    v=float(r[vi].replace(',',''))
    v = v/1e3 if r[ui]=='ns' else (v*1e3 if r[ui]=='ms' else v)
    k=r[ki].split('(')[0][:70]; t[k]+=v; n[k]+=1
tot=sum(t.values())
for k,v in t.most_common(): print(f'{k:70s} n={n[k]:3d} {v/1e3:9.3f} ms {100*v/tot:5.1f}%  avg {v/n[k]:8.1f} us')
PY
python tools/bench_extra.py --mips 256 1 2>&1 | tail -1
