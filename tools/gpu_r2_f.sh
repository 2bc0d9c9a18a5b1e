ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none --csv --log-file gpurun_out/launches_prof_range.csv python tools/prof_range.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_prof_range.csv')) if len(r)>10]
hdr=rows[0]; out={}
for r in rows[1:]:
    d=dict(zip(hdr,r)); k=(d['ID'],d['Kernel Name'][:70]); out.setdefault(k,{})[d['Metric Name']]=d['Metric Value']
for (i,n),m in out.items(): print(i,n,m)
PY
