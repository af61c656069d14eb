#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_qr262k.csv python bench.py --workload qr262k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_qr.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_qr262k.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict(); cnt=collections.Counter()
for r in rows[1:]:
    k=r[ki].split('(')[0][:60]; agg[k]=agg.get(k,0.0)+float(r[vi].replace(',','')); cnt[k]+=1
tot=sum(agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]): print(f"{v/1e6:9.2f} ms {cnt[k]:5d} launches  {k}")
print('total', tot/1e6)
PY
