#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg3d.csv python tools/bench_configs.py cfg3d --steps 60 > gpurun_out/ncu_cfg3d.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_cfg3d.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki],float(r[vi].replace(',',''))) for r in rows[1:]]
tail=seq[-400:]
agg=collections.OrderedDict()
for k,v in tail:
    k=k[:70]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print('%5d  %9.1f us total  %7.2f us avg  %s'%(c,t/1e3,t/1e3/c,k))
P
