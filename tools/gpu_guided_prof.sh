#!/bin/bash
set -u
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/guided_launches.csv python bench.py --workload guided --frames 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/guided_ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/guided_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value")
t=collections.defaultdict(lambda:[0,0.0,0.0])
for r in rows[1:]:
    k=r[ki][:50]
    v=float(r[vi].replace(",",""))
    if r[mi].startswith("gpu__time"): t[k][0]+=1; t[k][1]+=v
    else: t[k][2]+=v
for k,(n,us,inst) in sorted(t.items(), key=lambda kv:-kv[1][1])[:8]:
    print(f"{k:52s} n={n:3d} total={us/1e3:9.1f}us inst={inst/1e6:8.1f}M")
PY
