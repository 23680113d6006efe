#!/bin/bash
# ncu launch list (per-kernel durations) for a given pairs count: scripts/gpu_launches.sh <pairs> [extra bench args]
mkdir -p gpurun_out
P=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b$P.csv python bench.py --pairs $P --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_b$P.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_b$P.csv")) if len(r)>5]
hdr=None; data=[]
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr: data.append(dict(zip(hdr,r)))
per=[d for d in data if d.get("Metric Name")=="gpu__time_duration.sum"]
# last step = launches after the last k_bounds_init
idx=[i for i,d in enumerate(per) if d["Kernel Name"].startswith("k_bounds_init")]
last=per[idx[-2]:idx[-1]] if len(idx)>=2 else per[-24:]
tot=0
for d in last:
    v=float(d["Metric Value"].replace(",","")); tot+=v
    print(f"  {d['Kernel Name'][:50]:50s} grid {d['Grid Size']:>16s} {v/1e3:10.1f} us")
print("  total (us):", tot/1e3)
PY
