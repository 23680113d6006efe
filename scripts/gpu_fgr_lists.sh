#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fgr.py -m gpu -q --tb=short 2>&1 | tail -4
for mode in 0 1; do
MGICP_FGR_LISTS=$mode timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fgr_lists_$mode.csv python - <<'PY' > /dev/null 2>&1
import sys
sys.path.insert(0, "."); import numpy as np, mgicp_b200 as m
G = "tests/golden/nclt"
cl = [m.pcd_io.read_pcd_xyz(f"{G}/s{i}.pcd") for i in (0, 1)]
eng = m.Engine(0)
_, feats = eng.fpfh_clouds(cl, 0.2, 20, 1.0, 200)
PY
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/fgr_lists_$mode.csv")) if len(r)>5]
hdr=None
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr:
        d=dict(zip(hdr,r))
        if d.get("Metric Name")=="gpu__time_duration.sum" and ("hybrid" in d["Kernel Name"] or "pfh" in d["Kernel Name"]): print("lists mode $mode  %-50s grid %-18s %10.1f us" % (d["Kernel Name"][:50], d["Grid Size"], float(d["Metric Value"].replace(",",""))/1e3))
PY
done
