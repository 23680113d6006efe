#!/bin/bash
# tensor-core FGR matching: parity with the brute-force fp64 matcher (bit-identical poses), the oracle tests, then timings
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tail -30
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, torch
import mgicp_b200 as m, oracle
G = "tests/golden/nclt"
cl = [m.pcd_io.read_pcd_xyz(f"{G}/s{i}.pcd") for i in (0, 1, 17, 18)]
eng = m.Engine(0)
_, feats = eng.fpfh_clouds(cl, 0.2, 20, 1.0, 200)
pairs = [(1, 0), (0, 1), (3, 2), (2, 3), (1, 2)]
caps = [int(int((len(cl[s]) + len(cl[t])) / 2) * 0.2) for s, t in pairs]
kw = dict(division_factor=1.4, use_absolute_scale=True, decrease_mu=True, maximum_correspondence_distance=0.2, iteration_number=300,
          tuple_scale=0.95, maximum_tuple_count=caps, seeds=list(range(len(pairs))))
res = {}
for mode in ("0", "1"):
    os.environ["MGICP_FGR_MATCH"] = mode
    T, nc = eng.fgr_pairs(cl, feats, pairs, **kw)      # warm
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        T, nc = eng.fgr_pairs(cl, feats, pairs, **kw)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    res[mode] = (T, nc)
    print(f"match mode {mode}: {1e3 * dt:.2f} ms per call of {len(pairs)} pairs (whole mgicp_fgr_pairs incl. uploads), ncorr {nc.tolist()}")
same = np.array_equal(res["0"][0], res["1"][0]) and np.array_equal(res["0"][1], res["1"][1])
print("tensor-core matcher == brute-force fp64 matcher (poses bit-identical):", same, "max |dT|", np.abs(res["0"][0] - res["1"][0]).max())
assert same
PY
echo "script rc=$?"
timeout 600 python -m pytest tests/test_gpu_fgr.py -m gpu -q --tb=short 2>&1 | tail -5
for mode in 0 1; do
MGICP_FGR_MATCH=$mode timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fgr_launches_$mode.csv python - <<'PY' > /dev/null 2>&1
import sys
sys.path.insert(0, "."); import numpy as np, mgicp_b200 as m
G = "tests/golden/nclt"
cl = [m.pcd_io.read_pcd_xyz(f"{G}/s{i}.pcd") for i in (0, 1)]
eng = m.Engine(0)
_, feats = eng.fpfh_clouds(cl, 0.2, 20, 1.0, 200)
eng.fgr_pairs(cl, feats, [(1, 0)], division_factor=1.4, use_absolute_scale=True, decrease_mu=True, maximum_correspondence_distance=0.2,
              iteration_number=300, tuple_scale=0.95, maximum_tuple_count=3747, seeds=[0])
PY
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/fgr_launches_$mode.csv")) if len(r)>5]
hdr=None
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr:
        d=dict(zip(hdr,r))
        if d.get("Metric Name")=="gpu__time_duration.sum": print("mode $mode  %-60s grid %-18s %10.1f us" % (d["Kernel Name"][:60], d["Grid Size"], float(d["Metric Value"].replace(",",""))/1e3))
PY
done
