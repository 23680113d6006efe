#!/bin/bash
# FGR front end, batched over the 64 real NCLT pairs of tests/golden/nclt_seq.npz: tests, wall/device times, ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fgr.py tests/test_gpu_boundary.py -m gpu -q -x --tb=short 2>&1 | tail -5
cat > /tmp/fgr_batch.py <<'PY'
import sys, time
sys.path.insert(0, "."); import numpy as np, torch, mgicp_b200 as m
z = np.load("tests/golden/nclt_seq.npz"); zo = z["off"]
ncl = [z["xyz"][zo[i]:zo[i + 1]] for i in range(len(zo) - 1)]
fpairs = [tuple(p) for p in z["pairs"].tolist()]
caps = [int(int((len(ncl[s]) + len(ncl[t])) / 2) * 0.2) for s, t in fpairs]
kw = dict(division_factor=1.4, use_absolute_scale=True, decrease_mu=True, maximum_correspondence_distance=0.2, iteration_number=300,
          tuple_scale=0.95, maximum_tuple_count=caps, seeds=[m.pose_graph.pair_seed(0, s, t) for s, t in fpairs])
eng = m.Engine(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for r in range(reps):
    t0 = time.perf_counter()
    f = eng.fpfh_clouds(ncl, 0.2, 20, 1.0, 200, resident=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    T, nc = eng.fgr_pairs(None, f, fpairs, **kw); t2 = time.perf_counter()
    print(f"rep {r}: features {1e3 * (t1 - t0):.1f} ms, registration {1e3 * (t2 - t1):.1f} ms, {len(fpairs) / (t2 - t0):.1f} pairs/s")
PY
timeout 300 python /tmp/fgr_batch.py 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fgr_batch_launches.csv python /tmp/fgr_batch.py 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/fgr_batch_launches.csv")) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r[0] == "ID"]; rows = rows[h[0] + 1:]
agg = collections.OrderedDict()
for r in rows:
    try: name = r[4].split("(")[0]; v = float(r[-1].replace(",", ""))
    except Exception: continue
    u = r[-2]; v = v / 1e3 if u.startswith("us") else v / 1e6 if u.startswith("ns") else v * 1e3 if u in ("s", "second") else v
    agg.setdefault(name, [0, 0]); agg[name][0] += v; agg[name][1] += 1
print("total %.2f ms" % sum(a[0] for a in agg.values()))
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]: print("  %-28s %9.3f ms x%d" % (k, v, n))
PY
