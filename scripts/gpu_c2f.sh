#!/bin/bash
# where the batched Coarse_to_fine pipeline spends its time (64 NCLT pairs)
cat > /tmp/c2f.py <<'PY'
import sys, time
sys.path.insert(0, "."); import numpy as np, torch, mgicp_b200 as m
z = np.load("tests/golden/nclt_seq.npz"); zo = z["off"]
ncl = [z["xyz"][zo[i]:zo[i + 1]] for i in range(len(zo) - 1)]
fpairs = [tuple(p) for p in z["pairs"].tolist()]
eng = m.Engine(0)
eng.set_timing(True)
for rep in range(2):
    t0 = time.perf_counter()
    T, info, fit, rm = m.pose_graph.register_pairs(ncl, fpairs, 0.1, engine=eng)
    t1 = time.perf_counter()
    print(f"rep {rep}: {1e3 * (t1 - t0):.1f} ms", {k: (round(v, 2) if not isinstance(v, list) else [round(x, 2) for x in v[:3]]) for k, v in eng.get_timing().items()})
# the pieces
from mgicp_b200.registration import create_scales
vox = create_scales(3); vox.reverse()
b = eng.cloud_bounds(ncl); dif = b[:, 3:] - b[:, :3]; rad = [(d[0] * d[1] * d[2]) ** (1 / 3) for d in dif]
dists = np.asarray([[(rad[s] + rad[t]) / 2 * (2 ** (-i)) for i in range(3)] for s, t in fpairs])
print("voxels", vox, "dists[0]", dists[0])
T0 = np.stack([z["T_fgr"][i] for i in range(len(fpairs))])
for rep in range(2):
    t0 = time.perf_counter(); r = eng.run(ncl, fpairs, vox, dists, 100, T0); t1 = time.perf_counter()
    ev = eng.evaluate_clouds(ncl, fpairs, [0.1] * len(fpairs), r.transformation); t2 = time.perf_counter()
    print(f"run {1e3 * (t1 - t0):.1f} ms, evaluate_clouds {1e3 * (t2 - t1):.1f} ms, iterations mean {r.iterations.mean(axis=0)}", eng.get_timing()["icp_ms"])
PY
timeout 300 python /tmp/c2f.py
