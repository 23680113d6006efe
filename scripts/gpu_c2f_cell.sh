#!/bin/bash
# ALL_FUNCTIONS schedule (search radius ~ cloud size): ICP time against the ICP grid's cell factor
cat > /tmp/c2f_cell.py <<'PY'
import sys, time
sys.path.insert(0, "."); import numpy as np, torch, mgicp_b200 as m
z = np.load("tests/golden/nclt_seq.npz"); zo = z["off"]
ncl = [z["xyz"][zo[i]:zo[i + 1]] for i in range(len(zo) - 1)]
fpairs = [tuple(p) for p in z["pairs"].tolist()]
eng = m.Engine(0); eng.set_timing(True)
from mgicp_b200.registration import create_scales
vox = create_scales(3); vox.reverse()
b = eng.cloud_bounds(ncl); dif = b[:, 3:] - b[:, :3]; rad = [(d[0] * d[1] * d[2]) ** (1 / 3) for d in dif]
dists = np.asarray([[(rad[s] + rad[t]) / 2 * (2 ** (-i)) for i in range(3)] for s, t in fpairs])
T0 = np.stack([z["T_fgr"][i] for i in range(len(fpairs))])
ref = None
for cf in [float(x) for x in sys.argv[1:]]:
    o = eng.make_opts(icp_cell_factor=cf)
    r = eng.run(ncl, fpairs, vox, dists, 100, T0, o)
    t0 = time.perf_counter(); r = eng.run(ncl, fpairs, vox, dists, 100, T0, o); t1 = time.perf_counter()
    tm = eng.get_timing()
    same = "" if ref is None else f" identical to first: {np.array_equal(ref, r.transformation)}"
    if ref is None: ref = r.transformation
    print(f"icp_cell_factor {cf}: run {1e3 * (t1 - t0):.1f} ms, icp {tm['icp_ms']:.1f} ms, icp_grid {tm['icp_grid_ms']:.2f}, iterations {r.iterations.mean(axis=0)}{same}")
PY
timeout 400 python /tmp/c2f_cell.py $FACTORS
