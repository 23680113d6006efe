#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/fgr_one.py <<'PY'
import sys
sys.path.insert(0, "."); import numpy as np, mgicp_b200 as m
G = "tests/golden/nclt"
cl = [m.pcd_io.read_pcd_xyz(f"{G}/s{i}.pcd") for i in (0, 1)]
eng = m.Engine(0)
_, feats = eng.fpfh_clouds(cl, 0.2, 20, 1.0, 200)
for _ in range(2):
    eng.fgr_pairs(cl, feats, [(1, 0)], division_factor=1.4, use_absolute_scale=True, decrease_mu=True, maximum_correspondence_distance=0.2,
                  iteration_number=300, tuple_scale=0.95, maximum_tuple_count=3747, seeds=[0])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fgr_match_tc -s 1 -c 1 -o gpurun_out/prof_fgr_tc -f python /tmp/fgr_one.py > gpurun_out/prof_fgr_tc.log 2>&1
ls -la gpurun_out/prof_fgr_tc.ncu-rep
