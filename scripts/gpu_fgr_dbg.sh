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
                  iteration_number=3, tuple_scale=0.95, maximum_tuple_count=3747, seeds=[0])
PY
for lib in new ab/libmgicp_tc_nomma.so ab/libmgicp_tc_noepi.so ab/libmgicp_tc_neither.so; do
  if [ "$lib" = "new" ]; then unset MGICP_LIB; else export MGICP_LIB=$PWD/$lib; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_fgr_match_tc --csv --log-file gpurun_out/dbg.csv python /tmp/fgr_one.py > /dev/null 2>&1
  echo "$lib: $(grep gpu__time_duration gpurun_out/dbg.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
