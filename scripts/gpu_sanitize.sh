#!/bin/bash
# compute-sanitizer over the kernels written in round 2 (small inputs): memcheck on the stage tests and the FGR registration test,
# racecheck (shared-memory hazards) on one small preprocessing + registration and one small FGR call
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, mgicp_b200 as m, oracle
eng = m.Engine(0)
src, tgt, T0, _ = m.synthetic.make_pair(120, seed=3)
r = m.multiscale_gicp(src, tgt, [1.0, 0.5, 0.25], [3.0, 1.0, 0.25], 8, T0, engine=eng)
print("gicp ok", r.fitness, r.iterations)
ds = np.asarray(oracle.voxel_down_sample(src, 0.6))
nrm, fp = eng.fpfh_clouds([ds, ds[:-11]], 1.2, 10, 4.0, 40)
T, nc = eng.fgr_pairs([ds, ds[:-11]], fp, [(0, 1), (1, 0)], use_absolute_scale=True, decrease_mu=True, maximum_correspondence_distance=1.0,
                      iteration_number=8, maximum_tuple_count=300, seeds=[1, 2])
print("fgr ok", nc.tolist(), np.abs(T[0] - np.eye(4)).max())
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/san_racecheck.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "sor or knn or normals or float64" > gpurun_out/san_stages.log 2>&1; echo "stages memcheck rc=$?"; tail -4 gpurun_out/san_stages.log
