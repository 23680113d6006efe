#!/bin/bash
# kNN grid cell factor sweep with the histogram-selection kNN (bench stage split)
mkdir -p gpurun_out
for cf in $FACTORS; do
  timeout 300 python bench.py --pairs 296 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --cell-factor $cf > gpurun_out/cf_$cf.json 2> gpurun_out/cf_$cf.err || tail -3 gpurun_out/cf_$cf.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/cf_$cf.json")); x = d["detail"]; s = x.get("stage_ms_per_step", {})
    print("cell_factor $cf: value %.1f ms/step %.2f icp %.2f | knn_grid %.2f sor %.2f normals %.2f" % (d["value"], d["ms_per_step"], x["ms_icp_per_step"], s.get("knn_grid_ms", -1), s.get("sor_ms", -1), s.get("normals_ms", -1)))
except Exception as e: print("failed $cf", e)
PY
done
