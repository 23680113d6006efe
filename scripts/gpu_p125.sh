#!/bin/bash
mkdir -p gpurun_out
for P in 125 ; do
for G in 0 -2 -3 -4 -6; do
  timeout 600 python bench.py --pairs $P --steps 5 --no-cpu-baseline --config3-pairs 0 --config5-pairs 0 --ctas-per-pair=$G > gpurun_out/p${P}_$G.json 2> gpurun_out/p${P}_$G.err
  python - <<PY
import json
d=json.load(open("gpurun_out/p${P}_$G.json")); x=d["detail"]; s=x["stage_ms_per_step"]
print("pairs $P ctas_per_pair $G: value %.0f e2e %.0f ms/step %.2f | ds %.2f grid %.2f sor %.2f nrm %.2f igrid %.2f icp %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], s["downsample_ms"], s["knn_grid_ms"], s["sor_ms"], s["normals_ms"], s["icp_grid_ms"], s["icp_ms"]))
PY
done; done
