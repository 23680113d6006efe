#!/bin/bash
# ICP grid cell factor sweep on the script-2 schedule (bench default workload)
mkdir -p gpurun_out
for cf in $FACTORS; do
  timeout 300 python bench.py --pairs 296 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --icp-cell-factor $cf > gpurun_out/icf_$cf.json 2> gpurun_out/icf_$cf.err || tail -3 gpurun_out/icf_$cf.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/icf_$cf.json")); x = d["detail"]
    print("icp_cell_factor $cf: value %.1f ms/step %.2f icp %.2f" % (d["value"], d["ms_per_step"], x["ms_icp_per_step"]))
except Exception as e: print("failed $cf", e)
PY
done
