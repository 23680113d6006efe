#!/bin/bash
# FGR tests + one default bench run, summarised
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fgr.py tests/test_gpu_boundary.py -m gpu -q -x -s --tb=short 2>&1 | grep -E "popular|passed|failed|Error" | tail -6
timeout 900 python bench.py --gpus 1 --steps ${STEPS:-10} --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench.json")); x = d["detail"]
print("value %.1f e2e %.1f ms/step %.2f icp %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"]))
print("fgr", json.dumps({k: v for k, v in x["fgr_front_end"].items() if k != "workload"}))
print("c2f", json.dumps({k: v for k, v in x["coarse_to_fine"].items() if k not in ("workload", "note")}))
PY
