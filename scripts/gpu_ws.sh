#!/bin/bash
# warp-specialised task kernel (MGICP_WS=1): strict parity tests in task mode, then A/B timings
mkdir -p gpurun_out
MGICP_WS=1 timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "l1_strict or task_mode or config2" 2>&1 | tail -8
for ws in 0 1; do
  MGICP_WS=$ws timeout 150 python bench.py --pairs ${PAIRS:-296} --steps 3 --no-cpu-baseline --no-extras > gpurun_out/ws_$ws.json 2> gpurun_out/ws_$ws.err || tail -5 gpurun_out/ws_$ws.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ws_$ws.json")); x=d["detail"]
    print("MGICP_WS=$ws value=%.1f e2e=%.1f ms/step=%.2f icp_ms=%.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"]))
except Exception as e: print("failed ws=$ws", e)
PY
done
