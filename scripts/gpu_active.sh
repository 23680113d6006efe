#!/bin/bash
# sweep of (active pairs, chunks per pass) for the ICP task mode: "A:V" pairs
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || { timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log; }
for cfg in $SWEEP; do
  IFS=: read A V <<< "$cfg"
  export MGICP_ACTIVE_PAIRS=$A
  timeout 600 python bench.py --pairs ${PAIRS:-148} --steps 3 --no-cpu-baseline --ctas-per-pair=-$V > gpurun_out/act_${A}_$V.json 2> gpurun_out/act_${A}_$V.err || tail -5 gpurun_out/act_${A}_$V.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/act_${A}_$V.json")); x=d["detail"]
    print("active=$A V=$V value=%.1f pairs/s ms/step=%.2f icp_ms=%.2f prep_ms=%.2f" % (d["value"], d["ms_per_step"], x["ms_icp_per_step"], x["ms_preprocess_per_step"]))
except Exception as e: print("bench $A $V failed", e)
PY
done
