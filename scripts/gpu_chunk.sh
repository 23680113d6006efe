#!/bin/bash
# sweep of the adaptive chunk size (points per chunk) of the ICP task mode; "pairs:chunk_points"
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || { timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log; }
for cfg in $SWEEP; do
  IFS=: read P CP <<< "$cfg"
  export MGICP_CHUNK_POINTS=$CP
  timeout 600 python bench.py --pairs $P --steps 3 --no-cpu-baseline > gpurun_out/chunk_${P}_$CP.json 2> gpurun_out/chunk_${P}_$CP.err || tail -5 gpurun_out/chunk_${P}_$CP.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/chunk_${P}_$CP.json")); x=d["detail"]
    print("pairs=$P chunk_points=$CP value=%.1f pairs/s ms/step=%.2f icp_ms=%.2f prep_ms=%.2f" % (d["value"], d["ms_per_step"], x["ms_icp_per_step"], x["ms_preprocess_per_step"] or 0.0))
except Exception as e: print("bench $P $CP failed", e)
PY
done
