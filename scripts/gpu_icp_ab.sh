#!/bin/bash
# ICP A/B in one box session. SWEEP="lib:chunk_points:pairs[:task_blocks_per_sm] ..." (lib = new | path relative to the repo root; chunk_points 0 = default)
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -8
for cfg in $SWEEP; do
  IFS=: read LIBSEL CP P OCC <<< "$cfg"
  if [ -z "$OCC" ] || [ "$OCC" = "0" ]; then unset MGICP_TASK_OCC; else export MGICP_TASK_OCC=$OCC; fi
  if [ "$LIBSEL" = "new" ]; then unset MGICP_LIB; else export MGICP_LIB=$PWD/$LIBSEL; fi
  if [ "$CP" = "0" ]; then unset MGICP_CHUNK_POINTS; else export MGICP_CHUNK_POINTS=$CP; fi
  tag=$(basename $LIBSEL .so)_${CP}_${P}_$OCC
  timeout 600 python bench.py --pairs $P --steps ${STEPS:-3} --no-cpu-baseline --no-extras $BENCH_EXTRA > gpurun_out/icp_$tag.json 2> gpurun_out/icp_$tag.err || tail -5 gpurun_out/icp_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/icp_$tag.json")); x=d["detail"]
    print("lib=$LIBSEL chunk=$CP pairs=$P occ=$OCC value=%.1f e2e=%.1f ms/step=%.2f icp_ms=%.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"]))
except Exception as e: print("failed $cfg", e)
PY
done
