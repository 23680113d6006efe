#!/bin/bash
# A/B in one box session: "lib:pairs:mode" triples; lib = new | a path relative to the repo root
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || { timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log; }
for cfg in $SWEEP; do
  IFS=: read LIBSEL P G <<< "$cfg"
  if [ "$LIBSEL" = "new" ]; then unset MGICP_LIB; else export MGICP_LIB=$PWD/$LIBSEL; fi
  tag=$(basename $LIBSEL .so)_${P}_$G
  timeout 600 python bench.py --pairs $P --steps ${STEPS:-3} --no-cpu-baseline --ctas-per-pair=$G $BENCH_EXTRA > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || tail -5 gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$tag.json")); x=d["detail"]
    print("lib=$LIBSEL pairs=$P mode=$G value=%.1f pairs/s e2e=%.1f ms/step=%.2f icp_ms=%.2f prep_ms=%.2f launches=%d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"], x["ms_preprocess_per_step"] or 0.0, d["gpu_launches"]))
except Exception as e: print("bench $tag failed", e)
PY
done
