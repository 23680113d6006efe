#!/bin/bash
# kNN A/B in one box session: stage tests (bit parity of the outlier filter / normals), then stage timings per library / mode / cell factor
# SWEEP="lib:mode:cellfactor ..." (lib = new | path relative to the repo root)
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -8
for cfg in ${SWEEP:-new:0:0 new:1:0}; do
  IFS=: read LIBSEL MODE CF <<< "$cfg"
  if [ "$LIBSEL" = "new" ]; then unset MGICP_LIB; else export MGICP_LIB=$PWD/$LIBSEL; fi
  tag=$(basename $LIBSEL .so)_${MODE}_${CF}
  MGICP_KNN_MODE=$MODE timeout 600 python bench.py --pairs ${PAIRS:-296} --steps 3 --no-cpu-baseline --config3-pairs 0 --config5-pairs 0 --cell-factor $CF > gpurun_out/knn_$tag.json 2> gpurun_out/knn_$tag.err || tail -5 gpurun_out/knn_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/knn_$tag.json")); x=d["detail"]; s=x["stage_ms_per_step"]
    print("lib=$LIBSEL knn_mode=$MODE cell_factor=$CF value=%.1f e2e=%.1f ms/step=%.2f | sor %.2f normals %.2f icp %.2f sum %.2f | single pair %.3f ms" % (d["value"], d["e2e"]["value"], d["ms_per_step"], s["sor_ms"], s["normals_ms"], s["icp_ms"], s["serialised_sum_ms"], x["single_pair_ms"]["median"]))
except Exception as e: print("failed $cfg", e)
PY
done
