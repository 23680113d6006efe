#!/bin/bash
# iteration helper: gpu tests + bench sweeps over the ICP scheduling mode (ctas-per-pair: >0 static gang, <0 task mode)
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
for cfg in ${SWEEP:-148:1 148:0 148:-1 148:-2 148:-8 1:0 16:0 37:0 74:0 296:0}; do
  P=${cfg%%:*}; G=${cfg##*:}
  timeout 600 python bench.py --pairs $P --steps 3 --no-cpu-baseline --ctas-per-pair=$G $BENCH_EXTRA > gpurun_out/bench_${P}_$G.json 2> gpurun_out/bench_${P}_$G.err || tail -5 gpurun_out/bench_${P}_$G.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${P}_$G.json")); x=d["detail"]
    print("pairs=$P mode=$G value=%.1f pairs/s e2e=%.1f ms/step=%.2f icp_ms=%.2f prep_ms=%.2f launches=%d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"], x["ms_preprocess_per_step"], d["gpu_launches"]))
except Exception as e: print("bench $P $G failed", e)
PY
done
