#!/bin/bash
# regular iteration: gpu tests + three bench points
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
for cfg in "1" "16" "148"; do
  timeout 600 python bench.py --pairs $cfg --steps 3 --no-cpu-baseline $BENCH_EXTRA > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err || tail -5 gpurun_out/bench_$cfg.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$cfg.json")); x=d["detail"]
    print("pairs=$cfg value=%.1f pairs/s e2e=%.1f ms/step=%.2f icp_ms=%.2f prep_ms=%.2f iters=%s launches=%d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"], x["ms_preprocess_per_step"], [round(v,1) for v in x["iterations_mean_per_scale"]], d["gpu_launches"]))
except Exception as e: print("bench $cfg failed", e)
PY
done
