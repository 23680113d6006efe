#!/bin/bash
# round 2: whole GPU suite with -s (full log kept), then the bench (default flags) and the single-pair latency mode
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -s --tb=short > gpurun_out/r2_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_full.log
grep -v "^$" gpurun_out/r2_pytest_full.log | tail -${TAILN:-45}
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench.json"))
    x = d["detail"]
    print("value %.1f e2e %.1f ms/step %.2f icp %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"]))
    for k in ("stage_ms_per_step", "single_pair_ms", "config3", "config5", "parity", "cpu_stage_split", "preprocess_roofline"):
        print(k, json.dumps(x.get(k)))
    print("roofline", json.dumps(d["roofline"]))
    print("cpu", json.dumps(d.get("cpu_baseline")))
except Exception as e:
    print("bench parse failed", e)
PY
fi
