#!/bin/bash
# end-of-round run: whole GPU suite with -s, smoke, both bench arms the way the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -s --tb=short > gpurun_out/r2_pytest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log
tail -4 gpurun_out/r2_pytest_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_final.json")); x = d["detail"]
print("value %.1f e2e %.1f ms/step %.2f icp %.2f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"], d["gpu_launches"], d["clocks"]))
print("stage", json.dumps(x["stage_ms_per_step"]))
print("single", x["single_pair_ms"]["median"], "config3", x["config3"]["pairs_per_s"], "config5", x["config5"]["pairs_per_s"], "fgr", x["fgr_front_end"]["pairs_per_s"], x["fgr_front_end"]["ms_features_device"], x["fgr_front_end"]["ms_registration_device"])
print("parity", json.dumps(x["parity"]["gpu_vs_oracle"]))
print("cpu", d["cpu_baseline"]["value"])
PY
