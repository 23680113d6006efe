#!/bin/bash
# N GPUs of one box: the sharded test, then the bench the way the driver launches it
N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -s -q --tb=short 2>&1 | tail -6
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n$N.json")); x = d["detail"]
    print("N=$N value %.1f e2e %.1f ms/step %.2f icp %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], x["ms_icp_per_step"]))
    for k in ("config3", "config5"):
        print(k, json.dumps(x.get(k)))
except Exception as e:
    print("parse failed", e)
PY
