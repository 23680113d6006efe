#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -s -q --tb=short 2>&1 | tail -4
for sp in 1 2 3; do
  timeout 600 python bench.py --steps 3 --no-cpu-baseline --config3-pairs 125 --config5-pairs 0 --split $sp > gpurun_out/split_$sp.json 2> gpurun_out/split_$sp.err
  python - <<PY
import json
d=json.load(open("gpurun_out/split_$sp.json")); c=d["detail"]["config3"]
print("split $sp: 125 pairs in %.1f ms = %.0f pairs/s, batches %d" % (1e3*c["seconds"], c["pairs_per_s"], c["batches_on_rank0"]))
PY
done
