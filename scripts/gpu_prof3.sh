#!/bin/bash
mkdir -p gpurun_out
for cf in 6 8 12; do
  timeout 600 python bench.py --pairs 148 --steps 3 --no-cpu-baseline --cell-factor $cf 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('knn_cell=$cf ms/step=%.2f icp_ms=%.2f prep_ms=%.2f' % (d['ms_per_step'], x['ms_icp_per_step'], x['ms_preprocess_per_step']))"
done
for f in 2 4; do
  timeout 600 python bench.py --pairs 148 --steps 3 --no-cpu-baseline --icp-cell-factor $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('icp_cell=$f ms/step=%.2f icp_ms=%.2f prep_ms=%.2f' % (d['ms_per_step'], x['ms_icp_per_step'], x['ms_preprocess_per_step']))"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_icp_tasks -s 3 -c 1 -o gpurun_out/prof_tasks3_b148 -f python bench.py --pairs 148 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_tasks3.log 2>&1
ls -la gpurun_out/*.ncu-rep
