#!/bin/bash
for f in 1.5 2 2.5 3 4; do
  python bench.py --pairs 74 --steps 3 --no-cpu-baseline --icp-cell-factor $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('icp_cell=$f ms/step=%.2f icp_ms=%.2f prep_ms=%.2f' % (d['ms_per_step'], x['ms_icp_per_step'], x['ms_preprocess_per_step']))"
done
for f in 5 6 8 10 12; do
  python bench.py --pairs 74 --steps 3 --no-cpu-baseline --cell-factor $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('knn_cell=$f ms/step=%.2f icp_ms=%.2f prep_ms=%.2f' % (d['ms_per_step'], x['ms_icp_per_step'], x['ms_preprocess_per_step']))"
done
