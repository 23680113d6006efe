#!/bin/bash
for g in 16 24 32 48 64 96 128 148; do
  python bench.py --pairs 1 --steps 5 --no-cpu-baseline --ctas-per-pair $g 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('gang=$g ms/step=%.2f icp_ms=%.2f prep_ms=%.2f iters=%s' % (d['ms_per_step'], x['ms_icp_per_step'], x['ms_preprocess_per_step'], x['iterations_mean_per_scale']))"
done
