#!/bin/bash
for p in 148 296 444 592; do
  python bench.py --pairs $p --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); x=d['detail']
print('pairs=$p value=%.1f e2e=%.1f ms/step=%.2f icp_ms=%.2f prep_ms=%.2f gen_s=%.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], x['ms_icp_per_step'], x['ms_preprocess_per_step'], x['workload_gen_s']))"
done
