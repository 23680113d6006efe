#!/bin/bash
# gpu tests, then full ncu captures (with source) of the ICP task kernel and the outlier-filter kNN at 148 pairs
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || { timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log; }
timeout 600 python bench.py --pairs 148 --steps 3 --no-cpu-baseline > gpurun_out/bench_148.json 2> gpurun_out/bench_148.err; python -c "
import json; d=json.load(open('gpurun_out/bench_148.json')); x=d['detail']; print('148 pairs: %.1f pairs/s icp %.2f prep %.2f' % (d['value'], x['ms_icp_per_step'], x['ms_preprocess_per_step']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_icp_tasks -s 3 -c 1 -o gpurun_out/prof_tasks_b148 -f python bench.py --pairs 148 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_tasks.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn -s 3 -c 1 -o gpurun_out/prof_knn2_b148 -f python bench.py --pairs 148 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_knn2.log 2>&1
bash scripts/gpu_launches.sh 148
ls -la gpurun_out/*.ncu-rep
