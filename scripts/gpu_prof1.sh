#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b16.csv python bench.py --pairs 16 --steps 1 --warmup 3 --no-cpu-baseline --ctas-per-pair 1 > gpurun_out/ncu_b16.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b1.csv python bench.py --pairs 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_icp -s 3 -c 1 -o gpurun_out/prof_icp_b16 -f python bench.py --pairs 16 --steps 1 --warmup 3 --no-cpu-baseline --ctas-per-pair 1 > gpurun_out/ncu_icp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_knn -s 6 -c 2 -o gpurun_out/prof_knn_b16 -f python bench.py --pairs 16 --steps 1 --warmup 3 --no-cpu-baseline --ctas-per-pair 1 > gpurun_out/ncu_knn.log 2>&1
ls -la gpurun_out/
