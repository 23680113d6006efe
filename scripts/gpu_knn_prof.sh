#!/bin/bash
# launch list + full ncu capture of k_knn_cw at a small batch (analysis run; numbers under ncu are not bench values)
mkdir -p gpurun_out
P=${PAIRS:-74}
bash scripts/gpu_launches.sh $P --no-extras | grep -E "knn|sor|normals|total"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_cw -s 3 -c 1 -o gpurun_out/prof_knn_cw_b$P -f python bench.py --pairs $P --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_knn_cw.log 2>&1
ls -la gpurun_out/*.ncu-rep
