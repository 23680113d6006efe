#!/bin/bash
# round-2 evidence at the bench's default workload (296 pairs per GPU): launch list + full ncu captures of the dominant kernels
mkdir -p gpurun_out
P=${PAIRS:-296}
nvidia-smi -L > gpurun_out/box.log; nproc >> gpurun_out/box.log
timeout 900 bash scripts/gpu_launches.sh $P --no-extras
for K in k_icp_tasks k_knn_hist; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${K}_b$P -f python bench.py --pairs $P --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_$K.log 2>&1
done
bash scripts/gpu_fgr_prof.sh
ls -la gpurun_out/*.ncu-rep
