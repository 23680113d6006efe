#!/bin/bash
mkdir -p gpurun_out
MGICP_WS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_icp_tasks_ws -s 3 -c 1 -o gpurun_out/prof_ws -f python bench.py --pairs ${PAIRS:-148} --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_ws.log 2>&1
ls -la gpurun_out/prof_ws.ncu-rep
