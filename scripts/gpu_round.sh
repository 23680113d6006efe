#!/bin/bash
# round checkpoint: gpu tests, smoke, bench at 1/148 pairs, launch lists, full ncu capture of k_icp and k_knn
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.log; nproc >> gpurun_out/box.log
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 600 python bench.py --pairs 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; cat gpurun_out/bench_single.json
timeout 600 bash scripts/gpu_launches.sh 148
timeout 600 bash scripts/gpu_launches.sh 1
timeout 900 bash scripts/gpu_prof.sh k_icp 3 148 prof_icp_b148
timeout 900 bash scripts/gpu_prof.sh k_knn 6 148 prof_knn_b148
ls -la gpurun_out
