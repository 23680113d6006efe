#!/bin/bash
# round-end evidence at the bench's default workload (296 pairs per GPU): tests, smoke, bench (+ reference arm), launch list,
# full ncu captures of the two dominant kernels
mkdir -p gpurun_out
P=${PAIRS:-296}
nvidia-smi -L > gpurun_out/box.log; nproc >> gpurun_out/box.log
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -5 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
timeout 600 python bench.py --pairs 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; cat gpurun_out/bench_single.json
timeout 900 bash scripts/gpu_launches.sh $P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_icp_tasks -s 3 -c 1 -o gpurun_out/prof_tasks_b$P -f python bench.py --pairs $P --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_tasks.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn -s 3 -c 1 -o gpurun_out/prof_knn_b$P -f python bench.py --pairs $P --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_knn.log 2>&1
ls -la gpurun_out/*.ncu-rep
