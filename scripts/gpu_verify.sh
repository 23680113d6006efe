#!/bin/bash
# quick verification of the current tree on one B200: gpu tests, smoke, default bench, single-pair bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.log; nproc >> gpurun_out/box.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --tb=short ) 2>&1 | tail -15 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 600 python bench.py --pairs 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; cat gpurun_out/bench_single.json
