#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --pairs 16 --steps 2 --warmup 3 --cpu-sample 2 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?" >> gpurun_out/bench_small.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
timeout 600 python bench.py --pairs 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err
tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json; cat gpurun_out/bench_single.json; tail -3 gpurun_out/bench_single.err
