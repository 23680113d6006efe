#!/bin/bash
# N-GPU check: gpu tests, default bench at N=1, torchrun bench at N=$1, reference arm
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_box.log; nproc >> gpurun_out/multi_box.log
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -5 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
