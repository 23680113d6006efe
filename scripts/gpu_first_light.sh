#!/bin/bash
# first GPU contact: sanity + verbose parity run (scratch helper, results under gpurun_out/)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/fl.log 2>&1
python -c "import open3d" >> gpurun_out/fl.log 2>&1 || echo "open3d: not importable on the GPU box" >> gpurun_out/fl.log
nproc >> gpurun_out/fl.log
timeout 600 python __graft_entry__.py smoke >> gpurun_out/fl.log 2>&1
echo "smoke rc=$?" >> gpurun_out/fl.log
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -s >> gpurun_out/fl.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fl.log
tail -60 gpurun_out/fl.log
