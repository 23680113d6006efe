#!/bin/bash
# round 2, first call: first run of the FGR registration kernels (bounded), then the whole GPU suite with -s, full log kept
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_box.log; nproc >> gpurun_out/r2_box.log
python -c "import open3d" >> gpurun_out/r2_box.log 2>&1 || echo "open3d: not importable on the GPU box" >> gpurun_out/r2_box.log
MGICP_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_fgr.py -m gpu -s --tb=short > gpurun_out/r2_fgr_first.log 2>&1
echo "fgr rc=$?" >> gpurun_out/r2_fgr_first.log
tail -40 gpurun_out/r2_fgr_first.log
timeout 1200 python -m pytest tests -m gpu -s --tb=short --deselect tests/test_gpu_fgr.py > gpurun_out/r2_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_full.log
tail -8 gpurun_out/r2_pytest_full.log
