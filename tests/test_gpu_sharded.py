"""shard.register_sharded on real GPUs: a pair list shared by two ranks, each registering its contiguous block, poses
all-gathered, equal bit for bit to the whole list registered by one rank.  With two GPUs (gpurun --gpus 2) every rank has its own
B200 and the gather is NCCL; on a single-GPU box the two ranks share the GPU (two engines) and gather through gloo, so the test
never skips.  (The CPU-only version of the plumbing is tests/test_shard.py.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_register_sharded_two_gpus():
    import torch
    assert torch.cuda.device_count() >= 1
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "scripts", "sharded_check.py"), "200", "600"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("SHARDED ")][-1]
    out = json.loads(line[len("SHARDED "):])
    print(out)
    assert out["equal_T"] and out["equal_fitness"] and out["equal_rmse"], out
