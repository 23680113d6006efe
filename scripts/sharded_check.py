"""Worker of tests/test_gpu_sharded.py: shard.register_sharded on a pair list shared by all ranks (run under torchrun).
Every rank registers its contiguous block on its own GPU, the poses are all-gathered over NCCL; rank 0 checks the gathered
result against the whole list registered on one GPU and prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mgicp_b200 as m  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    az = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    ngpu = torch.cuda.device_count()
    nccl = ngpu >= world                     # one GPU per rank: NCCL; fewer GPUs than ranks (the single-GPU test box): the ranks share
    local = local % ngpu                     # GPUs and the poses are gathered through gloo on the host
    torch.cuda.set_device(local)
    if nccl:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    scene = m.synthetic.Scene(seed=12345)
    scans = [m.synthetic.make_scan(scene, m.synthetic.sensor_pose(k), az, seed=k).astype(np.float32) for k in range(n_pairs + 1)]
    pairs = [(i + 1, i) for i in range(n_pairs)]
    inits = []
    for i in range(n_pairs):
        T_true = np.linalg.inv(m.synthetic.sensor_pose(i)) @ m.synthetic.sensor_pose(i + 1)
        inits.append(m.synthetic.perturbation(np.random.default_rng(77 + i)) @ T_true)
    inits = np.stack(inits)
    vox, dists = [1.0, 0.5, 0.25], [3.0, 1.0, 0.25]
    eng = m.Engine(local)
    opts = eng.make_opts(loss="l1", ctas_per_pair=-2)       # fixed chunking: the same summation order whatever the block of pairs
    dist.barrier()
    t0 = time.perf_counter()
    T, fit, rm = m.shard.register_sharded(eng, scans, pairs, vox, dists, 100, inits, rank, world, opts)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        ref = eng.run(scans, pairs, vox, dists, 100, inits, opts)
        out = {"pairs": n_pairs, "world": world, "backend": dist.get_backend(), "gpus": ngpu, "seconds": dt, "equal_T": bool(np.array_equal(T, ref.transformation)),
               "equal_fitness": bool(np.array_equal(fit, ref.fitness)), "equal_rmse": bool(np.array_equal(rm, ref.inlier_rmse)),
               "max_abs_dT": float(np.abs(T - ref.transformation).max())}
        print("SHARDED " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
