"""Pair sharding across ranks, exercised with world_size 2 on CPU (gloo).  The per-pair work is stubbed by the CPU oracle
(test infrastructure) so the exchange logic -- partition, re-indexing, gather order -- is what is tested."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_partition_covers_everything(pkg):
    for n in (0, 1, 7, 8, 1000, 10001):
        for w in (1, 2, 3, 8):
            blocks = [pkg.shard.partition(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    pairs = [(i + 1, i) for i in range(10)]
    ids, local, (lo, hi) = pkg.shard.local_problem(pairs, 1, 2)
    assert (lo, hi) == (5, 10) and ids == [5, 6, 7, 8, 9, 10] and local == [(i + 1, i) for i in range(5)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import mgicp_b200 as m
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scans, inits, _ = m.synthetic.make_sequence(n_pairs + 1, azimuth_steps=60, seed=2)
    pairs = [(i + 1, i) for i in range(n_pairs)]
    ids, local_pairs, (lo, hi) = m.shard.local_problem(pairs, rank, world)
    rows = []
    for (s, t), T0 in zip(local_pairs, inits[lo:hi]):
        r = oracle.multiscale_gicp(scans[ids[s]], scans[ids[t]], [1.0], [3.0], 5, T0, loss="l2")
        rows.append(np.concatenate([r.transformation.reshape(16), [r.fitness, r.inlier_rmse]]))
    local = torch.tensor(np.array(rows).reshape(-1, 18), dtype=torch.float64)
    allres = m.shard.gather_results(local, n_pairs, rank, world)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), allres.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 4])
def test_world_size_2_gather_matches_single_process(pkg, oracle, tmp_path, n_pairs):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(tmp_path / f"rank{r}.npy") for r in range(world))
    assert a.shape == (n_pairs, 18) and np.array_equal(a, b)          # every rank holds all results, same order
    scans, inits, _ = pkg.synthetic.make_sequence(n_pairs + 1, azimuth_steps=60, seed=2)
    for i in range(n_pairs):
        r = oracle.multiscale_gicp(scans[i + 1], scans[i], [1.0], [3.0], 5, inits[i], loss="l2")
        assert np.array_equal(a[i, :16].reshape(4, 4), r.transformation) and a[i, 16] == r.fitness
