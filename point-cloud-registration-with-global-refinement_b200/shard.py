"""Pair-level sharding across ranks (one process per GPU).

A scan pair does not shard (its ICP loop is a serial chain over a < 10 MB working set), so whole pairs are
distributed: rank r of W takes a contiguous block of the pair list (consecutive pairs share clouds, so a block needs
only one extra cloud) and the per-pair results -- 16 pose entries, fitness, RMSE -- are gathered with ONE collective at
the end (NCCL on GPUs, gloo in the CPU tests).  Nothing is exchanged inside a pair.
The reference's loop over pairs is sequential and independent per pair (2_MGICP_refinement_in_NCLT_dataset.py:187-218).
"""
from __future__ import annotations

import numpy as np
import torch

RESULT_WIDTH = 18   # 16 pose entries (row-major 4x4), fitness, inlier_rmse


def partition(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the contiguous block of rank `rank`; block sizes differ by at most one, earlier ranks get the extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_problem(pairs, rank: int, world: int):
    """Pairs of this rank re-indexed against the minimal cloud subset it has to upload.
    returns (cloud_ids, local_pairs, (lo, hi))"""
    lo, hi = partition(len(pairs), rank, world)
    mine = list(pairs[lo:hi])
    cloud_ids = sorted({c for p in mine for c in p})
    remap = {c: i for i, c in enumerate(cloud_ids)}
    return cloud_ids, [(remap[s], remap[t]) for s, t in mine], (lo, hi)


def pack_results(T, fitness, rmse) -> torch.Tensor:
    """[B,4,4], [B], [B] tensors -> [B, 18] float64 tensor on the same device"""
    B = T.shape[0]
    return torch.cat([T.reshape(B, 16), fitness.reshape(B, 1), rmse.reshape(B, 1)], dim=1).to(torch.float64)


def gather_results(local: torch.Tensor, n_total: int, rank: int, world: int) -> torch.Tensor:
    """All ranks receive the [n_total, 18] results in global pair order.  `local` is this rank's [B_local, 18] block."""
    if world == 1:
        return local
    import torch.distributed as dist
    sizes = [partition(n_total, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    # NCCL gathers device tensors; a gloo group (CPU tests, or several ranks sharing one GPU) gathers on the host
    dev = local.device if dist.get_backend() == "nccl" else torch.device("cpu")
    pad = torch.zeros((bmax, RESULT_WIDTH), dtype=torch.float64, device=dev)
    pad[: local.shape[0]] = local.to(dev)
    out = torch.empty((world * bmax, RESULT_WIDTH), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0).to(local.device)


def unpack_results(res: torch.Tensor):
    a = res.cpu().numpy()
    return a[:, :16].reshape(-1, 4, 4), a[:, 16], a[:, 17]


def register_sharded(engine, clouds, pairs, voxel_sizes, max_dists, max_iters, T_init, rank: int, world: int, opts=None):
    """Multiscale GICP of `pairs` (global list, identical on every rank) sharded over the ranks; every rank returns all results."""
    cloud_ids, local_pairs, (lo, hi) = local_problem(pairs, rank, world)
    md = np.asarray(max_dists, np.float64)
    md = md if md.ndim == 1 else md.reshape(len(pairs), -1)[lo:hi]
    T0 = np.asarray(T_init, np.float64).reshape(len(pairs), 4, 4)[lo:hi]
    if hi > lo:
        r = engine.run([clouds[c] for c in cloud_ids], local_pairs, voxel_sizes, md, max_iters, T0, opts)
        local = pack_results(torch.from_numpy(r.transformation), torch.from_numpy(r.fitness), torch.from_numpy(r.inlier_rmse))
        local = local.to(engine.tdev)
    else:
        local = torch.zeros((0, RESULT_WIDTH), dtype=torch.float64, device=engine.tdev)
    return unpack_results(gather_results(local, len(pairs), rank, world))
