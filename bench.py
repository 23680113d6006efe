#!/usr/bin/env python
"""bench.py -- 3-scale GICP scan-pairs/sec on synthetic NCLT-shaped (~100k-point, HDL-32 pattern) pairs.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched under torch.distributed.run by the driver)
    python bench.py --impl reference ...                    (CPU arm: the oracle restatement of the Open3D path)

A step = one pass of the hot path (per-scale voxel down-sample, outlier removal, normals, 3-scale ICP loops) over one
batch of `--pairs` consecutive scan pairs per GPU (BASELINE.json configs[1] batched as in configs[2]); voxels
1.0/0.5/0.25 m, max correspondence distance 3x/2x/1x voxel, 100 iterations per scale, L1 kernel (the reference's
setting).  Whole pairs are sharded across ranks (weak scaling); only the resulting poses are gathered (NCCL).

`value`: inputs resident in HBM, CUDA events.  `e2e`: the package's streaming API (mgicp_b200.BatchStream) fed with the
pageable numpy clouds a caller of the reference holds (S2:169), results back in host memory, wall clock.
`detail` carries what BASELINE.json's other configs and the north-star target ask for: the latency of a single pair
(`single_pair_ms`), configs[2] = 1,000 fixed consecutive pairs partitioned over the ranks (`config3`, strong scaling),
configs[4] = 10,000 non-consecutive pairs with large initial offsets (`config5`), the per-stage split of a step on the GPU
and on the CPU, and the parity of the GPU poses against the CPU oracle on the pairs the oracle is timed on (`parity`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOXELS = [1.0, 0.5, 0.25]
DISTS = [3.0, 1.0, 0.25]
MAX_IT = 100
AZIMUTH = 3125          # 32 beams x 3125 firings ~ 100k points per scan


def _gen_scan(args):
    import mgicp_b200 as m
    k, seed, az = args
    scene = m.synthetic.Scene(seed=12345)
    return m.synthetic.make_scan(scene, m.synthetic.sensor_pose(k), az, seed=1000003 * seed + k).astype(np.float32)


class ScanCache:
    """scans of the synthetic circuit by global index, generated on demand by a process pool"""

    def __init__(self, az, seed=0):
        self.az, self.seed, self.scans = az, seed, {}

    def need(self, ks):
        import multiprocessing as mp
        todo = sorted(set(int(k) for k in ks) - set(self.scans))
        if not todo:
            return
        nproc = max(1, min(len(todo), (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
        with mp.get_context("fork").Pool(nproc) as pool:
            for k, s in zip(todo, pool.map(_gen_scan, [(k, self.seed, self.az) for k in todo])):
                self.scans[k] = s

    def get(self, ks):
        self.need(ks)
        return [self.scans[int(k)] for k in ks]


def consecutive_inits(k0, n_pairs, seed=0):
    """FGR-like initial poses and true motions of the pairs (k0+i+1 -> k0+i)"""
    import mgicp_b200 as m
    inits, truths = [], []
    for i in range(n_pairs):
        T_true = np.linalg.inv(m.synthetic.sensor_pose(k0 + i)) @ m.synthetic.sensor_pose(k0 + i + 1)
        rng = np.random.default_rng(77 + 1000003 * seed + k0 + i)
        inits.append(m.synthetic.perturbation(rng) @ T_true)
        truths.append(T_true)
    return np.stack(inits), np.stack(truths)


def make_workload(cache, n_pairs, rank):
    """n_pairs+1 consecutive scans (float32, the PCD-native dtype) + FGR-like initial poses; rank-specific stretch of the circuit"""
    k0 = rank * (n_pairs + 1)
    scans = cache.get(range(k0, k0 + n_pairs + 1))
    inits, truths = consecutive_inits(k0, n_pairs)
    pairs = [(i + 1, i) for i in range(n_pairs)]      # source = scan i+1, target = scan i (S2:191)
    return scans, pairs, inits, truths


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1 to every rank: not what a CPU arm wants)
    try:
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        oracle.set_num_threads(os.cpu_count() or 1)
    return oracle


def oracle_run(scans, pairs, inits, n_sample, sum_chunk=1024):
    """the CPU restatement of the reference's Open3D path, all host threads, on the first n_sample pairs of the workload.
    returns (pairs/s, seconds, threads, results, stage split)"""
    oracle = _oracle()
    oracle.set_sum_chunk(sum_chunk)
    oracle.stage_reset()
    res = []
    t0 = time.perf_counter()
    try:
        for (s, t), T0 in list(zip(pairs, inits))[:n_sample]:
            res.append(oracle.multiscale_gicp(scans[s].astype(np.float64), scans[t].astype(np.float64), VOXELS, DISTS, MAX_IT, T0, loss="l1"))
    finally:
        oracle.set_sum_chunk(1024)
    dt = time.perf_counter() - t0
    st = oracle.stage_seconds()
    S = len(VOXELS)
    split = {"per_pair_ms": {k[:-2]: 1e3 * st[k] / n_sample for k in ("downsample_s", "sor_s", "normals_s", "icp_s")},
             "ms_per_iteration_per_scale": [1e3 * st["icp_s_per_scale"][s] / max(st["icp_passes_per_scale"][s], 1.0) for s in range(S)],
             "iterations_mean_per_scale": [st["icp_passes_per_scale"][s] / n_sample - 1.0 for s in range(S)]}
    return n_sample / dt, dt, oracle.num_threads(), res, split


def parity_block(m, T_gpu, fit_gpu, rm_gpu, ref, ref2):
    """GPU vs the faithful oracle on the same pairs, next to the oracle's own envelope (its sums re-associated)"""
    def deltas(TA, fA, rA, TB, fB, rB):
        e = np.array([m.synthetic.pose_error(a, b) for a, b in zip(TA, TB)])
        df, dr = np.abs(np.asarray(fA) - np.asarray(fB)), np.abs(np.asarray(rA) - np.asarray(rB))
        inside = (e[:, 0] < 1e-4) & (e[:, 1] < 1e-4) & (df < 1e-5) & (dr < 1e-5)
        q = lambda x: {"median": float(np.median(x)), "p90": float(np.quantile(x, 0.9)), "max": float(np.max(x))}
        return {"rot_rad": q(e[:, 0]), "trans_m": q(e[:, 1]), "fitness": q(df), "rmse": q(dr), "frac_within_north_star": float(inside.mean())}
    n = len(ref)
    To, fo, ro = [r.transformation for r in ref], [r.fitness for r in ref], [r.inlier_rmse for r in ref]
    out = {"n": n, "tolerances": "1e-4 rad, 1e-4 m, 1e-5 fitness, 1e-5 rmse (north_star)",
           "gpu_vs_oracle": deltas(T_gpu[:n], fit_gpu[:n], rm_gpu[:n], To, fo, ro)}
    if ref2:
        n2 = len(ref2)
        out["oracle_self_envelope"] = dict(deltas([r.transformation for r in ref2], [r.fitness for r in ref2], [r.inlier_rmse for r in ref2],
                                                  To[:n2], fo[:n2], ro[:n2]), n=n2,
                                           how="the same oracle with its 27 sums re-associated (chunk 333 instead of 1024): the L1-IRLS "
                                               "loop is chaotic, this is the scatter any independently ordered implementation shows")
    return out


def ncu_traffic(pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed ncu --set full
    capture of this exact workload (profiles/traffic.json); null for other batch sizes."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return float(t["dram_bytes_per_launch"]) if int(t["pairs_per_gpu"]) == int(pairs) else None
    except Exception:
        return None


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def emit(line: dict, fd: int) -> None:
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    # exactly ONE line goes to stdout: libraries that print there (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=296, help="scan pairs per GPU per step (two per SM: the tail of a batch, when few long-running pairs are left, is amortised)")
    ap.add_argument("--azimuth", type=int, default=AZIMUTH)
    ap.add_argument("--cpu-sample", type=int, default=64, help="pairs timed on the CPU oracle for cpu_baseline (~10 s)")
    ap.add_argument("--ctas-per-pair", type=int, default=0)
    ap.add_argument("--cell-factor", type=float, default=0.0)
    ap.add_argument("--icp-cell-factor", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip detail.single_pair_ms / config3 / config5 / stage split (A/B runs)")
    ap.add_argument("--config3-pairs", type=int, default=1000)
    ap.add_argument("--config5-pairs", type=int, default=10000)
    ap.add_argument("--split", type=int, default=1, help="a rank's block of a fixed pair list that fits ONE batch is still streamed in this many batches, so that packing / upload of one overlap the kernels of the other")
    ap.add_argument("--single-engine", action="store_true", help="one workspace / stream instead of two alternating ones")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{a.pairs} consecutive NCLT-shaped synthetic HDL-32 scan pairs per GPU (~{32 * a.azimuth // 1000}k pts/scan), "
                          "3-scale GICP, voxel 1.0/0.5/0.25 m, max-dist 3.0/1.0/0.25 m, 100 it/scale, SOR(30,1.0), kNN-20 normals, L1 kernel",
              "pairs_per_gpu": a.pairs, "points_per_scan": 32 * a.azimuth, "l2_policy": "inputs_larger_than_L2 (no flush)",
              "parallelism": f"pairs sharded over {world} GPU(s), poses gathered"}
    cache = ScanCache(a.azimuth)

    # ------------------------------------------------------------------ reference arm (CPU) ----------
    if a.impl == "reference":
        if rank != 0:
            return
        n_s = max(2, min(a.pairs, 32))        # ~5 s of CPU work per step
        scans, pairs, inits, _ = make_workload(cache, n_s, 0)
        for _ in range(min(a.warmup, 1)):
            oracle_run(scans, pairs, inits, 1)
        times = []
        cores = 1
        split = None
        for _ in range(a.steps):
            pps, dt, cores, _, split = oracle_run(scans, pairs, inits, n_s)
            times.append(dt)
        tot = sum(times)
        val = n_s * a.steps / tot
        sample = f"{n_s} pairs of the same workload per step, {a.steps} steps, CPU oracle (C restatement of the Open3D path; Open3D itself is not installable offline)"
        emit({"impl": "reference", "metric": "3-scale GICP scan-pairs/sec (~100k pts)", "value": val, "unit": "pairs/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "detail": {"stage_split": split}}, out_fd)
        return

    # ------------------------------------------------------------------ B200 arm ------------------------
    import torch
    import mgicp_b200 as m
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    eng = m.Engine(local_rank)
    opts = eng.make_opts(loss="l1", ctas_per_pair=a.ctas_per_pair, cell_factor=a.cell_factor, icp_cell_factor=a.icp_cell_factor)
    t_gen = time.perf_counter()
    scans, pairs, inits, truths = make_workload(cache, a.pairs, rank)
    t_gen = time.perf_counter() - t_gen
    flat, off, _ = eng.pack_clouds(scans)
    B, S = len(pairs), len(VOXELS)
    ps, pt = [p[0] for p in pairs], [p[1] for p in pairs]
    md = np.broadcast_to(np.asarray(DISTS), (B, S)).copy()
    mi = np.full(S, MAX_IT, np.int32)
    xyz_dev = torch.from_numpy(flat).to(dev)
    T0_dev = torch.from_numpy(inits.reshape(B, 16).copy()).to(dev)
    gathered = [torch.empty((world * B, 18), dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None

    # Consecutive batches alternate between two engines (two workspaces) on two streams: the preprocessing kernels of
    # batch k+1 fill the SMs that the persistent ICP kernel of batch k leaves idle towards its end.
    # (few pairs per step = latency mode: one engine, so that ms_per_step is the latency of one batch)
    engs = [eng] if (a.single_engine or a.pairs * 8 <= 148) else [eng, m.Engine(local_rank)]
    work = [torch.cuda.Stream(device=dev) for _ in engs]

    def step_device(k):
        e = k % len(engs)
        with torch.cuda.stream(work[e]):
            engs[e].preprocess_device(xyz_dev, off, VOXELS, opts)
            ev_a.record()
            out = engs[e].register_device(ps, pt, md, mi, T0_dev, opts)
            ev_b.record()
            if world > 1:
                T, fit, rm = out[0], out[1], out[2]
                local = torch.cat([T.reshape(B, 16), fit[:, None], rm[:, None]], dim=1)
                dist.all_gather_into_tensor(gathered[e], local)
        return out

    def join_work():
        cur = torch.cuda.current_stream(dev)
        for w in work:
            ev = torch.cuda.Event()
            ev.record(w)
            cur.wait_event(ev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(a.warmup * len(engs)):
        out = step_device(k)
    torch.cuda.synchronize()
    for e_ in engs:
        e_.check()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sum(e_.kernel_launches() for e_ in engs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    icp_ms = []
    barrier()
    e0.record()
    for w in work:
        w.wait_event(e0)
    for k in range(a.steps):
        out = step_device(k)
        icp_ms.append((ev_a, ev_b))          # per-launch duration of the dominant kernel (events on the launching stream; no host sync here)
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    join_work()
    e1.record()
    barrier()
    launches = sum(e_.kernel_launches() for e_ in engs) - launches0
    ms = max_over_ranks(e0.elapsed_time(e1))
    icp = [x.elapsed_time(y) for x, y in icp_ms]

    # ---- e2e: the package's streaming API, pageable host clouds in, host results out, every step -------------------
    def gather_post(res):
        full = torch.empty((world * res.shape[0], res.shape[1]), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(full, res.contiguous())
        return full
    bs = m.BatchStream(VOXELS, DISTS, MAX_IT, engines=engs, opts=opts, post=gather_post if world > 1 else None)
    batch = (scans, pairs, inits)            # list of pageable float32 numpy arrays, as np.asarray(pcd.points) hands them over
    for _ in bs.run([batch] * 2):
        pass
    barrier()
    h2d0, d2h0 = bs.h2d_bytes, bs.d2h_bytes
    t0 = time.perf_counter()
    for res in bs.run([batch] * a.steps):
        pass
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d_step, d2h_step = (bs.h2d_bytes - h2d0) // a.steps, (bs.d2h_bytes - d2h0) // a.steps
    bs.close()
    clocks = sampler.stop() if rank == 0 else None

    T, fit, rm, it, nc, st = (x.cpu().numpy() for x in out)
    for e_ in engs:
        e_.check()
    assert np.array_equal(res.transformation[rank * B:(rank + 1) * B], T), "streaming API and device-resident path disagree"
    err = [m.synthetic.pose_error(T[b], truths[b]) for b in range(B)]

    # ---- per-stage split of one step (one engine, events at the stage boundaries; outside the timed regions) --------
    detail_extra = {}
    if not a.no_extras:
        eng.set_timing(True)
        stage = []
        for _ in range(2):
            eng.preprocess_device(xyz_dev, off, VOXELS, opts)
            eng.register_device(ps, pt, md, mi, T0_dev, opts)
            stage.append(eng.get_timing())
        eng.set_timing(False)
        stg = stage[-1]
        serial = sum(stg[k] for k in ("downsample_ms", "knn_grid_ms", "sor_ms", "normals_ms", "icp_grid_ms", "icp_ms"))
        detail_extra["stage_ms_per_step"] = {k: stg[k] for k in ("downsample_ms", "knn_grid_ms", "sor_ms", "normals_ms", "icp_grid_ms", "icp_ms")}
        detail_extra["stage_ms_per_step"]["serialised_sum_ms"] = serial
        detail_extra["stage_ms_per_pair"] = {k: v / B for k, v in detail_extra["stage_ms_per_step"].items()}
        detail_extra["batch_scale_wall_ms_mean"] = stg["scale_ms"][:S]

        # ---- north-star target: one ~100k-point pair, 3 scales, under 5 ms (gang mode, device-resident inputs) -------
        n2 = int(off[2])
        xyz1, off1 = xyz_dev[:n2], off[:3].copy()
        T01 = T0_dev[:1]
        one = []
        eng.set_timing(True)
        for r_ in range(3 + 20):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            eng.preprocess_device(xyz1, off1, VOXELS, opts)
            o1 = eng.register_device([1], [0], md[:1], mi, T01, opts)
            eb.record()
            eb.synchronize()
            if r_ >= 3:
                one.append(ea.elapsed_time(eb))
        st1 = eng.get_timing()
        eng.set_timing(False)
        it1 = o1[3].cpu().numpy()[0]
        d_rot, d_tr = m.synthetic.pose_error(o1[0].cpu().numpy()[0], T[0])
        detail_extra["single_pair_ms"] = {"median": statistics.median(one), "mean": sum(one) / len(one), "min": min(one), "max": max(one),
                                          "reps": len(one), "target_ms": 5.0, "mode": "static gang of thread blocks, one pair",
                                          "stage_ms": {k: st1[k] for k in ("downsample_ms", "knn_grid_ms", "sor_ms", "normals_ms", "icp_grid_ms", "icp_ms")},
                                          "iterations_per_scale": it1.tolist(),
                                          "ms_per_iteration_per_scale": [st1["scale_ms"][s] / (int(it1[s]) + 1) for s in range(S)],
                                          "pose_vs_batch_result": {"rot_rad": d_rot, "trans_m": d_tr,
                                                                   "note": "another block partition = another summation order (L1 chaos envelope)"}}

        # ---- BASELINE configs[2]: 1,000 fixed consecutive pairs partitioned over the ranks (strong scaling) ----------
        def run_fixed(tag, pair_list, T_init_all, batch_pairs):
            """pair_list: global list of (src_scan, tgt_scan) indices, identical on every rank; rank r takes a contiguous block
            (shard.partition), streams it in batches of <= batch_pairs through BatchStream from host memory, and the poses of all
            ranks are gathered with one collective at the end."""
            lo, hi = m.shard.partition(len(pair_list), rank, world)
            mine, T_mine = pair_list[lo:hi], T_init_all[lo:hi]
            cache.need({c for p in mine for c in p})
            if 0 < len(mine) <= batch_pairs and a.split > 1:
                batch_pairs = max(32, -(-len(mine) // a.split))
            batches = []
            for b0 in range(0, len(mine), batch_pairs):
                blk = mine[b0:b0 + batch_pairs]
                ids = sorted({c for p in blk for c in p})
                remap = {c: i for i, c in enumerate(ids)}
                batches.append((cache.get(ids), [(remap[s_], remap[t_]) for s_, t_ in blk], T_mine[b0:b0 + batch_pairs]))
            fbs = m.BatchStream(VOXELS, DISTS, MAX_IT, engines=engs, opts=opts)
            if batches:
                for _ in fbs.run((batches * 2)[:2]):       # warm both staging slots (pinned allocations) and the workspaces for this batch shape
                    pass
            barrier()
            t0_ = time.perf_counter()
            got = list(fbs.run(batches))
            if got:
                local = np.concatenate([np.concatenate([g.transformation.reshape(-1, 16), g.fitness[:, None], g.inlier_rmse[:, None]], axis=1) for g in got])
            else:
                local = np.zeros((0, 18))
            allres = m.shard.gather_results(torch.from_numpy(local).to(dev), len(pair_list), rank, world)
            barrier()
            dt_ = max_over_ranks(time.perf_counter() - t0_)
            fbs.close()
            its = np.concatenate([g.iterations for g in got]) if got else np.zeros((0, S))
            return {"pairs": len(pair_list), "pairs_per_s": len(pair_list) / dt_, "seconds": dt_, "n_gpus": world, "scaling": "strong",
                    "pairs_on_rank0": hi - lo, "batches_on_rank0": len(batches), "from": "pageable host clouds (BatchStream), poses gathered",
                    "mean_fitness": float(allres[:, 16].mean().item()), "iterations_mean_per_scale_rank0": its.mean(axis=0).tolist() if len(its) else None,
                    "iteration_cap_hit_frac_rank0": float((its >= MAX_IT).mean()) if len(its) else None}, allres

        n3 = a.config3_pairs
        if n3 > 0:
            t_g = time.perf_counter()
            pl3 = [(i + 1, i) for i in range(n3)]
            T3, truth3 = consecutive_inits(0, n3)
            r3, all3 = run_fixed("config3", pl3, T3, a.pairs)
            e3 = np.array([m.synthetic.pose_error(all3[b, :16].reshape(4, 4).cpu().numpy(), truth3[b])[1] for b in range(0, n3, max(1, n3 // 50))])
            r3["median_trans_err_vs_truth_m"] = float(np.median(e3))
            r3["workload"] = f"BASELINE configs[2]: {n3} consecutive pairs over {n3 + 1} scans, contiguous blocks per rank"
            r3["setup_s"] = time.perf_counter() - t_g - r3["seconds"]
            detail_extra["config3"] = r3
        # ---- BASELINE configs[4]: 10,000 non-consecutive pairs (|i-j| >= 50) with large initial offsets ------------------
        n5 = a.config5_pairs
        if n5 > 0 and n3 >= 1000:
            t_g = time.perf_counter()
            blocks, per = 8, n5 // 8
            pl5, T5 = [], []
            for b_ in range(blocks):
                rng = np.random.default_rng(5000 + b_)
                base = b_ * 125
                while len(pl5) < (b_ + 1) * per:
                    i_, j_ = rng.integers(0, 125, 2)
                    if abs(int(i_) - int(j_)) < 50:
                        continue
                    T_true = np.linalg.inv(m.synthetic.sensor_pose(base + int(j_))) @ m.synthetic.sensor_pose(base + int(i_))
                    pl5.append((base + int(i_), base + int(j_)))
                    T5.append(m.synthetic.perturbation(rng, rot_deg=10.0, trans=2.0) @ T_true)
            r5, _ = run_fixed("config5", pl5, np.stack(T5), per)
            r5["workload"] = (f"BASELINE configs[4]: {len(pl5)} non-consecutive pairs (|i-j| >= 50 scans) in {blocks} blocks of 125 scans of the "
                              "config-3 sequence, initial pose off by ~2 m / ~10 deg: low fitness, iteration caps")
            r5["setup_s"] = time.perf_counter() - t_g - r5["seconds"]
            detail_extra["config5"] = r5

    # ---- SURVEY 8(f) N3, the stage before: registro_FGR (hybrid normals + FPFH + feature matching on the tensor cores + graduated
    # non-convexity) on the real NCLT clouds of tests/golden/nclt_seq.npz, batched over consecutive pairs; CPU: the FGR oracle
    if rank == 0 and world == 1 and not a.no_extras and os.path.exists(os.path.join(ROOT, "tests", "golden", "nclt_seq.npz")):
        z = np.load(os.path.join(ROOT, "tests", "golden", "nclt_seq.npz"))
        zo = z["off"]
        ncl = [z["xyz"][zo[i]:zo[i + 1]] for i in range(len(zo) - 1)]
        fpairs = [tuple(p_) for p_ in z["pairs"].tolist()]
        caps = [int(int((len(ncl[s_]) + len(ncl[t_])) / 2) * 0.2) for s_, t_ in fpairs]        # AF:196
        fkw = dict(division_factor=1.4, use_absolute_scale=True, decrease_mu=True, maximum_correspondence_distance=0.2,
                   iteration_number=300, tuple_scale=0.95, maximum_tuple_count=caps, seeds=[m.pose_graph.pair_seed(0, s_, t_) for s_, t_ in fpairs])
        def fgr_once():
            ea_, eb_, ec_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ea_.record()
            feats_ = eng.fpfh_clouds(ncl, 0.2, 20, 1.0, 200, resident=True)        # descriptors stay in HBM, as registro_FGR keeps them
            eb_.record()
            Tf_, ncf_ = eng.fgr_pairs(ncl, feats_, fpairs, **fkw)
            ec_.record(); ec_.synchronize()
            return Tf_, ncf_, ea_.elapsed_time(eb_), eb_.elapsed_time(ec_)
        fgr_once()
        t_f = time.perf_counter()
        Tf, ncf, ms_feat, ms_reg = fgr_once()
        t_f = time.perf_counter() - t_f
        e_init = np.array([m.synthetic.pose_error(Tf[b_], z["T_golden"][b_]) for b_ in range(len(fpairs))])
        e_ship = np.array([m.synthetic.pose_error(z["T_fgr"][b_], z["T_golden"][b_]) for b_ in range(len(fpairs))])
        orc = _oracle()
        n_cpu = 2
        t_c = time.perf_counter()
        for b_ in range(n_cpu):
            orc.registro_FGR(ncl[fpairs[b_][0]].astype(np.float64), ncl[fpairs[b_][1]].astype(np.float64), 0.1, seed=0)
        t_c = time.perf_counter() - t_c
        # the reference's whole pairwise pipeline (Coarse_to_fine_FGR_M_GICP, AF:315-332, as full_registration calls it per pair),
        # batched: FGR from scratch, 3-scale M-GICP of the ALL_FUNCTIONS schedule from the FGR pose, information matrix
        m.pose_graph.register_pairs(ncl, fpairs[:4], 0.1, engine=eng)
        t_p = time.perf_counter()
        T_c2f, _info, fit_c2f, _rm = m.pose_graph.register_pairs(ncl, fpairs, 0.1, engine=eng)
        t_p = time.perf_counter() - t_p
        e_c2f = np.array([m.synthetic.pose_error(T_c2f[b_], z["T_golden"][b_]) for b_ in range(len(fpairs))])
        detail_extra["coarse_to_fine"] = {
            "workload": f"Coarse_to_fine_FGR_M_GICP (AF:315-332) batched over the same {len(fpairs)} NCLT pairs: registro_FGR, Multiscale_GICP "
                        "(ALL_FUNCTIONS schedule: voxels 0.4/0.2/0.1, search distances from the bounding boxes, L1, 100 iterations per scale) "
                        "from the FGR pose, get_information_matrix_from_point_clouds; host buffers in / out",
            "pairs_per_s": len(fpairs) / t_p, "seconds": t_p, "mean_fitness": float(np.mean(fit_c2f)),
            "median_trans_vs_shipped_refined_pose_m": float(np.median(e_c2f[:, 1])),
            "median_rot_vs_shipped_refined_pose_rad": float(np.median(e_c2f[:, 0])),
            "note": "the shipped refined poses come from the script-2 schedule (5 scales, 0.1 .. 0.5 m): same basin, not the same optimum"}
        detail_extra["fgr_front_end"] = {
            "workload": f"registro_FGR (AF:178-203) on {len(fpairs)} consecutive real NCLT pairs ({len(ncl)} clouds, ~{int(np.mean([len(c_) for c_ in ncl]))} pts): "
                        "hybrid normals + FPFH once per cloud, matching (tcgen05) + tuple test + 300 GNC iterations per pair, host buffers in / out",
            "pairs_per_s": len(fpairs) / t_f, "seconds": t_f, "ms_features_device": ms_feat, "ms_registration_device": ms_reg,
            "median_trans_vs_refined_pose_m": float(np.median(e_init[:, 1])), "shipped_fgr_median_trans_vs_refined_pose_m": float(np.median(e_ship[:, 1])),
            "cpu_oracle_pairs_per_s": n_cpu / t_c, "cpu_cores": orc.num_threads(), "cpu_sample": f"{n_cpu} pairs, {t_c:.1f} s"}

    if rank == 0:
        value = world * B * a.steps / (ms * 1e-3)
        e2e = world * B * a.steps / e2e_s
        # algorithmic bytes of the ICP loop: per scale (I+1)*24*M'_src + 72*sum_{i<I} K_i
        bytes_icp = float(np.sum(st[:, :, 7] * 24.0 * st[:, :, 0] + 72.0 * (st[:, :, 6] - st[:, :, 3])))
        n_raw = float(sum(len(s) for s in scans))
        icp_avg = sum(icp) / len(icp)
        peak, peak_src = peak_hbm()
        ach = bytes_icp / (icp_avg * 1e-3) / 1e9
        serial = detail_extra.get("stage_ms_per_step", {}).get("serialised_sum_ms")
        line = {"metric": "3-scale GICP scan-pairs/sec (~100k pts)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                        "api": "mgicp_b200.BatchStream.run: pageable numpy clouds -> pinned staging -> H2D -> preprocess + ICP -> D2H, per step"},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "k_icp_tasks (fused correspondence search + GICP linearisation + 6x6 solve loop, all pairs and "
                                       "scales in one launch)", "bound": "hbm",
                             "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                             "traffic": ncu_traffic(a.pairs), "algorithmic_bytes_per_launch": bytes_icp, "kernel_ms": icp_avg,
                             "share_of_step": (icp_avg / serial) if serial else None,
                             "note": "latency-bound chain of ~145 dependent passes per pair (gathers + fp64), see DESIGN.md; traffic = "
                                     "dram bytes read + written per launch from the committed ncu capture of this workload; share_of_step = "
                                     "kernel_ms / serialised sum of the step's stages (the other stages: detail.stage_ms_per_step)"},
                "clocks": clocks,
                "detail": {"engines": len(engs), "ms_icp_per_step": icp_avg,
                           "ms_preprocess_per_step": (ms / a.steps - icp_avg) if len(engs) == 1 else (serial - icp_avg if serial else None),
                           "ms_per_pair_icp_block": icp_avg, "iterations_mean_per_scale": st[:, :, 2].mean(axis=0).tolist(),
                           "points_after_sor_mean_per_scale": st[:, :, 0].mean(axis=0).tolist(),
                           "passes_per_pair_mean": float(st[:, :, 7].sum(axis=1).mean()),
                           "ms_per_iteration": icp_avg / float(st[:, :, 7].sum(axis=1).mean()),   # ICP kernel time / passes of a pair (one pair: latency of an iteration; a batch: amortised over the pairs in flight)
                           "pair_iterations_per_s": float(st[:, :, 7].sum()) / (icp_avg * 1e-3),
                           "median_trans_err_vs_truth_m": float(np.median([e[1] for e in err])),
                           "median_rot_err_vs_truth_rad": float(np.median([e[0] for e in err])),
                           "mean_fitness": float(fit.mean()), "raw_points_per_step": n_raw, "workload_gen_s": t_gen,
                           "ctas_per_pair": a.ctas_per_pair, "cell_factor": a.cell_factor, "icp_cell_factor": a.icp_cell_factor}}
        line["detail"].update(detail_extra)
        if serial:
            # the preprocessing half of the step against the same roofline: B_pre = sum over clouds and scales of b_in*N + 48*M + 80*M'
            # (M is not reported by the kernels; M' <= M, so this is a lower bound of the algorithmic bytes)
            b_pre = float(S * 12.0 * n_raw + 128.0 * sum(st[:, s, 0].sum() + st[-1:, s, 1].sum() for s in range(S)))
            pre_ms = serial - detail_extra["stage_ms_per_step"]["icp_ms"]
            line["detail"]["preprocess_roofline"] = {"algorithmic_bytes_per_step_lower_bound": b_pre, "ms": pre_ms,
                                                     "achieved_GBps": b_pre / (pre_ms * 1e-3) / 1e9, "frac": b_pre / (pre_ms * 1e-3) / 1e9 / peak,
                                                     "note": "issue- and latency-bound kNN / hash work, not bandwidth: see profiles/"}
        if world == 1 and not a.no_cpu_baseline:
            n_s = min(a.cpu_sample, B)
            pps, dt, cores, ref, split = oracle_run(scans, pairs, inits, n_s)
            line["cpu_baseline"] = {"value": pps, "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n_s} pairs of the same workload, {dt:.1f} s, CPU oracle (C/OpenMP restatement of the "
                                              "reference's Open3D path; Open3D is not installable offline)"}
            line["detail"]["cpu_stage_split"] = split
            n_env = min(n_s, 32)
            _, _, _, ref2, _ = oracle_run(scans, pairs, inits, n_env, sum_chunk=333)
            line["detail"]["parity"] = parity_block(m, T, fit, rm, ref, ref2)
        emit(line, out_fd)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
