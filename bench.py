#!/usr/bin/env python
"""bench.py -- 3-scale GICP scan-pairs/sec on synthetic NCLT-shaped (~100k-point, HDL-32 pattern) pairs.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched under torch.distributed.run by the driver)
    python bench.py --impl reference ...                    (CPU arm: the oracle restatement of the Open3D path)

A step = one pass of the hot path (per-scale voxel down-sample, outlier removal, normals, 3-scale ICP loops) over one
batch of `--pairs` consecutive scan pairs per GPU (BASELINE.json configs[1] batched as in configs[2]); voxels
1.0/0.5/0.25 m, max correspondence distance 3x/2x/1x voxel, 100 iterations per scale, L1 kernel (the reference's
setting).  Whole pairs are sharded across ranks (weak scaling); only the resulting poses are gathered (NCCL).
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


def make_workload(n_pairs, rank, az=AZIMUTH, seed=0):
    """n_pairs+1 consecutive scans (float32, the PCD-native dtype) + FGR-like initial poses; rank-specific stretch of the circuit"""
    import multiprocessing as mp
    import mgicp_b200 as m
    k0 = rank * (n_pairs + 1)
    jobs = [(k0 + i, seed, az) for i in range(n_pairs + 1)]
    nproc = max(1, min(len(jobs), (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    with mp.get_context("fork").Pool(nproc) as pool:
        scans = pool.map(_gen_scan, jobs)
    inits, truths = [], []
    for i in range(n_pairs):
        T_true = np.linalg.inv(m.synthetic.sensor_pose(k0 + i)) @ m.synthetic.sensor_pose(k0 + i + 1)
        rng = np.random.default_rng(77 + 1000003 * seed + k0 + i)
        inits.append(m.synthetic.perturbation(rng) @ T_true)
        truths.append(T_true)
    pairs = [(i + 1, i) for i in range(n_pairs)]      # source = scan i+1, target = scan i (S2:191)
    return scans, pairs, np.stack(inits), np.stack(truths)


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


def oracle_pairs_per_sec(scans, pairs, inits, n_sample):
    """the CPU restatement of the reference's Open3D path, all host threads, on the first n_sample pairs of the workload"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1 to every rank: not what a CPU arm wants)
    try:
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        oracle.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    for (s, t), T0 in list(zip(pairs, inits))[:n_sample]:
        oracle.multiscale_gicp(scans[s].astype(np.float64), scans[t].astype(np.float64), VOXELS, DISTS, MAX_IT, T0, loss="l1")
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, oracle.num_threads()


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

    # ------------------------------------------------------------------ reference arm (CPU) ----------
    if a.impl == "reference":
        if rank != 0:
            return
        n_s = max(2, min(a.pairs, 32))        # ~5 s of CPU work per step
        scans, pairs, inits, _ = make_workload(n_s, 0, a.azimuth)
        for _ in range(min(a.warmup, 1)):
            oracle_pairs_per_sec(scans, pairs, inits, 1)
        times = []
        cores = 1
        for _ in range(a.steps):
            pps, dt, cores = oracle_pairs_per_sec(scans, pairs, inits, n_s)
            times.append(dt)
        tot = sum(times)
        val = n_s * a.steps / tot
        sample = f"{n_s} pairs of the same workload per step, {a.steps} steps, CPU oracle (C restatement of the Open3D path; Open3D itself is not installable offline)"
        emit({"impl": "reference", "metric": "3-scale GICP scan-pairs/sec (~100k pts)", "value": val, "unit": "pairs/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}, out_fd)
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
    scans, pairs, inits, truths = make_workload(a.pairs, rank, a.azimuth)
    t_gen = time.perf_counter() - t_gen
    flat, off, _ = eng.pack_clouds(scans)
    B, S = len(pairs), len(VOXELS)
    ps, pt = [p[0] for p in pairs], [p[1] for p in pairs]
    md = np.broadcast_to(np.asarray(DISTS), (B, S)).copy()
    mi = np.full(S, MAX_IT, np.int32)
    flat_pin = torch.from_numpy(flat).pin_memory()
    T0_pin = torch.from_numpy(inits.reshape(B, 16).copy()).pin_memory()
    xyz_dev = flat_pin.to(dev)
    T0_dev = T0_pin.to(dev)
    gathered = [torch.empty((world * B, 18), dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None

    # Consecutive batches alternate between two engines (two workspaces) on two streams: the preprocessing kernels of
    # batch k+1 fill the SMs that the persistent ICP kernel of batch k leaves idle towards its end (measured +3 %).
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

    # end to end: every step's inputs come from pinned HOST memory and its results go back to the host.  The upload of
    # step k+1 runs on a copy stream while step k computes (two device buffers), like a caller streaming batches would do.
    copy_stream = torch.cuda.Stream(device=dev)
    xyz_buf = [torch.empty_like(xyz_dev) for _ in range(2)]
    T0_buf = [torch.empty_like(T0_dev) for _ in range(2)]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_pre = torch.cuda.Event()

    def upload_e2e(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[b])          # the step that last used this buffer is done
            xyz_buf[b].copy_(flat_pin, non_blocking=True)
            T0_buf[b].copy_(T0_pin, non_blocking=True)
            ev_up[b].record(copy_stream)

    d2h_stream = torch.cuda.Stream(device=dev)
    res_pin = [torch.empty(((world if world > 1 else 1) * B, 18), dtype=torch.float64).pin_memory() for _ in range(2)]
    ev_step = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    dbg = [] if os.environ.get("MGICP_BENCH_DEBUG") else None

    def run_e2e(n_steps):
        """Streams n_steps batches: upload of step k+1 (copy stream) and download of step k-1 (another stream) overlap the
        kernels of step k; the host blocks only on the download of step k-1 AFTER it has enqueued step k, so the GPU never
        waits for the host.  Every step's inputs cross PCIe from pinned host memory and its results land in host memory."""
        for b in range(2):
            ev_free[b].record(torch.cuda.current_stream(dev))
        upload_e2e(0)
        results = []
        pending = None                      # (device result tensor, slot) of the previous step
        for k in range(n_steps):
            b = k & 1
            e_ = engs[k % len(engs)]
            cur = work[k % len(engs)]
            cur.wait_event(ev_up[b])
            if dbg is not None:
                dbg.append([torch.cuda.Event(enable_timing=True) for _ in range(2)] + [time.perf_counter()])
                dbg[-1][0].record(cur)
            torch.cuda.set_stream(cur)
            e_.preprocess_device(xyz_buf[b], off, VOXELS, opts)
            if k + 1 < n_steps:
                # the next batch crosses PCIe while the (latency-bound) ICP kernel runs, not during the bandwidth- and
                # atomics-heavy preprocessing kernels
                ev_pre.record(cur)
                copy_stream.wait_event(ev_pre)
                upload_e2e(1 - b)
            out = e_.register_device(ps, pt, md, mi, T0_buf[b], opts)
            if dbg is not None:
                dbg[-1][1].record(cur)
                dbg[-1].append(time.perf_counter())
            ev_free[b].record(cur)
            T, fit, rm = out[0], out[1], out[2]
            local = torch.cat([T.reshape(B, 16), fit[:, None], rm[:, None]], dim=1)
            if world > 1:
                dist.all_gather_into_tensor(gathered[b], local)
                local = gathered[b].clone()
            ev_step[b].record(cur)
            if pending is not None:         # fetch the previous step's results while this step runs
                results.append(fetch_e2e(*pending))
            pending = (local, b)
        results.append(fetch_e2e(*pending))
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        return results[-1]

    def fetch_e2e(dev_res, b):
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(ev_step[b])
            res_pin[b].copy_(dev_res, non_blocking=True)
            ev_out[b].record(d2h_stream)
        ev_out[b].synchronize()
        return res_pin[b].clone()

    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(a.warmup * len(engs)):
        out = step_device(k)
    torch.cuda.synchronize()
    for e_ in engs:
        e_.check()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        if True:   # per-launch duration of the dominant kernel (events on the launching stream; no host sync here)
            icp_ms.append((ev_a, ev_b))
            ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    join_work()
    e1.record()
    barrier()
    launches = sum(e_.kernel_launches() for e_ in engs) - launches0
    ms = e0.elapsed_time(e1)
    icp = [x.elapsed_time(y) for x, y in icp_ms]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # e2e: host buffers in (pinned), host results out, every step
    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    if dbg is not None:
        dbg.clear()
    res = run_e2e(a.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dbg is not None and rank == 0:
        for k, (ea, eb, ta, tb) in enumerate(dbg):
            gap = dbg[k - 1][1].elapsed_time(ea) if k else 0.0
            print(f"e2e step {k}: device {ea.elapsed_time(eb):.2f} ms, idle before {gap:.2f} ms, host enqueue {1e3 * (tb - ta):.2f} ms at t={1e3 * (ta - t0):.1f}",
                  file=sys.stderr)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    T, fit, rm, it, nc, st = (x.cpu().numpy() for x in out)
    for e_ in engs:
        e_.check()
    err = [m.synthetic.pose_error(T[b], truths[b]) for b in range(B)]
    if rank == 0:
        value = world * B * a.steps / (ms * 1e-3)
        e2e = world * B * a.steps / e2e_s
        # algorithmic bytes of the ICP loop: per scale (I+1)*24*M'_src + 72*sum_{i<I} K_i
        bytes_icp = float(np.sum(st[:, :, 7] * 24.0 * st[:, :, 0] + 72.0 * (st[:, :, 6] - st[:, :, 3])))
        n_raw = float(sum(len(s) for s in scans))
        icp_avg = sum(icp) / len(icp)
        peak, peak_src = peak_hbm()
        ach = bytes_icp / (icp_avg * 1e-3) / 1e9
        line = {"metric": "3-scale GICP scan-pairs/sec (~100k pts)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(flat.nbytes + inits.nbytes),
                        "d2h_bytes_per_step": int(res.numel() * 8)},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "k_icp_tasks (fused correspondence search + GICP linearisation + 6x6 solve loop, all pairs and "
                                       "scales in one launch)", "bound": "hbm",
                             "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                             "traffic": ncu_traffic(a.pairs), "algorithmic_bytes_per_launch": bytes_icp, "kernel_ms": icp_avg,
                             "note": "latency-bound chain of ~145 dependent passes per pair (gathers + fp64), see DESIGN.md; traffic = "
                                     "dram bytes read + written per launch from the committed ncu capture of this workload"},
                "clocks": clocks,
                "detail": {"engines": len(engs), "ms_icp_per_step": icp_avg,
                           "ms_preprocess_per_step": (ms / a.steps - icp_avg) if len(engs) == 1 else None,
                           "ms_per_pair_icp_block": icp_avg, "iterations_mean_per_scale": st[:, :, 2].mean(axis=0).tolist(),
                           "points_after_sor_mean_per_scale": st[:, :, 0].mean(axis=0).tolist(),
                           "passes_per_pair_mean": float(st[:, :, 7].sum(axis=1).mean()),
                           "ms_per_iteration": icp_avg / float(st[:, :, 7].sum(axis=1).mean()),   # ICP kernel time / passes of a pair (one pair: latency of an iteration; a batch: amortised over the pairs in flight)
                           "pair_iterations_per_s": float(st[:, :, 7].sum()) / (icp_avg * 1e-3),
                           "median_trans_err_vs_truth_m": float(np.median([e[1] for e in err])),
                           "median_rot_err_vs_truth_rad": float(np.median([e[0] for e in err])),
                           "mean_fitness": float(fit.mean()), "raw_points_per_step": n_raw, "workload_gen_s": t_gen,
                           "ctas_per_pair": a.ctas_per_pair, "cell_factor": a.cell_factor, "icp_cell_factor": a.icp_cell_factor}}
        if world == 1 and not a.no_cpu_baseline:
            n_s = min(a.cpu_sample, B)
            pps, dt, cores = oracle_pairs_per_sec(scans, pairs, inits, n_s)
            line["cpu_baseline"] = {"value": pps, "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n_s} pairs of the same workload, {dt:.1f} s, CPU oracle (C/OpenMP restatement of the "
                                              "reference's Open3D path; Open3D is not installable offline)"}
        emit(line, out_fd)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
