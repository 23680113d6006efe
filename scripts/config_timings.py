"""Timings of BASELINE.json configs 4 and 5 on one GPU (reported in DESIGN.md; the bench line is config 1/2/3 shaped)."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import mgicp_b200 as m

eng = m.Engine(0)
out = {}
# config 4: ~2M-point TLS-like pair, script-2 4-scale schedule, L1
src, tgt, T0, Tt = m.synthetic.make_tls_pair(2_000_000, seed=2)
src, tgt = src.astype(np.float32), tgt.astype(np.float32)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = m.Multiscale_GICP(src, tgt, 4, 100, T0, engine=eng)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
rot, tr = m.synthetic.pose_error(r.transformation, Tt)
out["config4_2M_pair"] = {"seconds_end_to_end_host_buffers": dt, "iterations": r.iterations, "points_per_scale": r.stats[:, 0].astype(int).tolist(),
                          "fitness": r.fitness, "rmse": r.inlier_rmse, "err_m": tr, "err_rad": rot}
t = time.perf_counter(); ev = m.evaluate_registration(src, tgt, 0.1, r.transformation, engine=eng); torch.cuda.synchronize()
out["config4_evaluate_registration_2Mx2M"] = {"seconds": time.perf_counter() - t, "fitness": ev.fitness, "rmse": ev.inlier_rmse}
# config 5: loop-closure sweep, non-consecutive pairs with identity initial guesses
n = 120
scans, inits, truths = m.synthetic.make_sequence(n, azimuth_steps=3125, seed=8)
scans = [s.astype(np.float32) for s in scans]
pairs = [(i, j) for i in range(n) for j in range(n) if i - j >= 50][:2000]
T0 = np.stack([np.eye(4)] * len(pairs))
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = m.multiscale_gicp_batch(scans, pairs, [1.0, 0.5, 0.25], [3.0, 1.0, 0.25], 100, T0, engine=eng)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
out["config5_loop_closure"] = {"pairs": len(pairs), "clouds": n, "seconds": dt, "pairs_per_s": len(pairs) / dt,
                               "mean_iterations_per_scale": r.iterations.mean(0).tolist(), "mean_fitness": float(r.fitness.mean()),
                               "frac_fitness_gt_0.4": float((r.fitness > 0.4).mean())}
print(json.dumps(out, indent=1))
