"""Real data: 64 consecutive NCLT pairs (tests/golden/nclt_seq.npz, built by tests/golden/make_nclt_sequence.py from the
reference's shipped clouds, FGR poses and refined poses), refined with the reference's script-2 call
Multiscale_GICP(source, target, 5, 100, T_fgr) -- L1 kernel, 5 scales (2_MGICP_refinement_in_NCLT_dataset.py:187-218).

Checked: the GPU engine against (a) the CPU oracle's poses for the same pairs (computed in the build container by
oracle/pin_against_goldens.py --all, stored in the fixture) and (b) the reference's own shipped refined poses, where the
yardstick is the oracle's own pin (median 0.5 mm, 90 % within 6 mm / 5e-4 rad over all 791 pairs)."""
import os

import numpy as np
import pytest

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nclt_seq.npz")


def _load():
    z = np.load(FIX)
    off = z["off"]
    clouds = [z["xyz"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    return z, clouds


def test_fixture_reproduces_the_oracle_pin(pkg, oracle):
    """CPU: the oracle, run here on the fixture's clouds, lands on the poses stored in the fixture (same code, same machine
    class: bit-identical or at rounding level -- OpenMP chunking is deterministic)"""
    z, clouds = _load()
    assert clouds[0].dtype == np.float32 and len(z["pairs"]) == 64
    for b in (0, 33):
        s, t = z["pairs"][b]
        r = oracle.Multiscale_GICP(clouds[s].astype(np.float64), clouds[t].astype(np.float64), 5, 100, z["T_fgr"][b], schedule="script2")
        rot, tr = pkg.synthetic.pose_error(r.transformation, z["T_oracle"][b])
        assert rot < 1e-9 and tr < 1e-9, (b, rot, tr)
        assert r.iterations == z["oracle_iters"][b].tolist()


@pytest.mark.gpu
def test_nclt_sequence_gpu_vs_oracle_and_shipped_goldens(pkg, engine):
    z, clouds = _load()
    pairs = [tuple(p) for p in z["pairs"].tolist()]
    vox = pkg.create_scales_script2(5)
    dists = pkg.max_correspondence_distances(vox)
    r = pkg.multiscale_gicp_batch(clouds, pairs, vox, dists, 100, z["T_fgr"], engine=engine, loss="l1")
    B = len(pairs)
    vs_orc = np.array([pkg.synthetic.pose_error(r.transformation[b], z["T_oracle"][b]) for b in range(B)])
    vs_gold = np.array([pkg.synthetic.pose_error(r.transformation[b], z["T_golden"][b]) for b in range(B)])
    orc_gold = np.array([pkg.synthetic.pose_error(z["T_oracle"][b], z["T_golden"][b]) for b in range(B)])
    init_gold = np.array([pkg.synthetic.pose_error(z["T_fgr"][b], z["T_golden"][b]) for b in range(B)])
    dfit, drm = np.abs(r.fitness - z["oracle_fitness"]), np.abs(r.inlier_rmse - z["oracle_rmse"])
    q = lambda x: f"median {np.median(x):.2e} p90 {np.quantile(x, 0.9):.2e} max {x.max():.2e}"
    within = lambda e: float(np.mean((e[:, 1] <= 6e-3) & (e[:, 0] <= 5e-4)))
    print(f"\n{B} real NCLT pairs, script-2 schedule (5 scales, L1, 100 it), T_init = shipped FGR poses (%.10f text)")
    print(f"  FGR init  vs shipped golden: trans [{q(init_gold[:, 1])}] m")
    print(f"  oracle    vs shipped golden: trans [{q(orc_gold[:, 1])}] m, rot [{q(orc_gold[:, 0])}] rad, within 6 mm / 5e-4 rad: {100 * within(orc_gold):.0f} %")
    print(f"  GPU       vs shipped golden: trans [{q(vs_gold[:, 1])}] m, rot [{q(vs_gold[:, 0])}] rad, within 6 mm / 5e-4 rad: {100 * within(vs_gold):.0f} %")
    print(f"  GPU       vs oracle:         trans [{q(vs_orc[:, 1])}] m, rot [{q(vs_orc[:, 0])}] rad; dfitness [{q(dfit)}], drmse [{q(drm)}]; "
          f"inside 1e-4 m / 1e-4 rad: {100 * np.mean((vs_orc[:, 1] < 1e-4) & (vs_orc[:, 0] < 1e-4)):.0f} %")
    print(f"  iterations per scale (mean): GPU {r.iterations.mean(axis=0).round(1).tolist()}, oracle {z['oracle_iters'].mean(axis=0).round(1).tolist()}")
    # against the reference's own poses the GPU engine is pinned exactly as well as the oracle is
    assert np.median(vs_gold[:, 1]) <= 1.5 * np.median(orc_gold[:, 1]) + 1e-4
    assert within(vs_gold) >= within(orc_gold) - 0.05
    # against the oracle: real clouds sit on a 5 mm lattice (distance ties) and the L1 loop is chaotic: the bulk agrees to a
    # fraction of a millimetre, the tail is bounded by the oracle's own distance to the golden poses.  Two chaotic realisations of
    # the same loop lie about as far from each other as each lies from Open3D's: the bound follows the oracle's own median distance
    # to the shipped poses (0.17 mm).  Measured medians, GPU vs oracle: 0.16 mm with the kNN grid at 10 voxels, 0.24 mm at 12
    # (a different point order, i.e. another realisation; vs the shipped poses 0.15 / 0.20 mm, the oracle 0.17 mm)
    assert np.median(vs_orc[:, 1]) <= 2.0 * np.median(orc_gold[:, 1]) + 1e-4 and np.median(vs_orc[:, 0]) <= 2.0 * np.median(orc_gold[:, 0]) + 1e-5
    assert np.quantile(vs_orc[:, 1], 0.9) <= max(3e-3, 2.0 * np.quantile(orc_gold[:, 1], 0.9))
    assert np.median(dfit) < 1e-3 and np.median(drm) < 1e-4
