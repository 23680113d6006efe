"""Parity of the FGR front end's feature stage (hybrid-radius normals + FPFH, csrc/mgicp_fgr.cuh) against the oracle.

Normals are bit-exact (same neighbour lists in the same order, same closed-form eigenvector); descriptors are compared at 1e-9
because atan2 comes from libdevice on the GPU and libm in the oracle (a neighbour sitting on a bin boundary may move) -- on the
first B200 run they were bit-identical too (max |diff| 0.0 on the NCLT fixtures)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nclt")


def _compare(oracle, cloud, nrm, fp, rn, kn, rf, kf):
    ref_n = oracle.estimate_normals_hybrid(cloud, rn, kn)
    ref_f = oracle.compute_fpfh_feature(cloud, ref_n, rf, kf)
    print(f"n={len(cloud)} normals: equal rows {(nrm == ref_n).all(axis=1).mean():.4f}, max |diff| {np.abs(nrm - ref_n).max():.3e}; "
          f"fpfh: rows within 1e-9 {(np.abs(fp - ref_f).max(axis=1) < 1e-9).mean():.4f}, max |diff| {np.abs(fp - ref_f).max():.3e}, "
          f"finite {np.isfinite(fp).all()}")
    assert np.array_equal(nrm, ref_n)                                    # same lists, same order, same closed-form eigenvector
    # atan2 comes from libdevice on the GPU and libm in the oracle: a neighbour sitting on a bin boundary may move
    close = np.abs(fp - ref_f).max(axis=1) < 1e-9
    assert close.mean() > 0.995, close.mean()
    thirds = fp.reshape(-1, 3, 11).sum(axis=2)
    has = thirds.sum(axis=1) > 0
    assert np.allclose(thirds[has], 200.0, atol=1e-9)


def test_fpfh_on_nclt_fixture(pkg, oracle, engine):
    """the reference's parameters (voxel 0.1: radii 0.2 / 1.0, max_nn 20 / 200) on real pre-processed NCLT clouds"""
    clouds = [pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, f"s{i}.pcd")) for i in (0, 1)]
    nrm, fp = engine.fpfh_clouds(clouds, 0.2, 20, 1.0, 200)
    for c, n, f in zip(clouds, nrm, fp):
        _compare(oracle, c.astype(np.float64), n, f, 0.2, 20, 1.0, 200)


def test_fpfh_caps_and_edges(pkg, oracle, engine):
    """max_nn binding on most points, an isolated point, a tiny cloud, float64 input"""
    src, _, _, _ = pkg.synthetic.make_pair(300, seed=4)
    ds = np.asarray(oracle.voxel_down_sample(src, 0.3))
    lone = np.vstack([ds, [[500.0, 500.0, 50.0]]])
    nrm, fp = engine.fpfh_clouds([lone, ds[:2]], 1.0, 8, 4.0, 16)
    _compare(oracle, lone, nrm[0], fp[0], 1.0, 8, 4.0, 16)
    assert not fp[0][-1].any() and np.array_equal(nrm[0][-1], [0.0, 0.0, 1.0])
    _compare(oracle, ds[:2], nrm[1], fp[1], 1.0, 8, 4.0, 16)
    with pytest.raises(Exception):
        engine.fpfh_clouds([ds], -1.0, 8, 4.0, 16)


# ---- registration stage (matching, tuple test, graduated non-convexity) ----
# First B200 run (round 2, profiles/r2_first/r2_fgr_first.log): correspondences identical to the oracle, poses bit-identical to
# the oracle run in the kernel's reduction order and within 5e-15 of the sequentially summed oracle.
REF_OPTS = dict(division_factor=1.4, use_absolute_scale=True, decrease_mu=True, maximum_correspondence_distance=0.2,
                iteration_number=300, tuple_scale=0.95)


def test_fgr_pairs_on_nclt_fixture(pkg, oracle, engine):
    """descriptors from the oracle, the reference's option values, both orders of the pair (the larger cloud becomes `i`):
    same correspondences as the oracle (discrete decisions are exact), pose within 1e-8 (the 27 sums are reduced in a
    different order)"""
    s = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, "s18.pcd")).astype(np.float64)
    t = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, "s17.pcd")).astype(np.float64)
    fs = oracle.compute_fpfh_feature(s, oracle.estimate_normals_hybrid(s, 0.2, 20), 1.0, 200)
    ft = oracle.compute_fpfh_feature(t, oracle.estimate_normals_hybrid(t, 0.2, 20), 1.0, 200)
    cap = int(int((len(s) + len(t)) / 2) * 0.2)
    T, nc = engine.fgr_pairs([s, t], [fs, ft], [(0, 1), (1, 0)], maximum_tuple_count=cap, seeds=[7, 9], **REF_OPTS)
    for b, (a, c, fa, fc, seed) in enumerate(((s, t, fs, ft, 7), (t, s, ft, fs, 9))):
        ref, nref = oracle.registration_fgr_based_on_feature_matching(a, c, fa, fc, maximum_tuple_count=cap, seed=seed, **REF_OPTS)
        # ... and with the 27 sums of every iteration reduced in the kernel's order: should be the GPU result bit for bit
        ref_k, _ = oracle.registration_fgr_based_on_feature_matching(a, c, fa, fc, maximum_tuple_count=cap, seed=seed,
                                                                     engine="kernel_order", **REF_OPTS)
        print(f"pair {b}: correspondences {nc[b]} / {nref}, max |dT| vs oracle {np.abs(T[b] - ref).max():.3e}, "
              f"vs kernel-order oracle {np.abs(T[b] - ref_k).max():.3e}")
        assert nc[b] == nref
        assert np.abs(T[b] - ref).max() < 1e-8
        assert np.abs(T[b] - ref_k).max() < 1e-12
    assert np.abs(T[0] @ T[1] - np.eye(4)).max() < 0.05          # the two directions are (roughly) inverse to each other


def test_fgr_pairs_recovers_an_exact_motion(pkg, oracle, engine):
    src, _, _, _ = pkg.synthetic.make_pair(600, seed=11)
    scene = np.asarray(oracle.voxel_down_sample(src, 0.5))
    rng = np.random.default_rng(5)
    a = rng.normal(size=3); a /= np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    T = np.eye(4); T[:3, :3] = np.eye(3) + np.sin(0.4) * K + (1 - np.cos(0.4)) * (K @ K); T[:3, 3] = 3.0 * rng.normal(size=3)
    fs = oracle.compute_fpfh_feature(scene, oracle.estimate_normals_hybrid(scene, 1.0, 20), 5.0, 200)
    perm = rng.permutation(len(scene))[: len(scene) - 37]
    tgt, ft = (scene @ T[:3, :3].T + T[:3, 3])[perm], fs[perm]
    for absolute in (True, False):
        got, nc = engine.fgr_pairs([scene, tgt], [fs, ft], [(0, 1), (1, 0)], use_absolute_scale=absolute, decrease_mu=True,
                                   maximum_correspondence_distance=1.0 if absolute else 0.01, iteration_number=300,
                                   maximum_tuple_count=2000, seeds=[3, 3])
        assert nc.tolist() == [6000, 6000]
        assert np.abs(got[0] - T).max() < 1e-6 and np.abs(got[1] - np.linalg.inv(T)).max() < 1e-6


def test_registro_FGR_end_to_end(pkg, engine):
    """the reference's call on real NCLT clouds: as good a coarse alignment as the shipped FGR pose"""
    for a, b in ((1, 0), (18, 17)):
        src = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, f"s{a}.pcd"))
        tgt = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, f"s{b}.pcd"))
        T_fgr = np.loadtxt(os.path.join(GOLD, f"fgr_pose_{a}_{b}.txt"))
        T_ref = np.loadtxt(os.path.join(GOLD, f"golden_pose_{a}_{b}.txt"))
        r = pkg.registro_FGR(src, tgt, 0.1, engine=engine)
        rot, tr = pkg.synthetic.pose_error(r.transformation, T_ref)
        rot_s, tr_s = pkg.synthetic.pose_error(T_fgr, T_ref)
        print(f"pair {a}->{b}: {tr:.3f} m / {rot:.4f} rad from the refined pose (shipped FGR {tr_s:.3f} m / {rot_s:.4f} rad), fitness {r.fitness:.3f}")
        assert tr < 0.2 and rot < 0.03 and 0.0 < r.fitness <= 1.0



def test_Coarse_to_fine_FGR_M_GICP(pkg, oracle, engine):
    """ALL_FUNCTIONS.py:315-332 on real NCLT clouds: FGR from scratch, then the 3-scale refinement of the ALL_FUNCTIONS schedule
    (voxels 0.4/0.2/0.1, radius-derived search distances) and the information matrix at the voxel size.  The refined pose must
    land on the shipped refined pose (which came from the script-2 schedule, 5 scales: both converge to the same basin), and the
    information matrix must equal the one the raw-cloud evaluation gives at that pose."""
    for (a, b), loss in (((1, 0), "l1"), ((18, 17), "l2")):      # L1 = the reference's kernel (AF:284), with its ~40 m search radius
        src = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, f"s{a}.pcd"))
        tgt = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, f"s{b}.pcd"))
        T_ref = np.loadtxt(os.path.join(GOLD, f"golden_pose_{a}_{b}.txt"))
        res, info = pkg.Coarse_to_fine_FGR_M_GICP(src, tgt, 0.1, engine=engine, loss=loss)
        rot, tr = pkg.synthetic.pose_error(res.transformation, T_ref)
        print(f"Coarse_to_fine {a}->{b} ({loss}): {tr:.4f} m / {rot:.5f} rad from the shipped refined pose, fitness {res.fitness:.3f}, "
              f"rmse {res.inlier_rmse:.4f}, info[5,5] = {info[5, 5]:.0f}")
        # with the ALL_FUNCTIONS schedule every point finds a partner (search radius ~40 m, fitness 1): the reference's own, slightly
        # biased optimum; first B200 run: 15 mm / 3.6e-3 rad (L1) from the script-2 golden pose
        assert tr < 0.05 and rot < 6e-3 and res.fitness > 0.5
        # the refinement half against the oracle's Multiscale_GICP (same schedule, same kernel) from the same FGR pose
        T_fgr = pkg.registro_FGR(src, tgt, 0.1, engine=engine).transformation
        got = pkg.Multiscale_GICP(src, tgt, 3, 100, T_fgr, schedule="all_functions", loss=loss, engine=engine)
        ref = oracle.Multiscale_GICP(src.astype(np.float64), tgt.astype(np.float64), 3, 100, T_fgr, schedule="all_functions", loss=loss)
        rot_o, tr_o = pkg.synthetic.pose_error(got.transformation, ref.transformation)
        print(f"   Multiscale_GICP(all_functions, {loss}) vs oracle: {tr_o:.2e} m / {rot_o:.2e} rad, iterations {got.iterations} / {ref.iterations}, "
              f"fitness {got.fitness:.4f} / {ref.fitness:.4f}, rmse {got.inlier_rmse:.5f} / {ref.inlier_rmse:.5f}")
        assert np.array_equal(got.transformation, res.transformation)          # Coarse_to_fine is exactly these two steps
        if loss == "l2":
            assert tr_o < 1e-6 and rot_o < 1e-7 and got.iterations == ref.iterations      # first B200 run: 3.5e-8 m / 1.3e-9 rad
        else:
            assert tr_o < 5e-3 and rot_o < 1e-3 and abs(got.fitness - ref.fitness) < 1e-3
        ref_info = oracle.get_information_matrix_from_point_clouds(src.astype(np.float64), tgt.astype(np.float64), 0.1, res.transformation)
        assert info.shape == (6, 6) and info[5, 5] == info[4, 4] == info[3, 3] > 100
        assert np.allclose(info, ref_info, rtol=1e-11, atol=1e-8)


def test_resident_descriptors_equal_the_host_round_trip(pkg, engine):
    """fpfh_clouds(..., resident=True) keeps clouds, normals and descriptors in device memory for fgr_pairs: same descriptors,
    same poses, bit for bit, as handing the host arrays back in"""
    clouds = [pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, f"s{i}.pcd")) for i in (0, 1, 17, 18)]
    pairs = [(1, 0), (3, 2), (0, 1)]
    caps = [int(int((len(clouds[s]) + len(clouds[t])) / 2) * 0.2) for s, t in pairs]
    nrm, fp = engine.fpfh_clouds(clouds, 0.2, 20, 1.0, 200)
    res = engine.fpfh_clouds(clouds, 0.2, 20, 1.0, 200, resident=True)
    nrm_r, fp_r = res.host()
    for a, b in zip(nrm + fp, nrm_r + fp_r):
        assert np.array_equal(a, b)
    T_h, nc_h = engine.fgr_pairs(clouds, fp, pairs, maximum_tuple_count=caps, seeds=[1, 2, 3], **REF_OPTS)
    T_r, nc_r = engine.fgr_pairs(None, res, pairs, maximum_tuple_count=caps, seeds=[1, 2, 3], **REF_OPTS)
    assert np.array_equal(T_h, T_r) and np.array_equal(nc_h, nc_r)


@pytest.mark.parametrize("popular", [0.12, 0.7])
def test_tensor_core_matching_equals_brute_force_on_degenerate_descriptors(pkg, engine, popular):
    """The matcher's hard cases: descriptors repeated exactly within and across the clouds (distance 0, the lowest index must win),
    clusters of near-duplicates (1e-9 apart: nothing the split-fp16 product can separate), all-zero descriptors, and ordinary ones.
    With 12 % of the rows drawn from a small pool of repeated descriptors the unsettled rows go through the queued fp64 search
    (k_fgr_match_fb, < 1024 rows per direction); with 70 % the queue overflows and the direction falls to the plain brute force.
    The tensor-core path (MGICP_FGR_MATCH=1, default) and the fp64 brute force (MGICP_FGR_MATCH=0) must give the same nearest
    neighbours, i.e. bit-identical poses and correspondence counts."""
    rng = np.random.default_rng(11)
    n_a, n_b = 3100, 2900

    def fpfh_like(m):
        f = rng.gamma(0.6, 20.0, size=(m, 33))
        return f * (200.0 / f.reshape(m, 3, 11).sum(axis=2).repeat(11, axis=1))      # thirds sum to 200, like FPFH

    pool = fpfh_like(40)
    pool[:3] = 0.0                                                                  # isolated points: all-zero descriptors
    fa, fb = fpfh_like(n_a), fpfh_like(n_b)
    # shared structure so that the clouds do match: most rows of b are noisy copies of rows of a
    src = rng.integers(0, n_a, n_b)
    fb = fa[src] * (1.0 + 0.02 * rng.standard_normal((n_b, 33)))
    ia, ib = rng.random(n_a) < popular, rng.random(n_b) < popular
    fa[ia] = pool[rng.integers(0, 40, int(ia.sum()))]
    fb[ib] = pool[rng.integers(0, 40, int(ib.sum()))]
    near = ib & (rng.random(n_b) < 0.5)
    fb[near] *= 1.0 + 1e-9 * rng.standard_normal((int(near.sum()), 33))              # near-duplicates
    pa = rng.uniform(-20, 20, (n_a, 3))
    R = pkg.synthetic.perturbation(rng, rot_deg=20.0, trans=3.0)
    pb = pa[src] @ R[:3, :3].T + R[:3, 3]
    kw = dict(maximum_tuple_count=600, seeds=[3, 4], **REF_OPTS)
    old = os.environ.get("MGICP_FGR_MATCH")
    try:
        os.environ["MGICP_FGR_MATCH"] = "1"
        T1, nc1 = engine.fgr_pairs([pa, pb], [fa, fb], [(0, 1), (1, 0)], **kw)
        os.environ["MGICP_FGR_MATCH"] = "0"
        T0, nc0 = engine.fgr_pairs([pa, pb], [fa, fb], [(0, 1), (1, 0)], **kw)
    finally:
        if old is None:
            os.environ.pop("MGICP_FGR_MATCH", None)
        else:
            os.environ["MGICP_FGR_MATCH"] = old
    print(f"popular {popular}: correspondences {nc1} / {nc0}")
    assert np.array_equal(nc1, nc0) and np.array_equal(T1, T0)
    assert nc1.min() > 100
