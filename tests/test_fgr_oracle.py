"""CPU tests of the FGR front-end oracle (oracle/fgr_oracle.c): unit properties of the hybrid-radius normals,
FPFH and the FGR optimisation, and the soft pin against the reference's shipped NCLT poses.  No GPU, no product code."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _rigid(rng, ang, tr):
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    T[:3, 3] = tr * rng.normal(size=3)
    return T


@pytest.fixture(scope="module")
def scene(pkg):
    """a voxel-size-0.5 down-sampled synthetic scan: a few thousand points on ground, walls and boxes"""
    src, _, _, _ = pkg.synthetic.make_pair(600, seed=11)
    import oracle as orc
    return np.asarray(orc.voxel_down_sample(src, 0.5))


def test_hybrid_normals(oracle, scene):
    n = oracle.estimate_normals_hybrid(scene, 1.0, 20)
    assert n.shape == scene.shape and np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-12)
    # a dense horizontal patch: normals are +-z; an isolated point: (0, 0, 1) from the identity covariance
    g = np.stack(np.meshgrid(np.arange(20) * 0.1, np.arange(20) * 0.1), -1).reshape(-1, 2)
    patch = np.column_stack([g, np.zeros(len(g))])
    cloud = np.vstack([patch, [[100.0, 100.0, 5.0]]])
    nn = oracle.estimate_normals_hybrid(cloud, 0.35, 20)
    assert np.allclose(np.abs(nn[:-1, 2]), 1.0, atol=1e-9)
    assert np.array_equal(nn[-1], [0.0, 0.0, 1.0])
    # the radius cut is strict (d^2 < r^2) and max_nn caps the list: with max_nn = 20 >= all neighbours inside the radius
    # the result equals the plain 20-NN normals wherever the 20th neighbour is inside the radius
    k20 = oracle.estimate_normals(patch, 20)
    h = oracle.estimate_normals_hybrid(patch, 10.0, 20)
    assert np.array_equal(k20, h)


def test_fpfh_histograms(oracle, scene):
    nrm = oracle.estimate_normals_hybrid(scene, 1.0, 20)
    f = oracle.compute_fpfh_feature(scene, nrm, 5.0, 200)
    assert f.shape == (len(scene), 33) and np.isfinite(f).all() and (f >= 0).all()
    thirds = f.reshape(-1, 3, 11).sum(axis=2)
    has_nb = thirds.sum(axis=1) > 0
    assert has_nb.mean() > 0.95
    # each third: neighbours' SPFH renormalised to 100 + the point's own SPFH (100)
    assert np.allclose(thirds[has_nb], 200.0, atol=1e-9)
    # invariant under a rigid motion of points and normals (bin-boundary flips aside)
    T = _rigid(np.random.default_rng(0), 0.7, 5.0)
    f2 = oracle.compute_fpfh_feature(scene @ T[:3, :3].T + T[:3, 3], nrm @ T[:3, :3].T, 5.0, 200)
    same = np.abs(f - f2).max(axis=1) < 1e-6
    assert same.mean() > 0.97
    # ... and under a permutation of the cloud
    perm = np.random.default_rng(1).permutation(len(scene))
    f3 = oracle.compute_fpfh_feature(scene[perm], nrm[perm], 5.0, 200)
    assert (np.abs(f3 - f[perm]).max(axis=1) < 1e-6).mean() > 0.999
    # an isolated point has an all-zero descriptor
    lone = np.vstack([scene[:50], [[1e3, 1e3, 1e3]]])
    fl = oracle.compute_fpfh_feature(lone, np.vstack([nrm[:50], [[0, 0, 1]]]), 5.0, 200)
    assert not fl[-1].any()


@pytest.mark.parametrize("absolute_scale", [True, False])
def test_fgr_recovers_an_exact_rigid_motion(oracle, scene, absolute_scale):
    """target = the source moved rigidly and shuffled, descriptors moved along: every mutual match is correct, the
    graduated non-convexity must land on the motion itself (pins the Jacobian signs, the delta * trans composition, the
    normalisation and its undoing, and the final inversion)"""
    rng = np.random.default_rng(5)
    T = _rigid(rng, 0.4, 3.0)
    nrm = oracle.estimate_normals_hybrid(scene, 1.0, 20)
    fs = oracle.compute_fpfh_feature(scene, nrm, 5.0, 200)
    perm = rng.permutation(len(scene))[: len(scene) - 37]                 # unequal sizes: exercises the source/target swap
    tgt = (scene @ T[:3, :3].T + T[:3, 3])[perm]
    ft = fs[perm]
    for a, b, fa, fb, want in ((scene, tgt, fs, ft, T), (tgt, scene, ft, fs, np.linalg.inv(T))):
        got, nc = oracle.registration_fgr_based_on_feature_matching(a, b, fa, fb, use_absolute_scale=absolute_scale, decrease_mu=True,
                                                                    maximum_correspondence_distance=1.0 if absolute_scale else 0.01,
                                                                    iteration_number=300, maximum_tuple_count=2000, seed=3)
        assert nc == 3 * 2000
        assert np.abs(got - want).max() < 1e-6, np.abs(got - want).max()
    # fewer than 10 correspondences: identity (Open3D returns before optimising)
    got, nc = oracle.registration_fgr_based_on_feature_matching(scene[:3], tgt[:3], fs[:3], ft[:3], maximum_tuple_count=1)
    assert nc <= 3


def test_registro_fgr_on_nclt_fixtures(oracle, pkg):
    """real NCLT clouds, the reference's parameters (voxel 0.1): as good a coarse alignment as the shipped FGR pose"""
    for a, b in ((1, 0), (18, 17)):
        src = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, "nclt", f"s{a}.pcd")).astype(np.float64)
        tgt = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, "nclt", f"s{b}.pcd")).astype(np.float64)
        T_fgr = np.loadtxt(os.path.join(GOLD, "nclt", f"fgr_pose_{a}_{b}.txt"))
        T_ref = np.loadtxt(os.path.join(GOLD, "nclt", f"golden_pose_{a}_{b}.txt"))
        T, nc = oracle.registro_FGR(src, tgt, 0.1, seed=0)
        assert nc == 3 * int(int((len(src) + len(tgt)) / 2) * 0.2)          # the tuple cap is reached
        rot, tr = pkg.synthetic.pose_error(T, T_ref)
        rot_s, tr_s = pkg.synthetic.pose_error(T_fgr, T_ref)
        print(f"pair {a}->{b}: oracle FGR {tr:.3f} m / {rot:.4f} rad from the refined pose, shipped FGR {tr_s:.3f} m / {rot_s:.4f} rad")
        assert tr < 0.2 and rot < 0.03                                       # inside the basin the refinement converges from
        assert tr < tr_s + 0.1 and rot < rot_s + 0.02
        again, _ = oracle.registro_FGR(src, tgt, 0.1, seed=0)
        assert np.array_equal(T, again)                                      # deterministic for a fixed seed


def test_shared_per_point_functions_equal_the_oracle(oracle, pkg):
    """csrc/fpfh_math.cuh (what the CUDA kernels of the feature stage call per point) evaluated on the CPU over the oracle's
    neighbour lists == the independent restatement in oracle/fgr_oracle.c, bit for bit, with the reference's parameters"""
    cloud = pkg.pcd_io.read_pcd_xyz(os.path.join(GOLD, "nclt", "s17.pcd")).astype(np.float64)
    nrm, f = oracle.fpfh_engine(cloud, 0.2, 20, 1.0, 200)
    ref_n = oracle.estimate_normals_hybrid(cloud, 0.2, 20)
    assert np.array_equal(nrm, ref_n)
    assert np.array_equal(f, oracle.compute_fpfh_feature(cloud, ref_n, 1.0, 200))


def test_shared_fgr_functions_equal_the_oracle(oracle, pkg, scene):
    """csrc/fgr_math.cuh (descriptor distance, counter-based tuple test, one correspondence's normal-equation terms, undoing the
    normalisation -- what the FGR kernels call) run sequentially on the CPU == oracle/fgr_oracle.c bit for bit: both pair orders
    (source/target swap), absolute and relative scale"""
    rng = np.random.default_rng(2)
    T = _rigid(rng, 0.3, 2.0)
    nrm = oracle.estimate_normals_hybrid(scene, 1.0, 20)
    fs = oracle.compute_fpfh_feature(scene, nrm, 5.0, 200)
    tgt = (scene @ T[:3, :3].T + T[:3, 3])[rng.permutation(len(scene))[:-50]] + 0.01 * rng.normal(size=(len(scene) - 50, 3))
    ft = oracle.compute_fpfh_feature(tgt, oracle.estimate_normals_hybrid(tgt, 1.0, 20), 5.0, 200)
    for a, b, fa, fb in ((scene, tgt, fs, ft), (tgt, scene, ft, fs)):
        for absolute in (True, False):
            kw = dict(use_absolute_scale=absolute, decrease_mu=True, maximum_correspondence_distance=1.0 if absolute else 0.01,
                      iteration_number=100, maximum_tuple_count=500, seed=11)
            T0, n0 = oracle.registration_fgr_based_on_feature_matching(a, b, fa, fb, **kw)
            T1, n1 = oracle.registration_fgr_based_on_feature_matching(a, b, fa, fb, engine=True, **kw)
            assert n0 == n1 and n0 > 30
            assert np.array_equal(T0, T1)
            # the kernel's reduction order (512 strided partials, shuffle tree, warps in order) moves the pose by rounding only
            T2, n2 = oracle.registration_fgr_based_on_feature_matching(a, b, fa, fb, engine="kernel_order", **kw)
            assert n2 == n0 and np.abs(T2 - T0).max() < 1e-10


def test_fgr_pin_summary():
    """the committed soft pin over a sample of the 900 consecutive NCLT pairs (oracle/pin_fgr_against_goldens.py)"""
    p = os.path.join(GOLD, "nclt_fgr_pin.json")
    s = json.load(open(p))["summary"]
    assert s["pairs"] >= 50
    # the oracle's FGR is statistically as close to the refined poses as the reference's own FGR
    assert s["oracle_m_p50"] < 1.5 * s["shipped_m_p50"] + 0.02 and s["oracle_m_p90"] < 1.5 * s["shipped_m_p90"] + 0.05
    assert s["oracle_rad_p50"] < 1.5 * s["shipped_rad_p50"] + 0.005


def test_division_by_reciprocal_is_the_division(oracle):
    """k_fpfh divides 33 SPFH values by every neighbour's squared distance; the kernel (and the CPU engine through the same
    header) forms one reciprocal per neighbour and corrects each product twice (fpfh_math.cuh:div_by_recip).  The result must be
    the IEEE quotient for every operand pair: 6e7 pairs, three seeds."""
    for seed in (1, 2, 3):
        assert oracle.check_recip_div(20_000_000, seed) == 0
