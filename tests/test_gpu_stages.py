"""Stage-by-stage parity of the CUDA path against the CPU oracle, through the C ABI (gpu tests).

Rules (BASELINE.json north_star / SURVEY.md 8(c)): down-sampled point SETS bit-exact; kNN neighbour SETS
identical except for exact-distance ties; normals equal incl. sign; normal-equation sums to 1e-10 relative.
"""
import numpy as np
import pytest

from conftest import match_rows, sort_rows

pytestmark = pytest.mark.gpu

VOXELS = [1.0, 0.5, 0.25]


@pytest.fixture(scope="module")
def prepped(engine, pkg, pair30k):
    from mgicp_b200 import _lib
    src, tgt, T_init, T_true = pair30k
    opts = engine.make_opts(debug=True)
    flat, off, _ = engine.pack_clouds([src, tgt])
    xyz = engine.upload(flat)
    engine.preprocess_device(xyz, off, VOXELS, opts)
    engine.check()
    return dict(engine=engine, lib=_lib, clouds=[src, tgt], opts=opts, n=[len(src), len(tgt)])


@pytest.mark.parametrize("cloud", [0, 1])
@pytest.mark.parametrize("scale", [0, 1, 2])
def test_voxel_down_sample_bit_exact(prepped, oracle, cloud, scale):
    e, L = prepped["engine"], prepped["lib"]
    got = e.get_stage(cloud, scale, L.STAGE_DOWNSAMPLED, prepped["n"][cloud])
    ref = oracle.voxel_down_sample(prepped["clouds"][cloud], VOXELS[scale])
    assert got.shape == ref.shape
    assert np.array_equal(sort_rows(got), sort_rows(ref))          # bit-exact point set
    grid = e.get_stage(cloud, scale, L.STAGE_GRID_POINTS, prepped["n"][cloud])
    assert np.array_equal(sort_rows(grid), sort_rows(ref))          # the spatial hash holds the same set


def test_bounds(prepped):
    e, L = prepped["engine"], prepped["lib"]
    for c in (0, 1):
        b = e.get_stage(c, 0, L.STAGE_BOUNDS, 1)
        pts = prepped["clouds"][c]
        assert np.array_equal(b[:3], pts.min(axis=0)) and np.array_equal(b[3:], pts.max(axis=0))


@pytest.mark.parametrize("cloud", [0, 1])
@pytest.mark.parametrize("scale", [0, 1, 2])
def test_statistical_outlier_removal(prepped, oracle, cloud, scale):
    e, L = prepped["engine"], prepped["lib"]
    n = prepped["n"][cloud]
    grid = e.get_stage(cloud, scale, L.STAGE_GRID_POINTS, n)
    avg = e.get_stage(cloud, scale, L.STAGE_SOR_AVG, n)
    keep = e.get_stage(cloud, scale, L.STAGE_SOR_KEEP, n).astype(bool)
    kept_ref, mask_ref, avg_ref, thr = oracle.remove_statistical_outlier(grid, 30, 1.0)
    # mean neighbour distances: same neighbours, summed in the same (ascending) order -> identical up to ties
    assert np.allclose(avg, avg_ref, rtol=0, atol=1e-13)
    assert (avg == avg_ref).mean() > 0.999
    assert np.array_equal(keep, mask_ref)
    final = e.get_stage(cloud, scale, L.STAGE_POINTS, n)
    assert np.array_equal(final, kept_ref)                           # order-preserving compaction


@pytest.mark.parametrize("cloud,scale", [(0, 0), (1, 1), (0, 2)])
def test_knn_neighbour_sets(prepped, oracle, cloud, scale):
    e, L = prepped["engine"], prepped["lib"]
    n = prepped["n"][cloud]
    for what, stage_pts, k in ((L.STAGE_KNN_SOR, L.STAGE_GRID_POINTS, 30), (L.STAGE_KNN_NORMAL, L.STAGE_POINTS, 20)):
        pts = e.get_stage(cloud, scale, stage_pts, n)
        got = e.get_stage(cloud, scale, what, n, k)
        ref_idx, ref_d2, _ = oracle.knn(pts, pts, k + 1)
        tie = ref_d2[:, k - 1] == ref_d2[:, k]                       # k-th and (k+1)-th equidistant: excluded by the rules
        same = np.array([set(g) == set(r[:k]) for g, r in zip(got, ref_idx)])
        assert same[~tie].all(), f"{(~same & ~tie).sum()} neighbour sets differ"
        assert tie.mean() < 0.01


@pytest.mark.parametrize("cloud", [0, 1])
@pytest.mark.parametrize("scale", [0, 1, 2])
def test_normals(prepped, oracle, cloud, scale):
    e, L = prepped["engine"], prepped["lib"]
    n = prepped["n"][cloud]
    pts = e.get_stage(cloud, scale, L.STAGE_POINTS, n)
    nrm = e.get_stage(cloud, scale, L.STAGE_NORMALS, n)
    ref = oracle.estimate_normals(pts, 20)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-12)
    err = np.abs(nrm - ref).max(axis=1)                              # sign included
    # same neighbours in the same (ascending-distance) order, same fdlibm trig kernels: bit-identical normals,
    # except where two neighbours are exactly equidistant (summation order of the cumulants then differs)
    assert (err == 0).mean() > 0.999, (err != 0).sum()
    assert err.max() < 1e-6


@pytest.mark.parametrize("loss", ["l1", "l2"])
def test_single_pass_normal_equations(prepped, oracle, pair30k, loss):
    """one correspondence pass + linearisation at T_init vs the oracle's first iteration (1e-10 relative)"""
    e, L = prepped["engine"], prepped["lib"]
    _, _, T_init, _ = pair30k
    for scale, max_d in ((0, 3.0), (2, 0.25)):
        sp = e.get_stage(0, scale, L.STAGE_POINTS, prepped["n"][0])
        sn = e.get_stage(0, scale, L.STAGE_NORMALS, prepped["n"][0])
        tp = e.get_stage(1, scale, L.STAGE_POINTS, prepped["n"][1])
        tn = e.get_stage(1, scale, L.STAGE_NORMALS, prepped["n"][1])
        ref = oracle.gicp(sp, sn, tp, tn, max_d, T_init, 1, loss=loss, want_trace=True)
        opts = e.make_opts(loss=loss)
        got = e.evaluate(scale, [0], [1], [max_d], T_init.reshape(1, 4, 4), opts)
        assert got["K"][0] == ref.trace[0, 2]
        assert abs(got["fitness"][0] - ref.trace[0, 0]) < 1e-15
        assert abs(got["rmse"][0] - ref.trace[0, 1]) < 1e-12
        s_ref = ref.sys_trace[0]
        scale_ = np.abs(s_ref).max()
        # L2: pure rounding (closed-form W vs the oracle's inverse().sqrt(), different summation order).
        # L1: rows with r -> 0 carry weights 1/|r| up to ~1e9, which turn the 1e-16 rounding of r into ~1e-9 of the sums.
        tol = 1e-12 if loss == "l2" else 1e-8
        assert np.abs(got["sums"][0] - s_ref).max() / scale_ < tol


def test_float64_clouds_ordered_voxel_sums(pkg, oracle, engine):
    """Coordinates that need all 53 bits (not float32-representable): the voxel sums must not depend on the order in which
    atomics arrive.  The engine then sums every voxel's points in input order (Open3D's AccumulatedPoint order): centroids
    bit-identical to the oracle's, two runs bit-identical, and the whole path bit-identical run to run under L1."""
    from mgicp_b200 import _lib as L
    src, tgt, T_init, _ = pkg.synthetic.make_pair(400, seed=21)
    rng = np.random.default_rng(3)
    src64 = src + rng.uniform(-1e-4, 1e-4, src.shape)            # no longer float32-representable
    tgt64 = tgt + rng.uniform(-1e-4, 1e-4, tgt.shape)
    assert not np.array_equal(src64.astype(np.float32).astype(np.float64), src64)
    vox = [1.0, 0.5, 0.25]
    opts = engine.make_opts(loss="l1")
    runs = []
    for _ in range(2):
        flat, off, code = engine.pack_clouds([src64, tgt64])
        assert code == L.F64
        engine.preprocess_device(engine.upload(flat), off, vox, opts)
        engine.check()
        runs.append([engine.get_stage(c, s, L.STAGE_DOWNSAMPLED, len(src64) + len(tgt64)) for c in (0, 1) for s in range(3)])
    for a, b in zip(*runs):
        assert np.array_equal(a, b)
    for c, cloud in enumerate((src64, tgt64)):
        for s, v in enumerate(vox):
            ref = sort_rows(oracle.voxel_down_sample(cloud, v))
            got = sort_rows(runs[0][c * 3 + s])
            assert got.shape == ref.shape and np.array_equal(got, ref), (c, s, np.abs(got - ref).max())
    r1 = pkg.multiscale_gicp(src64, tgt64, vox, [3.0, 1.0, 0.25], 60, T_init, engine=engine)
    r2 = pkg.multiscale_gicp(src64, tgt64, vox, [3.0, 1.0, 0.25], 60, T_init, engine=engine)
    assert np.array_equal(r1.transformation, r2.transformation) and r1.iterations == r2.iterations


@pytest.mark.parametrize("sor_k,normal_k", [(8, 5), (32, 20), (30, 32)])
def test_other_neighbour_counts_and_odd_clouds(pkg, oracle, engine, sor_k, normal_k):
    """The histogram-selection kNN for other k (1..32), for clouds smaller than k, for a very sparse cloud (every query needs
    rings 2..3 or the whole-cloud scan) and for a cloud with exact duplicates of distances (a regular lattice: ties)."""
    from mgicp_b200 import _lib as L
    src, _, _, _ = pkg.synthetic.make_pair(300, seed=9)
    rng = np.random.default_rng(sor_k)
    sparse = rng.uniform(-400.0, 400.0, (700, 3)) * np.array([1.0, 1.0, 0.05])       # ~1 point per 30 x 30 m: far apart at voxel 1
    tiny = src[:7].copy()
    gx, gy = np.meshgrid(np.arange(24) * 0.5, np.arange(24) * 0.5)
    lattice = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1) + 0.25
    clouds = [src, sparse, tiny, lattice]
    opts = engine.make_opts(sor_k=sor_k, normal_k=normal_k, debug=True)
    flat, off, _ = engine.pack_clouds(clouds)
    vox = [1.0, 0.5]
    engine.preprocess_device(engine.upload(flat), off, vox, opts)
    engine.check()
    for c, cloud in enumerate(clouds):
        for s, v in enumerate(vox):
            n = len(cloud)
            grid = engine.get_stage(c, s, L.STAGE_GRID_POINTS, n)
            avg = engine.get_stage(c, s, L.STAGE_SOR_AVG, n)
            keep = engine.get_stage(c, s, L.STAGE_SOR_KEEP, n).astype(bool)
            lists = engine.get_stage(c, s, L.STAGE_KNN_SOR, n, sor_k)
            kk = min(sor_k, len(grid))
            ref_idx, ref_d2, _ = oracle.knn(grid, grid, min(kk + 1, len(grid)))
            # mean distance over the k nearest (the query itself included), exact up to ties at the k-th place
            ref_avg = np.sqrt(ref_d2[:, :kk]).sum(axis=1) / kk
            tie = (ref_d2[:, kk - 1] == ref_d2[:, kk]) if ref_d2.shape[1] > kk else np.zeros(len(grid), bool)
            assert np.allclose(avg, ref_avg, rtol=0, atol=1e-12), (c, s, np.abs(avg - ref_avg).max())
            got_sets = [set(int(t) for t in row if t >= 0) for row in lists]
            assert all(len(g) == kk for g in got_sets), (c, s)
            same = np.array([g == set(r[:kk].tolist()) for g, r in zip(got_sets, ref_idx)])
            assert same[~tie].all(), (c, s, int((~same & ~tie).sum()))
            if len(grid) > 1 and not tie.any():
                _, mask_ref, _, _ = oracle.remove_statistical_outlier(grid, sor_k, 1.0)
                assert np.array_equal(keep, mask_ref), (c, s)
