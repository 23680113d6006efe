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
