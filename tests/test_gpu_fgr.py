"""Parity of the FGR front end's feature stage (hybrid-radius normals + FPFH, csrc/mgicp_fgr.cuh) against the oracle.

These kernels were written after the round's GPU budget was spent: they compile for sm_100a and their per-point arithmetic
is checked on the CPU (tests/test_fgr_oracle.py::test_shared_per_point_functions_equal_the_oracle), but they have not run on
a GPU yet.  Until their first green run the tests are opt-in: MGICP_RUN_UNVERIFIED=1 python -m pytest tests -m gpu."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MGICP_RUN_UNVERIFIED") != "1", reason="first GPU run pending (set MGICP_RUN_UNVERIFIED=1)")]

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nclt")


def _compare(oracle, cloud, nrm, fp, rn, kn, rf, kf):
    ref_n = oracle.estimate_normals_hybrid(cloud, rn, kn)
    assert np.array_equal(nrm, ref_n)                                    # same lists, same order, same closed-form eigenvector
    ref_f = oracle.compute_fpfh_feature(cloud, ref_n, rf, kf)
    # acos / atan2 come from libdevice on the GPU and libm in the oracle: a neighbour sitting on a bin boundary may move
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
