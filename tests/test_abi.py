"""The C-ABI library loads on a CPU-only machine and exports every symbol include/mgicp.h declares (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    from mgicp_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mgicp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mgicp_[a-z_]+)\s*\(", hdr)))
    assert len(declared) >= 13
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/mgicp.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert b"sm_100a" in L.mgicp_version()


def test_default_opts_are_the_reference_constants(pkg):
    from mgicp_b200 import _lib
    o = _lib.Opts()
    _lib.load().mgicp_default_opts(C.byref(o))
    # knn_filtro=30, std_filtro=1.0 (AF:280-281); knn=20 (AF:301); L1Loss (AF:284); 1e-6/1e-6 (AF:309-310); epsilon 1e-3
    assert (o.sor_k, o.sor_std, o.normal_k, o.epsilon, o.loss, o.rel_fitness, o.rel_rmse) == (30, 1.0, 20, 1e-3, 1, 1e-6, 1e-6)


def test_icp_cell_factor_follows_the_schedule(pkg):
    """mgicp_auto_icp_cell_factor (host arithmetic, no GPU): the reference's two schedules and what lies between them"""
    from mgicp_b200 import _lib
    L = _lib.load()
    dp = C.POINTER(C.c_double)

    def factor(voxels, dists):
        v = np.ascontiguousarray(voxels, np.float64)
        d = np.ascontiguousarray(dists, np.float64).reshape(-1, len(v))
        return L.mgicp_auto_icp_cell_factor(len(v), v.ctypes.data_as(dp), d.shape[0], d.ctypes.data_as(dp))

    vox = pkg.create_scales_script2(5)                                  # 2_MGICP...py:102-120: radii of at most 3 voxels
    assert factor(vox, pkg.max_correspondence_distances(vox)) == 3.5
    assert factor([1.0, 0.5, 0.25], [3.0, 1.0, 0.25]) == 3.5            # the bench's three scales
    af = pkg.create_scales(3)
    af.reverse()                                                        # ALL_FUNCTIONS.py:260-278: 0.4 / 0.2 / 0.1, radius = the cloud's size
    assert factor(af, [[44.7, 22.4, 11.2], [30.0, 15.0, 7.5]]) == 16.0
    assert factor([1.0, 0.5], [[8.0, 2.0], [5.0, 3.0]]) == 8.0          # in between: the largest radius / voxel of the batch
    assert L.mgicp_auto_icp_cell_factor(0, None, 0, None) == 3.5


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.MgicpError):
        pkg.Engine()
    from mgicp_b200 import _lib
    h = C.c_void_p()
    assert _lib.load().mgicp_create(0, C.byref(h)) != 0 and not h.value
    with pytest.raises(pkg.MgicpError):
        pkg.Multiscale_GICP(np.zeros((10, 3)), np.zeros((10, 3)), 3, 10, np.eye(4))


def test_product_package_never_imports_the_oracle():
    pk = os.path.join(ROOT, "point-cloud-registration-with-global-refinement_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "libmgicp_oracle" not in txt and "orc_" not in txt, f


def test_pcd_and_pose_io_roundtrip(pkg, tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(100, 3)).astype(np.float32)
    p = str(tmp_path / "a.pcd")
    pkg.pcd_io.write_pcd_xyz(p, pts)
    back = pkg.pcd_io.read_pcd_xyz(p)
    assert back.dtype == np.float64 and np.array_equal(back, pts.astype(np.float64))
    nclt = pkg.pcd_io.read_pcd_xyz(os.path.join(ROOT, "tests", "golden", "nclt", "s0.pcd"))
    assert nclt.shape == (18421, 3)
    T = np.eye(4)
    T[:3, 3] = [1.5, -2.25, 1e-9]
    q = str(tmp_path / "pose.txt")
    pkg.pcd_io.write_pose(q, T)                        # %.18e like the reference's np.savetxt default
    assert np.array_equal(pkg.pcd_io.read_pose(q), T)
