"""The C-ABI boundary and the package's streaming API, on the GPU.

* mgicp_run_batch called exactly as INTEGRATION.md section 2 shows (raw ctypes, no Engine) equals
  mgicp_preprocess + mgicp_register_batch bit for bit;
* BatchStream (packing into pinned memory, upload / compute / download on three streams, two alternating engines)
  returns, batch by batch, exactly what the synchronous Engine.run returns;
* device-side error flags travel through the stream-ordered mgicp_job_errors;
* stage timing (mgicp_set_timing / mgicp_get_timing)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

VOXELS, DISTS = [1.0, 0.5, 0.25], [3.0, 1.0, 0.25]


def test_run_batch_as_in_integration_md(pkg, engine, pair_small):
    import torch
    from mgicp_b200 import _lib
    src, tgt, T_ini, _ = pair_small
    L = C.CDLL(_lib.lib_path())                                    # a maintainer's own binding: nothing from engine.py
    L.mgicp_last_error.restype = C.c_char_p

    class Opts(C.Structure):                                       # mgicp_opts, include/mgicp.h
        _fields_ = [("sor_k", C.c_int32), ("sor_std", C.c_double), ("normal_k", C.c_int32), ("epsilon", C.c_double),
                    ("loss", C.c_int32), ("loss_k", C.c_double), ("rel_fitness", C.c_double), ("rel_rmse", C.c_double),
                    ("cell_factor", C.c_double), ("icp_cell_factor", C.c_double), ("ctas_per_pair", C.c_int32), ("debug", C.c_int32)]
    h = C.c_void_p()
    assert L.mgicp_create(0, C.byref(h)) == 0
    o = Opts()
    L.mgicp_default_opts(C.byref(o))
    xyz = torch.from_numpy(np.concatenate([src, tgt]).astype(np.float32)).cuda()
    off = (C.c_int64 * 3)(0, len(src), len(src) + len(tgt))
    vox = (C.c_double * 3)(*VOXELS); md = (C.c_double * 3)(*DISTS); mi = (C.c_int32 * 3)(100, 100, 100)
    ps, pt = (C.c_int32 * 1)(0), (C.c_int32 * 1)(1)
    T0 = torch.from_numpy(np.ascontiguousarray(T_ini).reshape(1, 16)).cuda(); T = torch.empty_like(T0)
    fit = torch.empty(1, dtype=torch.float64, device="cuda"); rm = torch.empty_like(fit)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.mgicp_run_batch(h, st, 2, C.c_void_p(xyz.data_ptr()), off, 0, 3, vox, 1, ps, pt, md, mi, C.byref(o),
                           C.c_void_p(T0.data_ptr()), C.c_void_p(T.data_ptr()), C.c_void_p(fit.data_ptr()),
                           C.c_void_p(rm.data_ptr()), None, None, None)
    assert rc == 0, L.mgicp_last_error(h)
    assert L.mgicp_check(h) == 0
    # the same through preprocess + register (engine.py), float32 clouds, default options (L1)
    ref = engine.run([src.astype(np.float32), tgt.astype(np.float32)], [(0, 1)], VOXELS, DISTS, 100, T_ini[None])
    assert np.array_equal(T.cpu().numpy().reshape(4, 4), ref.transformation[0])
    assert fit.item() == ref.fitness[0] and rm.item() == ref.inlier_rmse[0]
    assert L.mgicp_destroy(h) == 0


def test_batch_stream_equals_synchronous_run(pkg, engine):
    """five batches of different sizes (float32 and float64 clouds) through the pipeline == Engine.run per batch"""
    batches = []
    for b, (n_az, dt) in enumerate([(250, np.float32), (400, np.float64), (180, np.float32), (300, np.float32), (220, np.float64)]):
        clouds, pairs, T0 = [], [], []
        for k in range(3):
            s, t, Ti, _ = pkg.synthetic.make_pair(n_az, seed=10 * b + k)
            clouds += [s.astype(dt), t.astype(dt)]
            pairs.append((2 * k, 2 * k + 1))
            T0.append(Ti)
        batches.append((clouds, pairs, np.stack(T0)))
    keep = [[c.copy() for c in b[0]] for b in batches]
    bs = pkg.BatchStream(VOXELS, DISTS, 60, device=0, loss="l1")
    got = list(bs.run(batches))
    bs.close()
    assert len(got) == len(batches) and bs.h2d_bytes > 0 and bs.d2h_bytes > 0
    opts = engine.make_opts(loss="l1")
    for b, r in zip(batches, got):
        ref = engine.run(b[0], b[1], VOXELS, DISTS, 60, b[2], opts)
        assert np.array_equal(r.transformation, ref.transformation)
        assert np.array_equal(r.fitness, ref.fitness) and np.array_equal(r.inlier_rmse, ref.inlier_rmse)
        assert np.array_equal(r.iterations, ref.iterations) and np.array_equal(r.num_correspondences, ref.num_correspondences)
        assert np.array_equal(r.stats, ref.stats)
    for b, k in zip(batches, keep):                                # inputs untouched (AF:289-290)
        assert all(np.array_equal(x, y) for x, y in zip(b[0], k))
    # the reference-facing batched call goes through the same pipeline
    r1 = pkg.multiscale_gicp_batch(batches[0][0], batches[0][1], VOXELS, DISTS, 60, batches[0][2], engine=engine, loss="l1")
    assert np.array_equal(r1.transformation, got[0].transformation)


def test_stream_reports_device_side_errors(pkg, engine):
    """a voxel size too small for the extent raises RuntimeError like Open3D, through the stream-ordered error flag"""
    s, t, Ti, _ = pkg.synthetic.make_pair(100, seed=1)
    far = np.vstack([s, [[3.0e6, 0.0, 0.0]]])
    bs = pkg.BatchStream([1.0, 0.5, 0.25], DISTS, 5, engine=engine, engines=1)
    with pytest.raises(RuntimeError):
        bs.run_one([far, t], [(0, 1)], Ti[None])
    ok = bs.run_one([s, t], [(0, 1)], Ti[None])                    # the stream stays usable
    bs.close()
    assert ok.fitness[0] > 0


def test_upload_reuses_pinned_staging_safely(pkg, engine):
    """two uploads with the same (dtype, numel) key back to back: the second must not overwrite the first in flight"""
    import torch
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(3_000_000,)), rng.normal(size=(3_000_000,))
    da = engine.upload(a)
    db = engine.upload(b)
    torch.cuda.synchronize()
    assert np.array_equal(da.cpu().numpy(), a) and np.array_equal(db.cpu().numpy(), b)


def test_stage_timing(pkg, engine, pair_small):
    src, tgt, T_ini, _ = pair_small
    engine.set_timing(True)
    try:
        engine.run([src, tgt], [(0, 1)], VOXELS, DISTS, 30, T_ini[None])
        t = engine.get_timing()
    finally:
        engine.set_timing(False)
    print("stage timing (ms):", {k: (round(v, 4) if not isinstance(v, list) else [round(x, 4) for x in v]) for k, v in t.items()})
    assert all(t[k] > 0 for k in ("downsample_ms", "sor_ms", "normals_ms", "icp_ms"))
    assert all(x > 0 for x in t["scale_ms"][:3]) and sum(t["scale_ms"][:3]) <= t["icp_ms"] * 1.05
