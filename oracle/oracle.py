"""ctypes front end of the CPU oracle (oracle/mgicp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports this module.

The Python-level orchestration mirrors the reference's two ``Multiscale_GICP`` definitions
(/root/reference/ALL_FUNCTIONS.py:272-313 and /root/reference/2_MGICP_refinement_in_NCLT_dataset.py:128-164).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmgicp_oracle.so")
_lib = None

LOSS = {"l2": 0, "l1": 1, "huber": 2, "cauchy": 3, "gm": 4, "tukey": 5}


class _GicpOpts(C.Structure):
    _fields_ = [("epsilon", C.c_double), ("loss", C.c_int), ("loss_k", C.c_double), ("rel_fitness", C.c_double),
                ("rel_rmse", C.c_double), ("max_iteration", C.c_int)]


class _Opts(C.Structure):
    _fields_ = [("sor_k", C.c_int), ("sor_std", C.c_double), ("normal_k", C.c_int), ("gicp", _GicpOpts)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "mgicp_oracle.c"), os.path.join(_HERE, "engine_order.cpp"), os.path.join(_HERE, "fgr_oracle.c"), os.path.join(_HERE, "fpfh_engine.cpp"),
            os.path.join(os.path.dirname(_HERE), "point-cloud-registration-with-global-refinement_b200", "csrc", "fpfh_math.cuh"),
            os.path.join(os.path.dirname(_HERE), "point-cloud-registration-with-global-refinement_b200", "csrc", "fgr_math.cuh"),
            os.path.join(os.path.dirname(_HERE), "point-cloud-registration-with-global-refinement_b200", "csrc", "mgicp_math.cuh")]
    stale = not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(f) for f in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building the oracle failed:\n" + r.stdout + r.stderr)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _check(rc, what):
    if rc == 1:
        raise RuntimeError(f"oracle {what}: invalid argument (Open3D raises RuntimeError here)")
    if rc != 0:
        raise MemoryError(f"oracle {what}: rc={rc}")


def set_sum_chunk(n: int) -> None:
    """summation chunk of the normal equations (default 1024); changing it only re-associates the sums"""
    lib().orc_set_sum_chunk(C.c_int64(n))


def gicp_engine_order(src_xyz, src_nrm, tgt_xyz, tgt_nrm, max_d, T_init, max_iteration, *, cl=1, epsilon=1e-3, loss="l1",
                      loss_k=1.0, rel_fitness=1e-6, rel_rmse=1e-6, want_trace=False):
    """registration_generalized_icp evaluated in the CUDA kernel's arithmetic order (oracle/engine_order.cpp)."""
    s, sn, t, tn = (_d(a).reshape(-1, 3) for a in (src_xyz, src_nrm, tgt_xyz, tgt_nrm))
    T0 = _d(T_init).reshape(16)
    T = np.empty(16)
    fit, rm, it, K = C.c_double(), C.c_double(), C.c_int32(), C.c_int64()
    trace = np.zeros((max_iteration + 1, 3)) if want_trace else None
    rc = lib().orc_gicp_engine_order(_p(s), _p(sn), C.c_int64(s.shape[0]), _p(t), _p(tn), C.c_int64(t.shape[0]),
                                     C.c_double(max_d), _p(T0), C.c_double(epsilon), C.c_int(LOSS[loss]), C.c_double(loss_k),
                                     C.c_double(rel_fitness), C.c_double(rel_rmse), C.c_int(max_iteration), C.c_int(cl), _p(T),
                                     C.byref(fit), C.byref(rm), C.byref(it), C.byref(K), _p(trace) if want_trace else None)
    _check(rc, "gicp_engine_order")
    return OracleResult(T.reshape(4, 4), fit.value, rm.value, [it.value], K.value,
                        trace=trace[: it.value + 1] if want_trace else None)


def multiscale_gicp_engine_order(stages, max_corr_dists, max_iters, T_init, *, cl=1, **kw):
    """Chain gicp_engine_order over scales.  stages[s] = (src_points, src_normals, tgt_points, tgt_normals) of scale s,
    in the engine's point order (the order fixes the thread-strided partial sums)."""
    T = _d(T_init).reshape(4, 4)
    res, iters = None, []
    for s, (sp, sn, tp, tn) in enumerate(stages):
        res = gicp_engine_order(sp, sn, tp, tn, max_corr_dists[s], T, int(max_iters[s]), cl=cl, **kw)
        T = res.transformation
        iters.append(res.iterations[0])
    res.iterations = iters
    return res


def probe(name, *arrays, out_len):
    """call a host probe of csrc/mgicp_math.cuh (probe_* in engine_order.cpp)"""
    out = np.zeros(out_len)
    args = []
    for a in arrays:
        if np.isscalar(a):
            args.append(C.c_double(a))
        else:
            a = _d(a)
            args.append(_p(a))
    getattr(lib(), "probe_" + name)(*args, _p(out))
    return out


def det_trig(x):
    L = lib()
    for f in (L.orc_det_sin, L.orc_det_cos, L.orc_det_acos):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    return L.orc_det_sin(x), L.orc_det_cos(x), L.orc_det_acos(min(max(x, -1.0), 1.0))


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))


def stage_reset() -> None:
    lib().orc_stage_reset()


def stage_seconds() -> dict:
    """wall-clock split of the multiscale_gicp calls since stage_reset(): down-sample / outlier removal / normals / ICP loop,
    plus ICP seconds and passes (iterations + 1) per scale index"""
    out = np.zeros(20)
    lib().orc_stage_seconds(_p(out))
    return {"downsample_s": out[0], "sor_s": out[1], "normals_s": out[2], "icp_s": out[3],
            "icp_s_per_scale": out[4:12].tolist(), "icp_passes_per_scale": out[12:20].tolist()}


def voxel_down_sample(xyz, voxel, return_index=False):
    xyz = _d(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    out = np.empty((max(n, 1), 3), np.float64)
    vox = np.empty((max(n, 1), 3), np.int32)
    m = C.c_int64(0)
    rc = lib().orc_voxel_down_sample(_p(xyz), C.c_int64(n), C.c_double(voxel), _p(out), _p(vox, C.c_int32), C.byref(m))
    _check(rc, "voxel_down_sample")
    if return_index:
        return out[: m.value].copy(), vox[: m.value].copy()
    return out[: m.value].copy()


def knn(xyz, queries, k):
    xyz = _d(xyz).reshape(-1, 3)
    q = _d(queries).reshape(-1, 3)
    idx = np.empty((q.shape[0], k), np.int32)
    d2 = np.empty((q.shape[0], k), np.float64)
    cnt = np.empty((q.shape[0],), np.int32)
    rc = lib().orc_knn(_p(xyz), C.c_int64(xyz.shape[0]), _p(q), C.c_int64(q.shape[0]), C.c_int(k), _p(idx, C.c_int32),
                       _p(d2), _p(cnt, C.c_int32))
    _check(rc, "knn")
    return idx, d2, cnt


def remove_statistical_outlier(xyz, k=30, ratio=1.0):
    """returns (kept_points, keep_mask, avg_dist, threshold)"""
    xyz = _d(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    keep = np.zeros((max(n, 1),), np.uint8)
    avg = np.zeros((max(n, 1),), np.float64)
    kept = C.c_int64(0)
    thr = C.c_double(0)
    rc = lib().orc_remove_statistical_outlier(_p(xyz), C.c_int64(n), C.c_int(k), C.c_double(ratio), _p(keep, C.c_uint8),
                                              _p(avg), C.byref(kept), C.byref(thr))
    _check(rc, "remove_statistical_outlier")
    mask = keep[:n].astype(bool)
    return xyz[mask].copy(), mask, avg[:n].copy(), thr.value


def estimate_normals(xyz, k=20):
    xyz = _d(xyz).reshape(-1, 3)
    out = np.empty_like(xyz)
    rc = lib().orc_estimate_normals(_p(xyz), C.c_int64(xyz.shape[0]), C.c_int(k), _p(out))
    _check(rc, "estimate_normals")
    return out


def fast_eigen3x3(cov6):
    cov6 = _d(cov6).reshape(6)
    out = np.empty(3)
    lib().orc_fast_eigen3x3(_p(cov6), _p(out))
    return out


def covariance_from_normal(n, eps=1e-3):
    n = _d(n).reshape(3)
    out = np.empty(9)
    lib().orc_covariance_from_normal(_p(n), C.c_double(eps), _p(out))
    return out.reshape(3, 3)


def ldlt_solve6(A, b):
    A = _d(A).reshape(36)
    b = _d(b).reshape(6)
    x = np.empty(6)
    lib().orc_ldlt_solve6(_p(A), _p(b), _p(x))
    return x


def vec6_to_mat4(x):
    x = _d(x).reshape(6)
    T = np.empty(16)
    lib().orc_vec6_to_mat4(_p(x), _p(T))
    return T.reshape(4, 4)


@dataclass
class OracleResult:  # noqa: D101
    """Mirrors Open3D's RegistrationResult fields the reference reads (S2:198,218; AF:331,357,366,369)."""
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    iterations: list = field(default_factory=list)
    num_correspondences: int = 0
    stats: np.ndarray | None = None
    trace: np.ndarray | None = None
    sys_trace: np.ndarray | None = None


def _gopts(max_iteration, epsilon, loss, loss_k, rel_fitness, rel_rmse):
    return _GicpOpts(epsilon, LOSS[loss], loss_k, rel_fitness, rel_rmse, int(max_iteration))


def gicp(src_xyz, src_nrm, tgt_xyz, tgt_nrm, max_d, T_init, max_iteration, *, epsilon=1e-3, loss="l1", loss_k=1.0,
         rel_fitness=1e-6, rel_rmse=1e-6, want_trace=False):
    """registration_generalized_icp on clouds that already carry normals (AF:304-311)."""
    s, sn, t, tn = (_d(a).reshape(-1, 3) for a in (src_xyz, src_nrm, tgt_xyz, tgt_nrm))
    T0 = _d(T_init).reshape(16)
    o = _gopts(max_iteration, epsilon, loss, loss_k, rel_fitness, rel_rmse)
    T = np.empty(16)
    fit, rm, it, K = C.c_double(), C.c_double(), C.c_int32(), C.c_int64()
    trace = np.zeros((max_iteration + 1, 3)) if want_trace else None
    sys_trace = np.zeros((max(max_iteration, 1), 27)) if want_trace else None
    rc = lib().orc_gicp(_p(s), _p(sn), C.c_int64(s.shape[0]), _p(t), _p(tn), C.c_int64(t.shape[0]), C.c_double(max_d),
                        _p(T0), C.byref(o), _p(T), C.byref(fit), C.byref(rm), C.byref(it), C.byref(K),
                        _p(trace) if want_trace else None, _p(sys_trace) if want_trace else None)
    _check(rc, "gicp")
    return OracleResult(T.reshape(4, 4), fit.value, rm.value, [it.value], K.value,
                        trace=trace[: it.value + 1] if want_trace else None,
                        sys_trace=sys_trace[: it.value] if want_trace else None)


def multiscale_gicp(source, target, voxel_sizes, max_corr_dists, max_iters, T_init, *, sor_k=30, sor_std=1.0,
                    normal_k=20, epsilon=1e-3, loss="l1", loss_k=1.0, rel_fitness=1e-6, rel_rmse=1e-6):
    """Body of Multiscale_GICP (AF:286-312) with an explicit schedule."""
    s = _d(getattr(source, "points", source)).reshape(-1, 3)
    t = _d(getattr(target, "points", target)).reshape(-1, 3)
    S = len(voxel_sizes)
    if np.isscalar(max_iters):
        max_iters = [int(max_iters)] * S
    vs, md = _d(voxel_sizes), _d(max_corr_dists)
    mi = np.ascontiguousarray(max_iters, np.int32)
    o = _Opts(sor_k, sor_std, normal_k, _gopts(0, epsilon, loss, loss_k, rel_fitness, rel_rmse))
    T0 = _d(T_init).reshape(16)
    T = np.empty(16)
    fit, rm, K = C.c_double(), C.c_double(), C.c_int64()
    iters = np.zeros(S, np.int32)
    stats = np.zeros((S, 8))
    rc = lib().orc_multiscale_gicp(_p(s), C.c_int64(s.shape[0]), _p(t), C.c_int64(t.shape[0]), C.c_int(S), _p(vs), _p(md),
                                   _p(mi, C.c_int32), _p(T0), C.byref(o), _p(T), C.byref(fit), C.byref(rm),
                                   _p(iters, C.c_int32), C.byref(K), _p(stats))
    _check(rc, "multiscale_gicp")
    return OracleResult(T.reshape(4, 4), fit.value, rm.value, iters.tolist(), K.value, stats=stats)


def evaluate_registration(source, target, max_correspondence_distance, transformation=None, *, want_corr=False, want_gtg=False):
    """o3d.pipelines.registration.evaluate_registration on the clouds as given (AF:809-822); optionally also the
    correspondences and the 6x6 GTG of get_information_matrix_from_point_clouds (AF:327-331) at the same pose."""
    s = _d(getattr(source, "points", source)).reshape(-1, 3)
    t = _d(getattr(target, "points", target)).reshape(-1, 3)
    T = _d(np.eye(4) if transformation is None else transformation).reshape(16)
    fit, rm, K = C.c_double(), C.c_double(), C.c_int64()
    corr = np.empty(s.shape[0], np.int32) if want_corr else None
    gtg = np.zeros(36) if want_gtg else None
    rc = lib().orc_evaluate_registration(_p(s), C.c_int64(s.shape[0]), _p(t), C.c_int64(t.shape[0]),
                                         C.c_double(max_correspondence_distance), _p(T), C.byref(fit), C.byref(rm), C.byref(K),
                                         _p(corr, C.c_int32) if want_corr else None, _p(gtg) if want_gtg else None)
    _check(rc, "evaluate_registration")
    r = OracleResult(T.reshape(4, 4).copy(), fit.value, rm.value, [], K.value)
    r.correspondence = corr
    r.information = gtg.reshape(6, 6) if want_gtg else None
    return r


def get_information_matrix_from_point_clouds(source, target, max_correspondence_distance, transformation):
    """o3d.pipelines.registration.get_information_matrix_from_point_clouds (AF:327-331, S3:317-320)"""
    return evaluate_registration(source, target, max_correspondence_distance, transformation, want_gtg=True).information


# ---- the two schedules of the reference (host float expressions reproduced verbatim) -------------
def create_scales_script2(n_scales):
    """2_MGICP_refinement_in_NCLT_dataset.py:102-106"""
    voxel_radius = 0.1
    voxel_radius = [voxel_radius + (0.1 * i) for i in range(n_scales)]
    voxel_radius.reverse()
    return voxel_radius


def max_correspondence_distances_script2(scales):
    """2_MGICP_refinement_in_NCLT_dataset.py:112-120"""
    n = len(scales)
    if n == 3:
        return [3 * scales[0], 2 * scales[1], scales[2]]
    if n == 4:
        return [3 * scales[0], 2.5 * scales[1], 2 * scales[2], scales[3]]
    if n == 5:
        return [3 * scales[0], 2.5 * scales[1], 2 * scales[2], 1.5 * scales[3], scales[4]]
    raise UnboundLocalError("max_correspondence_distances: the reference only defines 3, 4 or 5 scales")


def create_scales_all_functions(n_scales):
    """ALL_FUNCTIONS.py:260-264 (+ reverse at AF:275)"""
    v = [0.1]
    for _ in range(n_scales - 1):
        v.append(v[-1] + v[-1])
    v.reverse()
    return v


def radius_from_cloud_pair(s, t):
    """ALL_FUNCTIONS.py:1092-1101"""
    d1 = s.max(axis=0) - s.min(axis=0)
    d2 = t.max(axis=0) - t.min(axis=0)
    r1 = (d1[0] * d1[1] * d1[2]) ** (1 / 3)
    r2 = (d2[0] * d2[1] * d2[2]) ** (1 / 3)
    return (r1 + r2) / 2


def Multiscale_GICP(source, target, n_scales, itera_escala, T_ini, schedule="script2", **kw):
    s = _d(getattr(source, "points", source)).reshape(-1, 3)
    t = _d(getattr(target, "points", target)).reshape(-1, 3)
    if schedule == "script2":
        voxels = create_scales_script2(n_scales)
        dists = max_correspondence_distances_script2(voxels)
    elif schedule == "all_functions":
        voxels = create_scales_all_functions(n_scales)
        r = radius_from_cloud_pair(s, t)
        dists = [r * (2 ** (-i)) for i in range(n_scales)]
    else:
        raise ValueError(schedule)
    return multiscale_gicp(s, t, voxels, dists, [itera_escala] * n_scales, T_ini, **kw)


# ---- the stage before the refinement: registro_FGR (oracle/fgr_oracle.c) ---------------------------------
class _FgrOpts(C.Structure):
    _fields_ = [("division_factor", C.c_double), ("use_absolute_scale", C.c_int32), ("decrease_mu", C.c_int32),
                ("maximum_correspondence_distance", C.c_double), ("iteration_number", C.c_int32), ("tuple_scale", C.c_double),
                ("maximum_tuple_count", C.c_int32), ("seed", C.c_uint64)]


def estimate_normals_hybrid(xyz, radius, max_nn):
    """estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) -- ALL_FUNCTIONS.py:181-183"""
    p = _d(xyz).reshape(-1, 3)
    out = np.empty_like(p)
    _check(lib().orc_estimate_normals_hybrid(_p(p), C.c_int64(p.shape[0]), C.c_double(radius), C.c_int(max_nn), _p(out)),
           "estimate_normals_hybrid")
    return out


def compute_fpfh_feature(xyz, normals, radius, max_nn):
    """compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius, max_nn)) -> [n, 33] -- ALL_FUNCTIONS.py:185-187"""
    p, nr = _d(xyz).reshape(-1, 3), _d(normals).reshape(-1, 3)
    out = np.empty((p.shape[0], 33))
    _check(lib().orc_compute_fpfh(_p(p), _p(nr), C.c_int64(p.shape[0]), C.c_double(radius), C.c_int(max_nn), _p(out)), "compute_fpfh")
    return out


def registration_fgr_based_on_feature_matching(source, target, source_fpfh, target_fpfh, *, division_factor=1.4,
                                               use_absolute_scale=False, decrease_mu=False, maximum_correspondence_distance=0.025,
                                               iteration_number=64, tuple_scale=0.95, maximum_tuple_count=1000, seed=0, engine=False):
    """Open3D's call with its defaults; returns (T source->target, number of correspondences optimised)"""
    s, t = _d(source).reshape(-1, 3), _d(target).reshape(-1, 3)
    fs, ft = _d(source_fpfh).reshape(-1, 33), _d(target_fpfh).reshape(-1, 33)
    assert fs.shape[0] == s.shape[0] and ft.shape[0] == t.shape[0]
    o = _FgrOpts(division_factor, int(use_absolute_scale), int(decrease_mu), maximum_correspondence_distance, iteration_number,
                 tuple_scale, maximum_tuple_count, seed)
    T = np.empty(16)
    nc = C.c_int64()
    # engine: through csrc/fgr_math.cuh, the functions the kernels call; engine == 'kernel_order': sums reduced like k_fgr_pair
    fn = lib().orc_fgr_engine_kernel_order if engine == "kernel_order" else (lib().orc_fgr_engine if engine else lib().orc_fgr)
    _check(fn(_p(s), C.c_int64(s.shape[0]), _p(t), C.c_int64(t.shape[0]), _p(fs), _p(ft), C.byref(o), _p(T), C.byref(nc)), "fgr")
    return T.reshape(4, 4), int(nc.value)


def registro_FGR(source, target, voxel_size, seed=0):
    """ALL_FUNCTIONS.py:178-203 / 1_FGR_pairwise_registration_in_NCLT_dataset.py:41-66 with the reference's parameters;
    returns (T source->target, number of correspondences optimised)"""
    s, t = _d(source).reshape(-1, 3), _d(target).reshape(-1, 3)
    T = np.empty(16)
    nc = C.c_int64()
    _check(lib().orc_registro_fgr(_p(s), C.c_int64(s.shape[0]), _p(t), C.c_int64(t.shape[0]), C.c_double(voxel_size),
                                  C.c_uint64(seed), _p(T), C.byref(nc)), "registro_fgr")
    return T.reshape(4, 4), int(nc.value)


def hybrid_lists(xyz, radius, max_nn):
    """KDTreeFlann::SearchHybrid for every point of the cloud: (idx [n, max_nn], d2 [n, max_nn], cnt [n]); entry 0 is the point"""
    p = _d(xyz).reshape(-1, 3)
    idx, d2 = knn(p, p, max_nn)[:2]
    idx, d2 = np.ascontiguousarray(idx, dtype=np.int32), np.ascontiguousarray(d2, dtype=np.float64)
    cnt = (d2 < radius * radius).sum(axis=1).astype(np.int32)
    return idx, d2, cnt


def fpfh_engine(xyz, radius_normals, nn_normals, radius_fpfh, nn_fpfh):
    """normals and FPFH through the per-point functions of csrc/fpfh_math.cuh (the ones the CUDA kernels call), neighbour
    lists from the oracle's KD-tree -> (normals [n, 3], fpfh [n, 33])"""
    p = _d(xyz).reshape(-1, 3)
    i_n, _, c_n = hybrid_lists(p, radius_normals, nn_normals)
    i_f, d_f, c_f = hybrid_lists(p, radius_fpfh, nn_fpfh)
    nrm, f = np.empty_like(p), np.empty((p.shape[0], 33))
    _check(lib().orc_fpfh_engine(_p(p), C.c_int64(p.shape[0]), _p(i_n, C.c_int32), _p(c_n, C.c_int32), C.c_int(nn_normals),
                                 _p(i_f, C.c_int32), _p(d_f), _p(c_f, C.c_int32), C.c_int(nn_fpfh), _p(nrm), _p(f)), "fpfh_engine")
    return nrm, f


def check_recip_div(n=20_000_000, seed=1):
    """csrc/fpfh_math.cuh:div_by_recip(a, b, RN(1/b)) against a / b on n operand pairs (random, exact-quotient neighbourhoods,
    perturbed quotients): the number of pairs on which they differ (expected 0)."""
    L = lib()
    L.orc_check_recip_div.restype = C.c_int64
    L.orc_check_recip_div.argtypes = [C.c_int64, C.c_uint64]
    return int(L.orc_check_recip_div(int(n), int(seed)))
