"""The reference-facing interface: drop-ins for the reference's ``Multiscale_GICP``.

    Multiscale_GICP(source, target, n_scales, itera_escala, T_ini)
        /root/reference/ALL_FUNCTIONS.py:272-313               (schedule="all_functions")
        /root/reference/2_MGICP_refinement_in_NCLT_dataset.py:128-164   (schedule="script2", default:
        the shipped golden poses come from this variant)

Same positional arguments and meaning; ``source`` / ``target`` may be anything with a ``.points``
attribute (Open3D PointCloud) or an N x 3 array; the inputs are never modified (AF:289-290); the
return value carries ``.transformation`` (4x4, source -> target), ``.fitness`` and ``.inlier_rmse``
like Open3D's RegistrationResult (AF:313).  Errors follow Open3D: RuntimeError for a non-positive
voxel size or correspondence distance.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .engine import Engine

_default_engine = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


@dataclass
class RegistrationResult:
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    iterations: list = field(default_factory=list)
    num_correspondences: int = 0
    stats: np.ndarray | None = None
    correspondence_set: np.ndarray | None = None     # [K, 2] int32 (source index, target index) at the last scale (AF:1064)

    def __repr__(self):
        return (f"RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e}, and "
                f"correspondence_set size of {self.num_correspondences}\nAccess transformation to get result.")


# ---- schedules: the host float expressions of the reference, verbatim in effect -----------------------
def create_scales_script2(n_scales):
    """2_MGICP_refinement_in_NCLT_dataset.py:102-106 (0.1 + 0.1*2 == 0.30000000000000004 matters for voxel membership)"""
    voxel_radius = 0.1
    voxel_radius = [voxel_radius + (0.1 * i) for i in range(n_scales)]
    voxel_radius.reverse()
    return voxel_radius


def max_correspondence_distances(scales):
    """2_MGICP_refinement_in_NCLT_dataset.py:112-120; like the reference, only 3, 4 or 5 scales are defined"""
    n_scales = len(scales)
    if n_scales == 3:
        return [3 * scales[0], 2 * scales[1], scales[2]]
    if n_scales == 4:
        return [3 * scales[0], 2.5 * scales[1], 2 * scales[2], scales[3]]
    if n_scales == 5:
        return [3 * scales[0], 2.5 * scales[1], 2 * scales[2], 1.5 * scales[3], scales[4]]
    raise UnboundLocalError("cannot access local variable 'max_correspondence_distances' (the reference defines 3, 4 or 5 scales)")


def create_scales(n_scales):
    """ALL_FUNCTIONS.py:260-264"""
    voxel_radius = [0.1]
    for _ in range(n_scales - 1):
        voxel_radius.append(voxel_radius[-1] + voxel_radius[-1])
    return voxel_radius


def radius_from_cloud_pair(source, target, engine: Engine | None = None):
    """ALL_FUNCTIONS.py:1092-1101, bounds reduced on the GPU (kernel K0)."""
    eng = engine or default_engine()
    b = eng.cloud_bounds([_points(source), _points(target)])
    dif_1, dif_2 = b[0, 3:] - b[0, :3], b[1, 3:] - b[1, :3]
    rad_1 = (dif_1[0] * dif_1[1] * dif_1[2]) ** (1 / 3)
    rad_2 = (dif_2[0] * dif_2[1] * dif_2[2]) ** (1 / 3)
    return (rad_1 + rad_2) / 2


def _points(c):
    a = np.asarray(getattr(c, "points", c))
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("a cloud must be an N x 3 array or expose .points")
    return a


def multiscale_gicp(source, target, voxel_sizes, max_corr_dists, max_iters, T_init, *, sor_k=30, sor_std=1.0, normal_k=20,
                    epsilon=1e-3, loss="l1", loss_k=1.0, rel_fitness=1e-6, rel_rmse=1e-6, engine: Engine | None = None,
                    **tuning) -> RegistrationResult:
    """Explicit-schedule core: the body of Multiscale_GICP (AF:286-312) for one pair."""
    eng = engine or default_engine()
    opts = eng.make_opts(sor_k=sor_k, sor_std=sor_std, normal_k=normal_k, epsilon=epsilon, loss=loss, loss_k=loss_k,
                         rel_fitness=rel_fitness, rel_rmse=rel_rmse, **tuning)
    src = _points(source)
    r = eng.run([src, _points(target)], [(0, 1)], list(voxel_sizes), list(max_corr_dists), max_iters,
                np.asarray(T_init, np.float64).reshape(1, 4, 4), opts)
    # like Open3D's result, the correspondences of the last scale come along (indices into the clouds that scale registered:
    # Engine.get_stage(cloud, last_scale, STAGE_POINTS))
    corr = eng.correspondences(0, len(src))
    return RegistrationResult(r.transformation[0], float(r.fitness[0]), float(r.inlier_rmse[0]), r.iterations[0].tolist(),
                              int(r.num_correspondences[0]), r.stats[0], corr)


def multiscale_gicp_batch(clouds, pairs, voxel_sizes, max_corr_dists, max_iters, T_init, *, engine: Engine | None = None,
                          **kw):
    """Many pairs over a shared cloud list in one go -> engine.BatchResult ([B,4,4], [B], [B], ...).
    One batch through the streaming pipeline (stream.BatchStream: clouds packed straight into pinned staging memory, results
    and error flag brought home in one copy); callers with several batches should keep a BatchStream and feed it all of them,
    which overlaps packing, upload, compute and download across batches."""
    from .stream import BatchStream
    eng = engine or default_engine()
    bs = BatchStream(list(voxel_sizes), max_corr_dists, max_iters, engine=eng, engines=1, opts=eng.make_opts(**kw))
    try:
        return bs.run_one([_points(c) for c in clouds], list(pairs), T_init)
    finally:
        bs.close()


def Multiscale_GICP(source, target, n_scales, itera_escala, T_ini, schedule="script2", **kw) -> RegistrationResult:
    if schedule == "script2":
        voxel_sizes = create_scales_script2(n_scales)
        search_distances = max_correspondence_distances(voxel_sizes)
    elif schedule == "all_functions":
        voxel_sizes = create_scales(n_scales)
        voxel_sizes.reverse()
        max_correspondence_distance = radius_from_cloud_pair(source, target, kw.get("engine"))
        search_distances = [max_correspondence_distance * (2 ** (-i)) for i in range(n_scales)]
    else:
        raise ValueError(f"schedule must be 'script2' or 'all_functions', not {schedule!r}")
    return multiscale_gicp(source, target, voxel_sizes, search_distances, [itera_escala] * n_scales, T_ini, **kw)


# ---- the immediate consumers of the refined pose -----------------------------------------------
def evaluate_registration(source, target, max_correspondence_distance, transformation=None, *, engine: Engine | None = None):
    """o3d.pipelines.registration.evaluate_registration(source, target, max_correspondence_distance, transformation)
    as called by the reference (AF:809-822): fitness and inlier RMSE of a pose on the clouds as given."""
    eng = engine or default_engine()
    T = np.eye(4) if transformation is None else np.asarray(transformation, np.float64)
    r = eng.evaluate_clouds([_points(source), _points(target)], [(0, 1)], [max_correspondence_distance], T.reshape(1, 4, 4))
    return RegistrationResult(T.copy(), float(r["fitness"][0]), float(r["rmse"][0]), [], int(r["K"][0]))


def get_information_matrix_from_point_clouds(source, target, max_correspondence_distance, transformation, *,
                                             engine: Engine | None = None) -> np.ndarray:
    """o3d.pipelines.registration.get_information_matrix_from_point_clouds (AF:327-331, S3:317-320): the 6x6 GTG of the
    correspondences found at `transformation` within `max_correspondence_distance`."""
    eng = engine or default_engine()
    T = np.asarray(transformation, np.float64).reshape(1, 4, 4)
    r = eng.evaluate_clouds([_points(source), _points(target)], [(0, 1)], [max_correspondence_distance], T)
    return r["information"][0]


def calculate_RMSE_and_fitness(lista_nuvens, T_circuito, distancia, *, engine: Engine | None = None):
    """ALL_FUNCTIONS.py:801-824, all pairs of the circuit in one batched launch: cloud i+1 is evaluated against cloud i
    with T_circuito[i]; if there are as many poses as clouds the last one closes the loop (cloud 0 against the last)."""
    eng = engine or default_engine()
    n_nuvens, n_T = len(lista_nuvens), len(T_circuito)
    if n_nuvens == n_T:
        pairs = [(i + 1, i) for i in range(n_nuvens - 1)] + [(0, n_nuvens - 1)]
    elif n_nuvens - 1 == n_T:
        pairs = [(i + 1, i) for i in range(n_T)]
    else:
        print("The number of clouds and poses are inconsistent")       # the reference prints and returns two empty lists
        return [], []
    r = eng.evaluate_clouds([_points(c) for c in lista_nuvens], pairs, distancia, np.stack([np.asarray(T, np.float64) for T in T_circuito]))
    return r["rmse"].tolist(), r["fitness"].tolist()


def Coarse_to_fine_M_GICP(source, target, voxel_size, T_ini, *, n_scales=3, itera_escala=100, schedule="all_functions",
                          engine: Engine | None = None, **kw):
    """The refinement half of Coarse_to_fine_FGR_M_GICP (AF:316-332): Multiscale_GICP from a given coarse pose, then the
    information matrix of the refined pose at `voxel_size`.  (The FGR front end that produces T_ini: registro_FGR below.)"""
    result = Multiscale_GICP(source, target, n_scales, itera_escala, T_ini, schedule=schedule, engine=engine, **kw)
    info = get_information_matrix_from_point_clouds(source, target, voxel_size, result.transformation, engine=engine)
    return result, info


def registro_FGR(source, target, voxel_size, *, engine: Engine | None = None, seed: int = 0):
    """ALL_FUNCTIONS.py:178-203 == 1_FGR_pairwise_registration_in_NCLT_dataset.py:41-66: hybrid-radius normals (2 v, 20 nn), FPFH
    (10 v, 200 nn) and Fast Global Registration with the reference's option values; returns a RegistrationResult whose
    fitness / inlier_rmse come from evaluate_registration at 2 v, like Open3D's."""
    eng = engine or default_engine()
    src, tgt = _points(source), _points(target)
    feats = eng.fpfh_clouds([src, tgt], 2 * voxel_size, 20, 10 * voxel_size, 200, resident=True)
    n_pontos = int((len(src) + len(tgt)) / 2)
    T, nc = eng.fgr_pairs([src, tgt], feats, [(0, 1)], division_factor=1.4, use_absolute_scale=True, decrease_mu=True,
                          maximum_correspondence_distance=2 * voxel_size, iteration_number=300, tuple_scale=0.95,
                          maximum_tuple_count=int(n_pontos * 0.2), seeds=[seed])   # int(n_pontos * 0.2): AF:196
    ev = evaluate_registration(src, tgt, 2 * voxel_size, T[0], engine=eng)
    return RegistrationResult(T[0], float(ev.fitness), float(ev.inlier_rmse), [], ev.num_correspondences, np.asarray([nc[0]]))


def Coarse_to_fine_FGR_M_GICP(source, target, voxel_size, *, engine: Engine | None = None, seed: int = 0, **kw):
    """ALL_FUNCTIONS.py:315-332: registro_FGR, then Multiscale_GICP (ALL_FUNCTIONS schedule, 3 scales, 100 iterations per
    scale) from the FGR pose, then the information matrix of the refined pose at `voxel_size`.
    Returns (result_M_GICP, information_matrix)."""
    eng = engine or default_engine()
    result_FGR = registro_FGR(source, target, voxel_size, engine=eng, seed=seed)
    return Coarse_to_fine_M_GICP(source, target, voxel_size, result_FGR.transformation, n_scales=3, itera_escala=100,
                                 schedule="all_functions", engine=eng, **kw)

