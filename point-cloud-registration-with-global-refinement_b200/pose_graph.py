"""Pose-graph builder over a list of clouds: the reference's ``full_registration`` (ALL_FUNCTIONS.py:342-394), batched.

The reference registers every pair (source_id, target_id) with source_id < target_id <= source_id + k one after the other,
each time recomputing normals and FPFH descriptors of both clouds inside ``registro_FGR``.  Here the work is laid out for one
GPU: descriptors once per CLOUD, Fast Global Registration for all pairs in one batched call, Multiscale GICP for all pairs in
one batched call (per-pair ALL_FUNCTIONS schedule), information matrices for all pairs in one call; only the assembly of the
graph (odometry chaining, edge flags, success count at fitness > 0.40) is host work, in the reference's order.

Open3D's ``PoseGraph`` type is not available (Open3D is not installable offline); ``PoseGraph`` below carries the same fields
and reads / writes the JSON layout of ``o3d.io.write_pose_graph`` (column-major matrices), so that the file hand-off to
``o3d.io.read_pose_graph`` / ``global_optimization`` keeps working.  NOTE: the FGR registration stage this builds on has not
run on a GPU yet (csrc/mgicp_fgr.cuh); the pair enumeration, assembly and file format are covered by CPU tests.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

from .registration import (_points, create_scales, default_engine)


@dataclass
class PoseGraphNode:
    pose: np.ndarray


@dataclass
class PoseGraphEdge:
    source_node_id: int
    target_node_id: int
    transformation: np.ndarray
    information: np.ndarray
    uncertain: bool = False
    confidence: float = 1.0


@dataclass
class PoseGraph:
    nodes: list = field(default_factory=list)
    edges: list = field(default_factory=list)

    def to_json(self) -> dict:
        col = lambda m: [float(x) for x in np.asarray(m, np.float64).flatten(order="F")]
        return {"class_name": "PoseGraph",
                "edges": [{"class_name": "PoseGraphEdge", "confidence": float(e.confidence), "information": col(e.information),
                           "source_node_id": int(e.source_node_id), "target_node_id": int(e.target_node_id),
                           "transformation": col(e.transformation), "uncertain": bool(e.uncertain), "version_major": 1,
                           "version_minor": 0} for e in self.edges],
                "nodes": [{"class_name": "PoseGraphNode", "pose": col(nd.pose), "version_major": 1, "version_minor": 0} for nd in self.nodes],
                "version_major": 1, "version_minor": 0}

    @staticmethod
    def from_json(d: dict) -> "PoseGraph":
        if d.get("class_name") != "PoseGraph":
            raise ValueError("not a PoseGraph file")
        m = lambda a, n: np.asarray(a, np.float64).reshape(n, n, order="F")
        g = PoseGraph()
        g.nodes = [PoseGraphNode(m(nd["pose"], 4)) for nd in d.get("nodes", [])]
        g.edges = [PoseGraphEdge(int(e["source_node_id"]), int(e["target_node_id"]), m(e["transformation"], 4), m(e["information"], 6),
                                 bool(e["uncertain"]), float(e.get("confidence", 1.0))) for e in d.get("edges", [])]
        return g


def write_pose_graph(path: str, graph: PoseGraph) -> None:
    with open(path, "w") as f:
        json.dump(graph.to_json(), f, indent=4)


def read_pose_graph(path: str) -> PoseGraph:
    with open(path) as f:
        return PoseGraph.from_json(json.load(f))


def registration_pairs(n_clouds: int, k: int) -> list:
    """(source_id, target_id) in the reference's loop order: source_id < target_id <= source_id + k
    (k (n - k) + (k^2 - k) / 2 pairs for k <= n - 1)"""
    return [(s, t) for s in range(n_clouds) for t in range(s + 1, n_clouds) if t - s <= k]


def assemble_pose_graph(pairs, transformations, informations, fitness, verbose: bool = False):
    """the graph bookkeeping of full_registration given the pairwise results (in `registration_pairs` order): odometry chained
    over the consecutive pairs, node i + 1 = inverse of the accumulated odometry, every pair an edge (uncertain = not
    consecutive); returns (PoseGraph, number of pairs with fitness > 0.40)"""
    graph = PoseGraph()
    odometry = np.identity(4)
    graph.nodes.append(PoseGraphNode(odometry.copy()))
    ok = 0
    for (s, t), T, info, fit in zip(pairs, transformations, informations, fitness):
        T = np.asarray(T, np.float64)
        if t == s + 1:
            odometry = np.dot(T, odometry)
            graph.nodes.append(PoseGraphNode(np.linalg.inv(odometry)))
            graph.edges.append(PoseGraphEdge(s, t, T, np.asarray(info, np.float64), uncertain=False))
        else:
            graph.edges.append(PoseGraphEdge(s, t, T, np.asarray(info, np.float64), uncertain=True))
        ok += 1 if fit > 0.40 else 0
        if verbose:
            print(f"{'Odometric case' if t == s + 1 else 'Caso loopclosure'}: cloud {s} in cloud {t}: {'Sucesso' if fit > 0.40 else 'Falhou'}")
    return graph, ok


def pair_seed(seed: int, source_id: int, target_id: int) -> int:
    """FGR's tuple test is random (Open3D draws from a global engine); here the draw depends on the seed and on WHICH clouds are
    registered, not on the pair's position in the batch: the same pair gives the same pose in any batch."""
    return (int(seed) * 1000003 + int(source_id)) * 1000003 + int(target_id)


def register_pairs(clouds, pairs, voxel_size, *, engine=None, seed: int = 0, n_scales: int = 3, itera_escala: int = 100):
    """Coarse_to_fine_FGR_M_GICP (ALL_FUNCTIONS.py:315-332) for many pairs over a shared cloud list, batched:
    returns (transformations [B,4,4], informations [B,6,6], fitness [B], inlier_rmse [B])"""
    eng = engine or default_engine()
    pts = [_points(c) for c in clouds]
    B = len(pairs)
    # registro_FGR: descriptors once per cloud, then all pairs
    feats = eng.fpfh_clouds(pts, 2 * voxel_size, 20, 10 * voxel_size, 200, resident=True)
    caps = [int(int((len(pts[s]) + len(pts[t])) / 2) * 0.2) for s, t in pairs]
    T_fgr, _ = eng.fgr_pairs(pts, feats, pairs, division_factor=1.4, use_absolute_scale=True, decrease_mu=True,
                             maximum_correspondence_distance=2 * voxel_size, iteration_number=300, tuple_scale=0.95,
                             maximum_tuple_count=caps, seeds=[pair_seed(seed, s_, t_) for s_, t_ in pairs])
    # Multiscale_GICP, ALL_FUNCTIONS schedule: voxels 0.1 * 2^k reversed, distances from the pair's bounding boxes
    voxels = create_scales(n_scales)
    voxels.reverse()
    b = eng.cloud_bounds(pts)
    dif = b[:, 3:] - b[:, :3]
    rad = [(d[0] * d[1] * d[2]) ** (1 / 3) for d in dif]                     # radius_from_cloud_pair (ALL_FUNCTIONS.py:1092-1101)
    dists = np.asarray([[(rad[s] + rad[t]) / 2 * (2 ** (-i)) for i in range(n_scales)] for s, t in pairs])
    r = eng.run(pts, pairs, voxels, dists, itera_escala, T_fgr)
    ev = eng.evaluate_clouds(pts, pairs, [voxel_size] * B, r.transformation)
    return r.transformation, ev["information"], r.fitness, r.inlier_rmse


def full_registration(lista_nuvens, voxel_size, k, *, engine=None, seed: int = 0, verbose: bool = True) -> PoseGraph:
    """ALL_FUNCTIONS.py:342-394 with the same arguments; returns the pose graph (nodes: absolute poses registering cloud n in
    cloud 0; edges: relative poses with information matrices, uncertain = loop-closure edge)."""
    n = len(lista_nuvens)
    pairs = registration_pairs(n, k)
    if verbose:
        print(f"\nPara n={n} | k={k} serao feitos {len(pairs)} registros em pares\n")
    T, info, fit, _ = register_pairs(lista_nuvens, pairs, voxel_size, engine=engine, seed=seed)
    graph, ok = assemble_pose_graph(pairs, T, info, fit, verbose=verbose)
    if verbose and pairs:
        print(f"{ok} sucessos de {len(pairs)} registros em pares. Taxa: {ok / len(pairs)}")
    return graph
