"""Pose-graph bookkeeping of full_registration (ALL_FUNCTIONS.py:342-394): pair enumeration, graph assembly, the JSON layout of
o3d.io.write_pose_graph.  CPU only; the registrations themselves are GPU work (tests/test_gpu_fgr.py, test_gpu_parity.py)."""
import json

import numpy as np
import pytest

from mgicp_b200 import pose_graph as pg


def _rigid(rng):
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    th = 0.3 * rng.normal()
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    T[:3, 3] = rng.normal(size=3)
    return T


@pytest.mark.parametrize("n,k", [(2, 1), (5, 1), (5, 3), (8, 7), (6, 10), (1, 3)])
def test_pair_enumeration_matches_the_reference_loops(n, k):
    want = []
    for source_id in range(n):                                   # the reference's loops, literally
        for target_id in range(source_id + 1, n):
            if target_id == source_id + 1:
                want.append((source_id, target_id))
            elif target_id != source_id + 1 and target_id - source_id <= k:
                want.append((source_id, target_id))
    assert pg.registration_pairs(n, k) == want
    if 1 <= k <= n - 1:
        assert len(want) == k * (n - k) + (k ** 2 - k) / 2       # the count the reference prints


def test_graph_assembly_follows_the_reference():
    rng = np.random.default_rng(0)
    n, k = 5, 2
    pairs = pg.registration_pairs(n, k)
    Ts = [_rigid(rng) for _ in pairs]
    infos = [np.diag(rng.uniform(1, 2, size=6)) for _ in pairs]
    fit = [0.9, 0.3, 0.41, 0.40, 0.8, 0.2, 0.95]
    g, ok = pg.assemble_pose_graph(pairs, Ts, infos, fit)
    assert ok == sum(f > 0.40 for f in fit) == 4
    assert len(g.nodes) == n and len(g.edges) == len(pairs)
    odo = np.eye(4)
    node = 1
    for (s, t), T, e in zip(pairs, Ts, g.edges):
        assert (e.source_node_id, e.target_node_id) == (s, t) and e.uncertain == (t != s + 1)
        assert np.array_equal(e.transformation, T)
        if t == s + 1:
            odo = T @ odo
            assert np.allclose(g.nodes[node].pose, np.linalg.inv(odo), atol=1e-14)
            node += 1
    assert np.array_equal(g.nodes[0].pose, np.eye(4))


def test_pose_graph_json_layout_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    g = pg.PoseGraph([pg.PoseGraphNode(np.eye(4)), pg.PoseGraphNode(_rigid(rng))],
                     [pg.PoseGraphEdge(0, 1, _rigid(rng), rng.normal(size=(6, 6)), uncertain=True)])
    p = tmp_path / "graph.json"
    pg.write_pose_graph(str(p), g)
    d = json.load(open(p))
    # Open3D's IJsonConvertible layout: class names, versions, Eigen matrices flattened column-major
    assert d["class_name"] == "PoseGraph" and d["version_major"] == 1 and d["version_minor"] == 0
    e = d["edges"][0]
    assert e["class_name"] == "PoseGraphEdge" and e["uncertain"] is True and e["confidence"] == 1.0
    assert len(e["transformation"]) == 16 and len(e["information"]) == 36 and len(d["nodes"][1]["pose"]) == 16
    T = g.edges[0].transformation
    assert e["transformation"][1] == T[1, 0] and e["transformation"][4] == T[0, 1] and e["transformation"][12] == T[0, 3]
    back = pg.read_pose_graph(str(p))
    assert np.array_equal(back.edges[0].transformation, T) and np.array_equal(back.edges[0].information, g.edges[0].information)
    assert np.array_equal(back.nodes[1].pose, g.nodes[1].pose) and back.edges[0].uncertain
    with pytest.raises(ValueError):
        pg.PoseGraph.from_json({"class_name": "Other"})
