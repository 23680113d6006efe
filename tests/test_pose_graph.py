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


class _FakeEngine:
    """records what the batched pipeline asks of the engine and answers with canned results (no GPU)"""

    def __init__(self, clouds):
        self.calls = []
        self.bounds = np.asarray([np.concatenate([c.min(axis=0), c.max(axis=0)]) for c in clouds])

    def fpfh_clouds(self, clouds, rn, kn, rf, kf, resident=False):
        self.calls.append(("fpfh", len(clouds), rn, kn, rf, kf))
        assert resident                                                  # the descriptors stay on the device between the stages
        return ("resident", [np.full((len(c), 33), float(i)) for i, c in enumerate(clouds)])

    def fgr_pairs(self, clouds, feats, pairs, **kw):
        self.calls.append(("fgr", list(pairs), kw))
        assert feats[0] == "resident"
        feats = feats[1]
        assert all(f.shape == (len(c), 33) for f, c in zip(feats, clouds))
        T = np.stack([np.eye(4)] * len(pairs))
        T[:, 0, 3] = np.arange(len(pairs)) + 1.0
        return T, np.zeros(len(pairs), np.int32)

    def cloud_bounds(self, clouds):
        return self.bounds

    def run(self, clouds, pairs, voxels, dists, iters, T0):
        self.calls.append(("run", list(pairs), list(voxels), np.asarray(dists), iters, np.asarray(T0)))
        from mgicp_b200.engine import BatchResult
        B = len(pairs)
        T = np.asarray(T0).copy()
        T[:, 1, 3] = 0.5
        return BatchResult(T, np.linspace(0.3, 0.9, B), np.full(B, 0.05), np.zeros((B, 3), np.int32), np.zeros(B, np.int32), np.zeros((B, 3, 8)))

    def evaluate_clouds(self, clouds, pairs, max_dists, T, want_corr=False):
        self.calls.append(("eval", list(pairs), list(max_dists), np.asarray(T)))
        return dict(information=np.stack([np.eye(6) * (b + 1) for b in range(len(pairs))]))


def test_full_registration_orchestration_matches_the_reference_parameters():
    """what full_registration -> Coarse_to_fine_FGR_M_GICP -> registro_FGR / Multiscale_GICP pass to Open3D, per pair
    (ALL_FUNCTIONS.py:178-203, 260-278, 315-332, 1092-1101), arrives at the batched engine calls"""
    rng = np.random.default_rng(3)
    clouds = [rng.uniform(-1, 1, size=(n, 3)) * (10 + i) for i, n in enumerate((101, 140, 77, 120))]
    eng = _FakeEngine(clouds)
    v, k = 0.1, 2
    g = pg.full_registration(clouds, v, k, engine=eng, seed=5, verbose=False)
    pairs = pg.registration_pairs(len(clouds), k)
    kinds = [c[0] for c in eng.calls]
    assert kinds == ["fpfh", "fgr", "run", "eval"]                       # one batched call per stage
    assert eng.calls[0][1:] == (4, 2 * v, 20, 10 * v, 200)               # descriptors once per cloud, the reference's radii
    _, fgr_pairs, kw = eng.calls[1]
    assert fgr_pairs == pairs
    assert kw["division_factor"] == 1.4 and kw["use_absolute_scale"] is True and kw["decrease_mu"] is True
    assert kw["maximum_correspondence_distance"] == 2 * v and kw["iteration_number"] == 300 and kw["tuple_scale"] == 0.95
    assert kw["maximum_tuple_count"] == [int(int((len(clouds[s]) + len(clouds[t])) / 2) * 0.2) for s, t in pairs]
    assert kw["seeds"] == [pg.pair_seed(5, s, t) for s, t in pairs]     # a function of the seed and the pair, not of the batch position
    assert len(set(kw["seeds"])) == len(pairs)
    _, run_pairs, voxels, dists, iters, T0 = eng.calls[2]
    assert run_pairs == pairs and voxels == [0.4, 0.2, 0.1] and iters == 100
    assert np.array_equal(T0[:, 0, 3], np.arange(len(pairs)) + 1.0)      # the FGR poses are the initial transforms
    for (s, t), d in zip(pairs, dists):
        dif_1, dif_2 = clouds[s].max(0) - clouds[s].min(0), clouds[t].max(0) - clouds[t].min(0)
        r = ((dif_1[0] * dif_1[1] * dif_1[2]) ** (1 / 3) + (dif_2[0] * dif_2[1] * dif_2[2]) ** (1 / 3)) / 2
        assert d.tolist() == [r * (2 ** (-i)) for i in range(3)]
    _, ev_pairs, md, T = eng.calls[3]
    assert ev_pairs == pairs and md == [v] * len(pairs) and np.array_equal(T[:, 1, 3], np.full(len(pairs), 0.5))
    assert len(g.nodes) == len(clouds) and len(g.edges) == len(pairs)
    assert [e.uncertain for e in g.edges] == [t != s + 1 for s, t in pairs]
    assert np.array_equal(g.edges[2].information, np.eye(6) * 3)
