"""SLERP / LUM global refinement mirrors against goldens produced by the reference's own functions
(tests/golden/make_refinement_goldens.py) and against size-independent properties."""
import os

import numpy as np
import pytest

import mgicp_b200 as m
from mgicp_b200 import global_refinement as gr
from mgicp_b200 import poses

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "refinement_goldens.npz"))
CIRCUITS = [c for c in range(5)]


def _rot_err(A, B):
    return max(float(np.abs(a[:3, :3] - b[:3, :3]).max()) for a, b in zip(A, B))


def _tr_err(A, B):
    return max(float(np.abs(a[:3, 3] - b[:3, 3]).max()) for a, b in zip(A, B))


@pytest.mark.parametrize("c", CIRCUITS)
def test_lum_matches_reference_dense_solve(c):
    """closed-form O(n) adjustment == the reference's inv(A'A) A'Lb (pure numpy in the reference: a true pin)"""
    rel = G[f"c{c}_in"]
    got = gr.reconstruir_Ts_para_origem_LUM(rel)
    want = G[f"c{c}_lum"]
    assert len(got) == len(want) == len(rel)
    assert np.array_equal(got[0], np.identity(4))
    assert _rot_err(got, want) < 1e-13
    assert _tr_err(got, want) < 1e-9 * max(1.0, float(np.abs(want[:, :3, 3]).max()))   # dense 3(n-1)-square inverse on the other side


@pytest.mark.parametrize("c", CIRCUITS)
def test_weighted_lum_matches_reference(c):
    rel, w = G[f"c{c}_in"], G[f"c{c}_w"]
    got = gr.reconstruir_Ts_para_origem_LUM(rel, list(w))
    want = G[f"c{c}_lum_w"]
    assert _rot_err(got, want) < 1e-13
    assert _tr_err(got, want) < 1e-9 * max(1.0, float(np.abs(want[:, :3, 3]).max()))


@pytest.mark.parametrize("c", CIRCUITS)
def test_slerp_orchestration_matches_reference(c):
    rel = G[f"c{c}_in"]
    for got, key in ((gr.reconstruir_Ts_para_origem_SLERP(rel), "slerp"), (gr.reconstruir_Ts_para_origem_SLERP_LUM(rel), "slerp_lum"),
                     (gr.reconstruir_Ts_para_origem_SLERP_LUM(rel, list(G[f"c{c}_w"])), "slerp_lum_w")):
        want = G[f"c{c}_{key}"]
        assert len(got) == len(want)
        assert _rot_err(got, want) < 1e-12, key
        assert _tr_err(got, want) < 1e-9 * max(1.0, float(np.abs(want[:, :3, 3]).max())), key
    # ALL_FUNCTIONS.py's variant composes the same quaternions in a different order: same poses to rounding
    assert _rot_err(gr.reconstruir_Ts_para_origem_SLERP(rel), G[f"c{c}_slerp_af"]) < 1e-11
    assert _tr_err(gr.reconstruir_Ts_para_origem_SLERP(rel), G[f"c{c}_slerp_af"]) < 1e-9


def test_pair_functions():
    A, B = G["pair_in"]
    assert np.abs(gr.Invert_pose(A) - G["pair_inv"]).max() < 2e-15
    assert np.abs(gr.Acumulate_Two_Poses(A, B) - G["pair_acc"]).max() < 2e-15
    for t, want in zip((0.0, 0.25, 0.5, 1.0), G["pair_interp"]):
        assert np.abs(gr.interpolar_duas_T(A, B, t) - want).max() < 1e-13
    assert np.abs(gr.interpolar_duas_T(A, B, 0.0) - A).max() < 1e-14
    assert np.abs(gr.interpolar_duas_T(A, B, 1.0) - B).max() < 1e-14


def _random_unit(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return gr.Quaternion(*q)


def test_quaternion_primitives():
    rng = np.random.default_rng(3)
    for _ in range(200):
        q, p = _random_unit(rng), _random_unit(rng)
        R, P = gr.as_rotation_matrix(q), gr.as_rotation_matrix(p)
        assert np.abs(R @ R.T - np.identity(3)).max() < 1e-14 and abs(np.linalg.det(R) - 1.0) < 1e-14
        assert np.abs(gr.as_rotation_matrix(q * p) - R @ P).max() < 1e-14                    # Hamilton product = matrix product
        assert np.abs(gr.as_rotation_matrix(q ** (-1)) - R.T).max() < 1e-14
        assert np.abs(gr.as_rotation_matrix(q * 3.0) - R).max() < 1e-14                       # non-unit input is normalised
        for nonorth in (True, False):
            b = gr.from_rotation_matrix(R, nonorthogonal=nonorth)
            assert b.w >= 0.0
            s = 1.0 if q.w >= 0 else -1.0
            assert np.abs(b.components - s * q.components).max() < 1e-12
        # slerp: end points, short arc whatever the sign of the second argument, constant angular velocity
        assert np.abs(gr.as_rotation_matrix(gr.slerp(q, p, 0, 1, 0.0)) - R).max() < 1e-13
        assert np.abs(gr.as_rotation_matrix(gr.slerp(q, p, 0, 1, 1.0)) - P).max() < 1e-13
        a, b = gr.slerp(q, p, 0, 1, 0.3), gr.slerp(q, -p, 0, 1, 0.3)
        assert np.abs(gr.as_rotation_matrix(a) - gr.as_rotation_matrix(b)).max() < 1e-13
        half = gr.as_rotation_matrix(gr.slerp(q, p, 2.0, 4.0, 3.0))
        assert np.abs(half @ R.T @ half - P).max() < 1e-12                                    # (midpoint relative rotation)^2 = whole


def test_from_rotation_matrix_of_text_rounded_rotation():
    """rotations read from `%.10f` text are orthonormal to 1e-10 only: the optimal quaternion is still within 1e-10"""
    rng = np.random.default_rng(5)
    q = _random_unit(rng)
    R = np.round(gr.as_rotation_matrix(q), 10)
    b = gr.from_rotation_matrix(R)
    s = 1.0 if q.w >= 0 else -1.0
    assert np.abs(b.components - s * q.components).max() < 2e-10


@pytest.mark.parametrize("n", [2, 5, 64, 901])
def test_closed_circuit_is_a_fixed_point(n):
    """a circuit that closes exactly is left as composed, by every method, at the reference's full circuit length (901)"""
    rng = np.random.default_rng(n)
    absolute = [np.identity(4)]
    for k in range(1, n):
        T = np.identity(4)
        T[:3, :3] = gr.as_rotation_matrix(_random_unit(rng))
        T[:3, 3] = rng.normal(size=3) * 5
        absolute.append(T)
    # relative poses in the reference's convention: abs[k+1] = Acumulate_Two_Poses(rel[k], abs[k])
    rel = []
    for k in range(n):
        a, b = absolute[k], absolute[(k + 1) % n]
        T = np.identity(4)
        T[:3, :3] = b[:3, :3] @ a[:3, :3].T
        T[:3, 3] = a[:3, :3].T @ (b[:3, 3] - a[:3, 3])
        rel.append(T)
    composed = poses.relative_to_absolute_poses(rel)
    assert max(np.abs(c - a).max() for c, a in zip(composed, absolute)) < 1e-9
    for f in (gr.reconstruir_Ts_para_origem_LUM, gr.reconstruir_Ts_para_origem_SLERP, gr.reconstruir_Ts_para_origem_SLERP_LUM):
        out = f(rel)
        assert len(out) == n
        assert max(np.abs(o - a).max() for o, a in zip(out, absolute)) < 1e-8, f.__name__


def test_lum_distributes_the_closure_and_is_a_least_squares_solution():
    rng = np.random.default_rng(11)
    n = 30
    L = rng.normal(size=(n, 3))
    w = rng.uniform(0.2, 3.0, size=n)
    for weights in (None, w):
        X = gr.lum_translations(L.reshape(-1, 1), weights)
        Xp = np.vstack([np.zeros(3), X, np.zeros(3)])
        res = L - (Xp[1:] - Xp[:-1])                               # l_i - (x_i - x_{i-1}), with x_{-1} = x_{n-1} = 0
        ww = np.ones(n) if weights is None else weights
        # normal equations: w_i res_i equal for all i (the Lagrange multiplier of the closure constraint)
        assert np.abs(res * ww[:, None] - (res * ww[:, None])[0]).max() < 1e-12
        assert np.abs(res.sum(axis=0) - L.sum(axis=0)).max() < 1e-12


def test_shapes_and_errors():
    with pytest.raises(ValueError):
        gr.reconstruir_Ts_para_origem_LUM([np.zeros((3, 3))])
    with pytest.raises(ValueError):
        gr.from_rotation_matrix(np.identity(4))
    one = gr.reconstruir_Ts_para_origem_LUM([np.identity(4)])
    assert len(one) == 1 and np.array_equal(one[0], np.identity(4))
    assert m.global_refinement is gr


def test_quaternion_layer_against_scipy(pkg):
    """The numpy-quaternion operations restated in global_refinement.py (the package is not installable offline), pinned against
    an INDEPENDENT implementation: scipy's Rotation / Slerp.  Conventions checked: component order (w, x, y, z), the rotation a
    quaternion stands for (as_rotation_matrix(from_rotation_matrix(R)) = R, not its transpose), the w >= 0 sign, the optimal
    quaternion of a slightly non-orthonormal matrix (rotations read from %.10f text, S3:22-40), slerp's direction,
    parametrisation and short-arc rule."""
    from scipy.spatial.transform import Rotation, Slerp
    gr = pkg.global_refinement
    rng = np.random.default_rng(11)
    for trial in range(40):
        rot = Rotation.from_rotvec(rng.normal(size=3) * rng.uniform(0.01, 3.0))
        R = rot.as_matrix()
        q = gr.from_rotation_matrix(R)
        xyzw = rot.as_quat()
        ref = np.array([xyzw[3], xyzw[0], xyzw[1], xyzw[2]])
        ref = -ref if ref[0] < 0 else ref
        assert q.w >= 0 and np.allclose(q.components, ref, atol=1e-12)
        assert np.allclose(gr.as_rotation_matrix(q), R, atol=1e-12)
        # a rotation as the reference reads it from %.10f text: the nearest rotation (polar decomposition) is what the optimal
        # quaternion stands for, to the order of the non-orthonormality squared
        R10 = np.array([[float(f"{v:.10f}") for v in row] for row in R])
        U, _, Vt = np.linalg.svd(R10)
        nearest = U @ Vt
        q10 = gr.from_rotation_matrix(R10)
        assert abs(np.sqrt(q10.norm2()) - 1.0) < 1e-12
        assert np.allclose(gr.as_rotation_matrix(q10), nearest, atol=5e-10)
        # non-unit quaternions still give the rotation (division by |q|^2)
        assert np.allclose(gr.as_rotation_matrix(gr.Quaternion(*(2.5 * q.components))), R, atol=1e-12)
        # slerp between two rotations at an arbitrary time, against scipy's Slerp
        rot2 = Rotation.from_rotvec(rng.normal(size=3) * rng.uniform(0.01, 3.0))
        t1, t2 = rng.uniform(-2, 0), rng.uniform(1, 3)
        t = rng.uniform(t1, t2)
        got = gr.slerp(q, gr.from_rotation_matrix(rot2.as_matrix()), t1, t2, t)
        want = Slerp([t1, t2], Rotation.concatenate([rot, rot2]))(t).as_matrix()
        assert np.allclose(gr.as_rotation_matrix(got), want, atol=1e-10)
    # short arc: interpolating towards -q2 is the same rotation path
    q1, q2 = gr.from_rotation_matrix(Rotation.from_rotvec([0.1, 0.2, 0.3]).as_matrix()), gr.from_rotation_matrix(Rotation.from_rotvec([-0.4, 0.1, 2.9]).as_matrix())
    a, b = gr.slerp(q1, q2, 0.0, 1.0, 0.3), gr.slerp(q1, -q2, 0.0, 1.0, 0.3)
    assert np.allclose(gr.as_rotation_matrix(a), gr.as_rotation_matrix(b), atol=1e-12)
    # composition: the product of quaternions is the product of the rotation matrices, in the same order
    A, B = Rotation.from_rotvec([0.3, -0.2, 0.5]), Rotation.from_rotvec([-0.7, 0.4, 0.1])
    qa, qb = gr.from_rotation_matrix(A.as_matrix()), gr.from_rotation_matrix(B.as_matrix())
    assert np.allclose(gr.as_rotation_matrix(qa * qb), A.as_matrix() @ B.as_matrix(), atol=1e-12)
    assert np.allclose(gr.as_rotation_matrix(qa / qb), A.as_matrix() @ B.as_matrix().T, atol=1e-12)
    assert np.allclose(gr.as_rotation_matrix(qa ** 0.5) @ gr.as_rotation_matrix(qa ** 0.5), A.as_matrix(), atol=1e-12)
