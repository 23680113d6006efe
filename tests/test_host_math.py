"""The engine's per-point math header (csrc/mgicp_math.cuh) compiled for the host and probed on the CPU: closed forms
against numpy / the faithful oracle.  No GPU."""
import numpy as np


def test_fast_eigen_bit_identical_to_oracle(oracle):
    rng = np.random.default_rng(0)
    for _ in range(500):
        A = rng.normal(size=(3, 3)) * rng.uniform(0.01, 3, size=(1, 3))
        C = A @ A.T
        c6 = [C[0, 0], C[0, 1], C[0, 2], C[1, 1], C[1, 2], C[2, 2]]
        assert np.array_equal(oracle.probe("fast_eigen3x3", c6, out_len=3), oracle.fast_eigen3x3(c6))


def test_closed_form_weight_matrix_is_inverse_sqrt(oracle):
    """W = (C_t + C_s)^(-1/2) for C = I - (1-eps) m m^T, including near-parallel and near-antiparallel normals"""
    rng = np.random.default_rng(1)
    k = 1 - 1e-3
    worst = 0.0
    for t in range(2000):
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        if t % 4 == 0:
            b = a + 1e-7 * rng.normal(size=3)
        elif t % 4 == 1:
            b = -a + 1e-5 * rng.normal(size=3)
        else:
            b = rng.normal(size=3)
        b /= np.linalg.norm(b)
        W6 = oracle.probe("weight_matrix", a, b, k, out_len=6)
        W = np.array([[W6[0], W6[1], W6[2]], [W6[1], W6[3], W6[4]], [W6[2], W6[4], W6[5]]])
        M = 2 * np.eye(3) - k * (np.outer(a, a) + np.outer(b, b))
        w, V = np.linalg.eigh(M)
        ref = V @ np.diag(w ** -0.5) @ V.T
        worst = max(worst, np.abs(W - ref).max())
    # near-parallel normals: the smallest eigenvalue of M is 2*eps = 0.002, formed as 2 - 1.998 (3 digits cancel) in ANY
    # implementation that adds the two covariances, Open3D's included; W's entries (<= 22.4) then carry ~1e-12
    assert worst < 2e-11, worst


def test_ldlt_and_pose_update_match_oracle(oracle):
    rng = np.random.default_rng(2)
    for _ in range(100):
        A = rng.normal(size=(6, 6))
        A = A @ A.T + 0.01 * np.eye(6)
        b = rng.normal(size=6)
        sums = np.concatenate([A[np.triu_indices(6)], b])
        x = oracle.probe("ldlt_solve6", sums, out_len=6)
        assert np.array_equal(x, oracle.ldlt_solve6(A, -b))          # same pivoted LDL^T, same operation order
        assert np.allclose(x, np.linalg.solve(A, -b), rtol=1e-8)
        T = oracle.probe("vec6_to_mat4", 0.05 * x, out_len=16).reshape(4, 4)
        assert np.allclose(T, oracle.vec6_to_mat4(0.05 * x), atol=1e-16)


def test_deterministic_trig_same_on_both_sides(oracle):
    rng = np.random.default_rng(3)
    for x in rng.uniform(-3.5, 3.5, 2000):
        s, c, a = oracle.det_trig(float(x))
        assert np.array_equal(oracle.probe("trig", float(x), out_len=3), [s, c, a])
