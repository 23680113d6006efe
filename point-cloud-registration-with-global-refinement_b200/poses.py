"""Pose bookkeeping either side of the refinement stage (the quaternion-free part).

The reference chains the refined relative poses T_{i+1 -> i} of a circuit into absolute poses, measures the loop-closure
error and compares pose lists (2_MGICP_refinement_in_NCLT_dataset.py:43-96; ALL_FUNCTIONS.py:110-150, 476-533, 831-838,
967-982).  These are n x 4 x 4 host-side operations (microseconds; there is nothing to put on a GPU) written here as
batched numpy; the conventions are the reference's own -- including ``compor_duas_poses``' R21 @ R10 rotation order --
and are pinned (to 2e-15: numpy's small matrix products vary in the last bit with operand alignment) by golden vectors
generated from the reference's functions (tests/golden/make_pose_goldens.py).
The SLERP / LUM global refinement itself (quaternion interpolation, weighted least squares) lives in global_refinement.py.
"""
from __future__ import annotations

import numpy as np


def _as_stack(poses) -> np.ndarray:
    a = np.asarray(poses, dtype=np.float64)
    if a.ndim != 3 or a.shape[1:] != (4, 4):
        raise ValueError("expected a list of 4 x 4 poses")
    return a


def _assemble(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    T = np.zeros(R.shape[:-2] + (4, 4))
    T[..., :3, :3] = R
    T[..., :3, 3] = t
    T[..., 3, 3] = 1.0
    return T


def Transformar_de_volta(T_4x4):
    """ALL_FUNCTIONS.py:110-114: inverse of a rigid pose, [R^T | -R^T t]."""
    T = np.asarray(T_4x4, dtype=np.float64)
    Rt = T[:3, :3].T
    return _assemble(Rt, -Rt @ T[:3, 3])


def compor_duas_poses(T21, T10):
    """ALL_FUNCTIONS.py:142-147: R20 = R21 @ R10, t20 = R10 @ t21 + t10 (the reference's convention, kept as is)."""
    T21, T10 = np.asarray(T21, dtype=np.float64), np.asarray(T10, dtype=np.float64)
    return _assemble(T21[:3, :3] @ T10[:3, :3], T10[:3, :3] @ T21[:3, 3] + T10[:3, 3])


def _rotations_to_origin(T: np.ndarray) -> np.ndarray:
    """out[k] = R_k R_{k-1} ... R_1 R_0 for k = 0..n-1 (R_j = rotation block of T[j]), associated left to right exactly
    like the reference's inner loop, which starts from the identity and multiplies with j descending:
    ((((I R_k) R_{k-1}) ...) R_0)."""
    n = T.shape[0]
    out = np.empty((n, 3, 3))
    for k in range(n):
        acc = np.identity(3)
        for j in range(k, -1, -1):
            acc = acc @ T[j, :3, :3]
        out[k] = acc
    return out


def relative_to_absolute_poses(T_circuito):
    """2_MGICP_refinement_in_NCLT_dataset.py:46-72 == ALL_FUNCTIONS.py:503-530 (poses_relativas_para_absolutas):
    relative poses T10, T21, ..., Tn_n-1 -> absolute poses [I, T10, T20, ..., T(n-1)_0]; the last composed pose (the
    closure) is dropped, the identity is inserted in front."""
    T = _as_stack(T_circuito)
    n = T.shape[0]
    R = _rotations_to_origin(T)
    t = np.empty((n, 3))
    t[0] = T[0, :3, 3]
    for i in range(n - 1):
        t[i + 1] = R[i] @ T[i + 1, :3, 3] + t[i]
    poses = [np.identity(4)] + [_assemble(R[i], t[i]) for i in range(n)]
    del poses[-1]
    return poses


poses_relativas_para_absolutas = relative_to_absolute_poses


def Calcular_Erro_LoopClosure(T_circuito, verbose: bool = False):
    """ALL_FUNCTIONS.py:476-497: 3 x 4 closure pose [R_loop | t_loop] of a closed circuit of relative poses (the
    reference also prints it and its Frobenius distance to the identity; printing is optional here)."""
    T = _as_stack(T_circuito)
    n = T.shape[0]
    R = _rotations_to_origin(T)
    t = T[0, :3, 3].copy()
    for i in range(n - 1):
        t = t + R[i] @ T[i + 1, :3, 3]
    closure = np.hstack((R[n - 1], t[:, None]))
    if verbose:
        print(f"POSE Closure error:\n{closure}")
        print(f"Distancia (Frobenious) para a identidade:\n{np.linalg.norm(R[n - 1] - np.identity(3), 'fro')}")
    return closure


def poses_absolutas_para_relativas(poses_absolutas):
    """ALL_FUNCTIONS.py:831-838: n absolute poses (the first the identity) -> n-1 relative poses
    compor_duas_poses(abs[i+1], inverse(abs[i]))."""
    _as_stack(poses_absolutas)                                    # shape check only
    P = [np.asarray(T, dtype=np.float64) for T in poses_absolutas]
    return [compor_duas_poses(P[i + 1], Transformar_de_volta(P[i])) for i in range(len(P) - 1)]


def subtract_squared_poses(list_poses_1, list_poses_2):
    """ALL_FUNCTIONS.py:967-982 / 2_MGICP...py:78-96: per pose, the Frobenius distance of the rotation blocks and the
    Euclidean distance of the translations.  Returns (distances_R, distances_t)."""
    if len(list_poses_1) != len(list_poses_2):
        raise Exception("The list of poses should be the same size")
    if len(list_poses_1) == 0:
        return [], []
    d2 = (_as_stack(list_poses_1) - _as_stack(list_poses_2)) ** 2
    # the reference sums the 3 x 3 block column-wise and then across: same association here
    d_R = [float(sum(sum(b[:3, :3])) ** 0.5) for b in d2]
    d_t = [float(sum(b[:3, 3]) ** 0.5) for b in d2]
    return d_R, d_t
