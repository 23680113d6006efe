"""Closed-circuit global refinement: SLERP on the rotations, Lu & Milios (LUM) on the translations, and both
(SURVEY.md 8(f) N4 -- the stage after the pairwise refinement).

Mirrors, with the reference's names, argument order and return shapes:

* 3_Global_Optimizations_in_NCLT_dataset.py:22-40   ``Invert_pose``, ``Acumulate_Two_Poses``
* 3_Global_Optimizations_in_NCLT_dataset.py:133-147 ``Montar_Vetor_Lb_translacoes``
* 3_Global_Optimizations_in_NCLT_dataset.py:155-187 ``Ajustamento_Quaternios_SLERP``
* 3_Global_Optimizations_in_NCLT_dataset.py:194-224 ``reconstruir_Ts_para_origem_LUM``
* 3_Global_Optimizations_in_NCLT_dataset.py:230-255 ``reconstruir_Ts_para_origem_SLERP``
* 3_Global_Optimizations_in_NCLT_dataset.py:263-292 ``reconstruir_Ts_para_origem_SLERP_LUM``
* ALL_FUNCTIONS.py:119-136                          ``interpolar_duas_T``
* ALL_FUNCTIONS.py:446-453, 597-668                 the weighted variants (``Pesos``), exposed as ``weights=``

These are n x 4 x 4 host-side operations on a circuit of <= a few thousand poses: microseconds of work next to the
registration itself, so they stay batched numpy (nothing here belongs on a GPU).  Two things differ from a transliteration:

* **The adjustment is solved in closed form.**  The reference builds the 3n x 3(n-1) design matrix A (+I on the
  diagonal, -I on the block sub-diagonal), forms A'PA and inverts it densely -- O(n^3), 2700 x 2700 for the NCLT circuit.
  The observations are the n rotated relative translations l_i with l_0 = x_0, l_i = x_i - x_{i-1}, l_{n-1} = -x_{n-2}:
  a chain constrained to close.  The weighted least-squares solution distributes the closure c = sum l_i over the links
  in proportion to 1/w_i:  x_j = sum_{i<=j} (l_i - c (1/w_i) / sum_k (1/w_k)).  O(n), and better conditioned than the
  dense inverse (the goldens generated from the reference's dense solve agree to ~1e-12 relative).
* **Quaternions.**  The reference uses the `numpy-quaternion` package (``import quaternion as quat``), which is not
  installable offline.  The few operations it needs are restated below from that package's published behaviour:
  ``from_rotation_matrix`` (default ``nonorthogonal=True``: dominant eigenvector of Bar-Itzhack's symmetric 4 x 4 matrix),
  Hamilton product, inverse, ``as_rotation_matrix`` (normalised by |q|^2) and ``quaternion_time_series.slerp`` =
  ``(q2 / q1) ** tau * q1`` on the short arc (q2 is negated when the chordal distance exceeds sqrt 2).  The eigenvector's
  sign is LAPACK's choice in the package; here it is fixed to w >= 0, which does not change any rotation matrix returned.
  Parity of this part against the package itself is **unpinned**; the orchestration (which quaternions are composed,
  inverted and interpolated at which parameter) is pinned by running the reference's own functions on this module's
  quaternion type (tests/golden/make_refinement_goldens.py).
"""
from __future__ import annotations

import numpy as np

from .poses import _as_stack, _assemble

__all__ = ["Invert_pose", "Acumulate_Two_Poses", "Montar_Vetor_Lb_translacoes", "Ajustamento_Quaternios_SLERP",
           "reconstruir_Ts_para_origem_LUM", "reconstruir_Ts_para_origem_SLERP", "reconstruir_Ts_para_origem_SLERP_LUM",
           "interpolar_duas_T", "Quaternion", "from_rotation_matrix", "as_rotation_matrix", "slerp", "lum_translations"]


# ------------------------------------------------------------------------------------------------------------------
# quaternions (w, x, y, z), Hamilton convention
# ------------------------------------------------------------------------------------------------------------------
class Quaternion:
    """The subset of ``quaternion.quaternion`` the reference touches: ``*``, ``/``, ``** scalar``, ``-q``, components."""
    __slots__ = ("w", "x", "y", "z")

    def __init__(self, w, x, y, z):
        self.w, self.x, self.y, self.z = float(w), float(x), float(y), float(z)

    @property
    def components(self) -> np.ndarray:
        return np.array([self.w, self.x, self.y, self.z])

    def __repr__(self):
        return f"quaternion({self.w!r}, {self.x!r}, {self.y!r}, {self.z!r})"

    def __neg__(self):
        return Quaternion(-self.w, -self.x, -self.y, -self.z)

    def __mul__(self, o):
        if isinstance(o, Quaternion):
            return Quaternion(self.w * o.w - self.x * o.x - self.y * o.y - self.z * o.z,
                              self.w * o.x + self.x * o.w + self.y * o.z - self.z * o.y,
                              self.w * o.y - self.x * o.z + self.y * o.w + self.z * o.x,
                              self.w * o.z + self.x * o.y - self.y * o.x + self.z * o.w)
        return Quaternion(self.w * o, self.x * o, self.y * o, self.z * o)

    def norm2(self) -> float:
        return self.w * self.w + self.x * self.x + self.y * self.y + self.z * self.z

    def inverse(self):
        n = self.norm2()
        return Quaternion(self.w / n, -self.x / n, -self.y / n, -self.z / n)

    def __truediv__(self, o):
        return self * o.inverse() if isinstance(o, Quaternion) else self * (1.0 / o)

    def log(self):
        """principal logarithm (the package's quaternion_log)"""
        b = np.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)
        n = np.sqrt(self.norm2())
        if b <= 1e-14 * abs(self.w):
            if self.w < 0.0:
                # log of a negative real: pi along x (the package's convention for the ambiguous axis)
                return Quaternion(np.log(-self.w), np.pi, 0.0, 0.0) if abs(self.w + 1.0) > 1e-14 else Quaternion(0.0, np.pi, 0.0, 0.0)
            return Quaternion(np.log(self.w), 0.0, 0.0, 0.0)
        v = np.arctan2(b, self.w)
        f = v / b
        return Quaternion(np.log(n), f * self.x, f * self.y, f * self.z)

    def exp(self):
        vn = np.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)
        e = np.exp(self.w)
        if vn > 1e-14:
            s = e * np.sin(vn) / vn
            return Quaternion(e * np.cos(vn), s * self.x, s * self.y, s * self.z)
        return Quaternion(e, 0.0, 0.0, 0.0)

    def __pow__(self, s):
        s = float(s)
        if s == -1.0:
            return self.inverse()
        if self.norm2() == 0.0:
            return Quaternion(1.0, 0, 0, 0) if s == 0.0 else Quaternion(0.0, 0, 0, 0)
        return (self.log() * s).exp()


def from_rotation_matrix(rot, nonorthogonal: bool = True) -> Quaternion:
    """``quaternion.from_rotation_matrix``: for the default ``nonorthogonal=True`` the optimal quaternion of a possibly
    slightly non-orthogonal matrix, i.e. the dominant eigenvector of Bar-Itzhack's K3 (rotations read from ``%.10f`` text
    are orthonormal to 1e-10 only, so this matters); sign fixed to w >= 0."""
    R = np.asarray(rot, dtype=np.float64)
    if R.shape != (3, 3):
        raise ValueError("expected a 3 x 3 rotation matrix")
    if nonorthogonal:
        K = np.empty((4, 4))
        K[0, 0] = (R[0, 0] - R[1, 1] - R[2, 2]) / 3.0
        K[0, 1] = (R[1, 0] + R[0, 1]) / 3.0
        K[0, 2] = (R[2, 0] + R[0, 2]) / 3.0
        K[0, 3] = (R[1, 2] - R[2, 1]) / 3.0
        K[1, 1] = (R[1, 1] - R[0, 0] - R[2, 2]) / 3.0
        K[1, 2] = (R[2, 1] + R[1, 2]) / 3.0
        K[1, 3] = (R[2, 0] - R[0, 2]) / 3.0
        K[2, 2] = (R[2, 2] - R[0, 0] - R[1, 1]) / 3.0
        K[2, 3] = (R[0, 1] - R[1, 0]) / 3.0
        K[3, 3] = (R[0, 0] + R[1, 1] + R[2, 2]) / 3.0
        K = np.triu(K) + np.triu(K, 1).T
        _, vec = np.linalg.eigh(K)
        v = vec[:, 3]                         # (x, y, z, w) of the conjugate: the package negates the vector part
        q = Quaternion(v[3], -v[0], -v[1], -v[2])
    else:
        # Shepperd-style branch on the largest diagonal combination (orthonormal input)
        d0, d1, d2 = R[0, 0], R[1, 1], R[2, 2]
        tr = d0 + d1 + d2
        if tr >= max(d0, d1, d2):
            w = np.sqrt(1.0 + tr) / 2.0
            q = Quaternion(w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w))
        elif d0 >= max(d1, d2):
            x = np.sqrt(1.0 + d0 - d1 - d2) / 2.0
            q = Quaternion((R[2, 1] - R[1, 2]) / (4 * x), x, (R[0, 1] + R[1, 0]) / (4 * x), (R[0, 2] + R[2, 0]) / (4 * x))
        elif d1 >= d2:
            y = np.sqrt(1.0 - d0 + d1 - d2) / 2.0
            q = Quaternion((R[0, 2] - R[2, 0]) / (4 * y), (R[0, 1] + R[1, 0]) / (4 * y), y, (R[1, 2] + R[2, 1]) / (4 * y))
        else:
            z = np.sqrt(1.0 - d0 - d1 + d2) / 2.0
            q = Quaternion((R[1, 0] - R[0, 1]) / (4 * z), (R[0, 2] + R[2, 0]) / (4 * z), (R[1, 2] + R[2, 1]) / (4 * z), z)
    return -q if q.w < 0.0 else q


def as_rotation_matrix(q: Quaternion) -> np.ndarray:
    """``quaternion.as_rotation_matrix``: valid for non-unit q (divides by |q|^2)."""
    n = q.norm2()
    if n == 0.0:
        raise ZeroDivisionError("quaternion has zero norm")
    s = 2.0 / n
    w, x, y, z = q.w, q.x, q.y, q.z
    return np.array([[1.0 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1.0 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1.0 - s * (x * x + y * y)]])


def slerp(R1: Quaternion, R2: Quaternion, t1, t2, t_out) -> Quaternion:
    """``quaternion.quaternion_time_series.slerp(R1, R2, t1, t2, t_out)``: (R2/R1)^tau R1 with tau = (t_out-t1)/(t2-t1),
    on the short arc (R2 negated when |R1 - R2| > sqrt 2)."""
    tau = (float(t_out) - float(t1)) / (float(t2) - float(t1))
    d = R1.components - R2.components
    if np.sqrt(d @ d) > 1.414213562373096:
        R2 = -R2
    return ((R2 / R1) ** tau) * R1


class _TimeSeries:
    slerp = staticmethod(slerp)


quaternion_time_series = _TimeSeries()          # so that `quat.quaternion_time_series.slerp(...)` reads like the reference


# ------------------------------------------------------------------------------------------------------------------
# poses
# ------------------------------------------------------------------------------------------------------------------
def Invert_pose(T_4x4):
    """3_Global_Optimizations...py:22-26 (== ALL_FUNCTIONS.py:110-114 Transformar_de_volta): [R^T | -R^T t]."""
    T = np.asarray(T_4x4, dtype=np.float64)
    Rt = T[:3, :3].T
    return _assemble(Rt, -Rt @ T[:3, 3])


def Acumulate_Two_Poses(T21, T10):
    """3_Global_Optimizations...py:35-40: R20 = R21 @ R10, t20 = R10 @ t21 + t10 (the reference's convention)."""
    T21, T10 = np.asarray(T21, dtype=np.float64), np.asarray(T10, dtype=np.float64)
    return _assemble(T21[:3, :3] @ T10[:3, :3], T10[:3, :3] @ T21[:3, 3] + T10[:3, 3])


def Montar_Vetor_Lb_translacoes(T_circuito, lista_rotacoes_origem):
    """3_Global_Optimizations...py:133-147: the 3n x 1 observation vector, l_i = R_i(origin) @ t_i."""
    T = _as_stack(T_circuito)
    R = np.asarray(lista_rotacoes_origem, dtype=np.float64)
    if R.shape[0] < T.shape[0]:
        raise ValueError("need one absolute rotation per relative pose")
    return np.einsum("nij,nj->ni", R[:T.shape[0]], T[:, :3, 3]).reshape(-1, 1)


def lum_translations(Lb, weights=None) -> np.ndarray:
    """Closed-form solution of the reference's  X = inv(A'PA) A'P Lb  (module docstring): adjusted absolute translations
    x_0 .. x_{n-2} as an (n-1) x 3 array."""
    L = np.asarray(Lb, dtype=np.float64).reshape(-1, 3)
    n = L.shape[0]
    if n < 2:
        return np.zeros((0, 3))
    iw = np.ones(n) if weights is None else 1.0 / np.asarray(weights, dtype=np.float64)[:n]
    c = L.sum(axis=0)
    return np.cumsum(L - np.outer(iw / iw.sum(), c), axis=0)[: n - 1]


def _absolute_rotations(T: np.ndarray) -> np.ndarray:
    """[I, R_0, R_1 R_0, ...]: `absolute_R = T[i].R @ absolute_R`, the value before the update being pose i's rotation
    (3_Global_Optimizations...py:199-203)."""
    out = np.empty((T.shape[0], 3, 3))
    acc = np.identity(3)
    for i in range(T.shape[0]):
        out[i] = acc
        acc = T[i, :3, :3] @ acc
    return out


def reconstruir_Ts_para_origem_LUM(T_circuito, Pesos=None):
    """3_Global_Optimizations...py:194-224 (ALL_FUNCTIONS.py:597-629 with ``Pesos``): rotations composed as they are,
    translations adjusted so that the circuit closes.  Returns n absolute poses, the first the identity."""
    T = _as_stack(T_circuito)
    n = T.shape[0]
    R = _absolute_rotations(T)
    X = lum_translations(Montar_Vetor_Lb_translacoes(T, R), Pesos)
    return [np.identity(4)] + [_assemble(R[i], X[i - 1]) for i in range(1, n)]


def Ajustamento_Quaternios_SLERP(relative_quat):
    """3_Global_Optimizations...py:155-187: absolute rotations composed forwards (r_i = q_{i-1} r_{i-1}) and backwards
    (the inverse of q_{n-1}^{-1}... accumulated from the closing end), interpolated pairwise at t = i/n.  Returns n
    quaternions, the first the identity."""
    n = len(relative_quat)
    fwd, bwd = [], []
    a = Quaternion(1.0, 0.0, 0.0, 0.0)
    r = Quaternion(1.0, 0.0, 0.0, 0.0)
    for i in range(1, n):
        a = relative_quat[i - 1] * a
        r = r * relative_quat[-i]
        fwd.append(a)
        bwd.append(r ** (-1))
    out = [Quaternion(1.0, 0.0, 0.0, 0.0)]
    for i in range(1, n):
        out.append(slerp(fwd[i - 1], bwd[-i], 0, 1, t_out=i / n))
    return out


def _circuit_quaternions(T: np.ndarray):
    return [from_rotation_matrix(T[i, :3, :3]) for i in range(T.shape[0])]


def reconstruir_Ts_para_origem_SLERP(T_circuito):
    """3_Global_Optimizations...py:230-255: rotations adjusted by SLERP, translations chained with the adjusted
    rotations (t_{i+1} = R_i t_i(rel) + t_i)."""
    T = _as_stack(T_circuito)
    n = T.shape[0]
    Rs = [as_rotation_matrix(q) for q in Ajustamento_Quaternios_SLERP(_circuit_quaternions(T))]
    poses = []
    t = np.zeros(3)
    for i in range(n):
        poses.append(_assemble(Rs[i], t))
        t = Rs[i] @ T[i, :3, 3] + t
    return poses


def reconstruir_Ts_para_origem_SLERP_LUM(T_circuito, Pesos=None):
    """3_Global_Optimizations...py:263-292 (ALL_FUNCTIONS.py:637-668 with ``Pesos``): SLERP rotations, then LUM on the
    translations rotated by them -- the reference's proposed global refinement."""
    T = _as_stack(T_circuito)
    n = T.shape[0]
    Rs = np.stack([as_rotation_matrix(q) for q in Ajustamento_Quaternios_SLERP(_circuit_quaternions(T))])
    X = lum_translations(Montar_Vetor_Lb_translacoes(T, Rs), Pesos)
    return [np.identity(4)] + [_assemble(Rs[i], X[i - 1]) for i in range(1, n)]


def interpolar_duas_T(T1, T2, t):
    """ALL_FUNCTIONS.py:119-136: linear interpolation of the translations, SLERP of the rotations at parameter t."""
    T1, T2 = np.asarray(T1, dtype=np.float64), np.asarray(T2, dtype=np.float64)
    q = slerp(from_rotation_matrix(T1[:3, :3]), from_rotation_matrix(T2[:3, :3]), 0, 1, t)
    return _assemble(as_rotation_matrix(q), T1[:3, 3] * (1 - t) + T2[:3, 3] * t)
