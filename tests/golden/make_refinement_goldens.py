"""Generates tests/golden/refinement_goldens.npz from the REFERENCE's own global-refinement functions.

Run in the build container only (needs /root/reference).  The reference scripts import Open3D / numpy-quaternion at the
top (neither installable here), so the functions are lifted out of the source files with `ast` and executed unmodified:

* the LUM adjustment (3_Global_Optimizations...py:194-224, ALL_FUNCTIONS.py:597-629 weighted) is pure numpy -> a true pin
  of the closed-form solve in global_refinement.py against the reference's dense  inv(A'PA) A'P Lb;
* the SLERP functions call `quat.*`; they are executed with `quat` bound to this repo's restatement of the few
  numpy-quaternion operations they use, which pins the ORCHESTRATION (which quaternions are composed, inverted and
  interpolated at which parameter, how the poses are assembled), not the quaternion package itself.
"""
import ast
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
S3 = "3_Global_Optimizations_in_NCLT_dataset.py"
WANT_S3 = ["Invert_pose", "Acumulate_Two_Poses", "Montar_Vetor_Lb_translacoes", "Ajustamento_Quaternios_SLERP",
           "reconstruir_Ts_para_origem_LUM", "reconstruir_Ts_para_origem_SLERP", "reconstruir_Ts_para_origem_SLERP_LUM"]
WANT_AF = ["Montar_Matriz_Diagonal_Pesos", "Montar_Vetor_Lb_translacoes", "Ajustamento_Quaternios_SLERP", "interpolar_duas_T",
           "reconstruir_Ts_para_origem_LUM", "reconstruir_Ts_para_origem_SLERP", "reconstruir_Ts_para_origem_SLERP_LUM"]


def lift(path, names, quat):
    tree = ast.parse(open(path).read())
    ns = {"np": np, "quat": quat, "print": lambda *a, **k: None}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def random_pose(rng, ang=0.05, tr=1.0):
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    th = ang * rng.normal()
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    T[:3, 3] = tr * rng.normal(size=3)
    return T


def circuit(rng, n, closure_err=0.02):
    """n relative poses of a nearly closed circuit: a loop of n poses on a circle, plus noise"""
    absolute = [np.eye(4)]
    for k in range(1, n):
        th = 2 * np.pi * k / n
        T = np.eye(4)
        T[:3, :3] = [[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]
        T[:3, 3] = [10 * np.sin(th), 10 * (1 - np.cos(th)), 0.1 * k]
        absolute.append(T)
    rel = []
    for k in range(n):
        a, b = absolute[k], absolute[(k + 1) % n]
        # reference convention: abs[k+1] = compose(rel[k], abs[k]) with R20 = R21 R10, t20 = R10 t21 + t10
        R = b[:3, :3] @ a[:3, :3].T
        t = a[:3, :3].T @ (b[:3, 3] - a[:3, 3])
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        rel.append(random_pose(rng, closure_err, closure_err) @ T)
    return rel


def main():
    gr = importlib.import_module("mgicp_b200.global_refinement")
    s3 = lift(os.path.join(REF, S3), WANT_S3, gr)
    af = lift(os.path.join(REF, "ALL_FUNCTIONS.py"), WANT_AF, gr)
    rng = np.random.default_rng(20261018)
    out = {}
    for c, n in enumerate((2, 3, 7, 40, 300)):
        rel = circuit(rng, n)
        w = list(rng.uniform(0.5, 2.0, size=n))
        out[f"c{c}_in"] = np.stack(rel)
        out[f"c{c}_w"] = np.asarray(w)
        out[f"c{c}_lum"] = np.stack(s3["reconstruir_Ts_para_origem_LUM"](rel))
        out[f"c{c}_lum_w"] = np.stack(af["reconstruir_Ts_para_origem_LUM"](rel, w))
        out[f"c{c}_slerp"] = np.stack(s3["reconstruir_Ts_para_origem_SLERP"](rel))
        out[f"c{c}_slerp_af"] = np.stack(af["reconstruir_Ts_para_origem_SLERP"](rel))
        out[f"c{c}_slerp_lum"] = np.stack(s3["reconstruir_Ts_para_origem_SLERP_LUM"](rel))
        out[f"c{c}_slerp_lum_w"] = np.stack(af["reconstruir_Ts_para_origem_SLERP_LUM"](rel, w))
    A, B = random_pose(rng, 0.8, 3.0), random_pose(rng, 0.8, 3.0)
    out["pair_in"] = np.stack([A, B])
    out["pair_inv"] = s3["Invert_pose"](A)
    out["pair_acc"] = s3["Acumulate_Two_Poses"](A, B)
    out["pair_interp"] = np.stack([af["interpolar_duas_T"](A, B, t) for t in (0.0, 0.25, 0.5, 1.0)])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "refinement_goldens.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    sys.exit(main())
