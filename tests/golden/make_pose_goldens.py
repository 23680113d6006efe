"""Generates tests/golden/pose_goldens.npz from the REFERENCE's own pose utilities.

Run in the build container only (needs /root/reference): the reference modules import Open3D at the top, which is not
installable here, so the pure-numpy functions are lifted out of the source files with `ast` and executed unmodified.
Inputs: seeded random circuits of rigid poses plus the two shipped NCLT golden poses.
"""
import ast
import os
import sys

import numpy as np

REF = "/root/reference"
WANT = {"ALL_FUNCTIONS.py": ["Transformar_de_volta", "compor_duas_poses", "Calcular_Erro_LoopClosure", "poses_relativas_para_absolutas",
                             "poses_absolutas_para_relativas", "subtract_squared_poses"],
        "2_MGICP_refinement_in_NCLT_dataset.py": ["relative_to_absolute_poses"]}


def lift(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "print": lambda *a, **k: None}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def random_pose(rng, ang=0.2, tr=2.0):
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    th = ang * rng.normal()
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    T[:3, 3] = tr * rng.normal(size=3)
    return T


def main():
    af = lift(os.path.join(REF, "ALL_FUNCTIONS.py"), WANT["ALL_FUNCTIONS.py"])
    s2 = lift(os.path.join(REF, "2_MGICP_refinement_in_NCLT_dataset.py"), WANT["2_MGICP_refinement_in_NCLT_dataset.py"])
    rng = np.random.default_rng(20261017)
    out = {}
    for c, n in enumerate((1, 2, 5, 17)):
        circ = [random_pose(rng) for _ in range(n)]
        out[f"c{c}_in"] = np.stack(circ)
        out[f"c{c}_abs_af"] = np.stack(af["poses_relativas_para_absolutas"](circ))
        out[f"c{c}_abs_s2"] = np.stack(s2["relative_to_absolute_poses"](circ))
        out[f"c{c}_closure"] = af["Calcular_Erro_LoopClosure"](circ)
        absolute = [np.eye(4)] + [random_pose(rng) for _ in range(n)]
        out[f"c{c}_abs_in"] = np.stack(absolute)
        out[f"c{c}_rel"] = np.stack(af["poses_absolutas_para_relativas"](absolute))
        other = [random_pose(rng) for _ in range(n)]
        dR, dt = af["subtract_squared_poses"](circ, other)
        out[f"c{c}_other"] = np.stack(other)
        out[f"c{c}_dR"], out[f"c{c}_dt"] = np.asarray(dR), np.asarray(dt)
    A, B = random_pose(rng), random_pose(rng)
    out["pair_in"] = np.stack([A, B])
    out["pair_comp"] = af["compor_duas_poses"](A, B)
    out["pair_inv"] = af["Transformar_de_volta"](A)
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pose_goldens.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    sys.exit(main())
