"""Pose bookkeeping against golden vectors produced by the reference's own functions
(tests/golden/make_pose_goldens.py lifts them out of /root/reference and runs them unmodified)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pose_goldens.npz"))


@pytest.mark.parametrize("c", [0, 1, 2, 3])
def test_circuit_functions_match_reference_goldens(pkg, c):
    P = pkg.poses
    # same association as the reference; numpy's small matrix products may still differ in the last bit with operand
    # alignment (BLAS head/tail paths), hence 2 ulp of the largest entries instead of bit equality
    eq = lambda a, b: np.allclose(a, b, rtol=0, atol=2e-15)
    circ = list(G[f"c{c}_in"])
    got = np.stack(P.relative_to_absolute_poses(circ))
    assert eq(got, G[f"c{c}_abs_s2"]) and eq(got, G[f"c{c}_abs_af"])
    assert eq(np.stack(P.poses_relativas_para_absolutas(circ)), G[f"c{c}_abs_af"])
    assert eq(P.Calcular_Erro_LoopClosure(circ), G[f"c{c}_closure"])
    assert eq(np.stack(P.poses_absolutas_para_relativas(list(G[f"c{c}_abs_in"]))), G[f"c{c}_rel"])
    dR, dt = P.subtract_squared_poses(circ, list(G[f"c{c}_other"]))
    assert np.allclose(dR, G[f"c{c}_dR"], rtol=1e-15, atol=0) and np.allclose(dt, G[f"c{c}_dt"], rtol=1e-15, atol=0)
    assert np.array_equal(circ[0], G[f"c{c}_in"][0])                     # inputs untouched


def test_pair_functions_and_properties(pkg):
    P = pkg.poses
    A, B = G["pair_in"]
    assert np.allclose(P.compor_duas_poses(A, B), G["pair_comp"], rtol=0, atol=2e-15)
    assert np.allclose(P.Transformar_de_volta(A), G["pair_inv"], rtol=0, atol=2e-15)
    assert np.allclose(P.Transformar_de_volta(A) @ A, np.eye(4), atol=1e-14)
    # absolute -> relative -> absolute is the identity map on a circuit that starts at the identity
    absolute = list(G["c3_abs_in"])
    rel = P.poses_absolutas_para_relativas(absolute)
    back = P.relative_to_absolute_poses(rel + [np.eye(4)])                # the dropped closure slot
    assert np.allclose(np.stack(back), np.stack(absolute), atol=1e-12)
    with pytest.raises(Exception):
        P.subtract_squared_poses([A], [A, B])
    assert P.subtract_squared_poses([], []) == ([], [])
    with pytest.raises(ValueError):
        P.relative_to_absolute_poses([np.eye(3)])


def test_pose_files_roundtrip_both_reference_formats(pkg, tmp_path):
    """'%.10f' (1_FGR...py:177, the hand-off INTO the refinement) and '%.18e' (np.savetxt default, the hand-off OUT)"""
    T = G["pair_in"][0]
    p18, p10 = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    pkg.pcd_io.write_pose(p18, T)
    pkg.pcd_io.write_pose(p10, T, fmt="%.10f")
    assert np.array_equal(pkg.pcd_io.read_pose(p18), T)
    assert np.abs(pkg.pcd_io.read_pose(p10) - T).max() < 5.1e-11
