"""Minimal PCD v0.7 reader/writer and 4x4 pose text IO.

The reference hands data between stages as binary PCD clouds (``SIZE 4 TYPE F``, fields ``x y z`` or
``x y z rgb``; loaded at 2_MGICP_refinement_in_NCLT_dataset.py:169) and 4x4 text poses written with
``np.savetxt`` (1_FGR_pairwise_registration_in_NCLT_dataset.py:176-177).  Open3D promotes the float32
coordinates to float64 on load; so does this reader.
"""
from __future__ import annotations

import numpy as np


def read_pcd_xyz(path: str, remove_nan_points: bool = False, remove_infinite_points: bool = False) -> np.ndarray:
    """Return the N x 3 float64 coordinates of a PCD file (binary or ascii, float32 fields).
    Like o3d.io.read_point_cloud (Open3D >= 0.13, which the reference needs for registration_generalized_icp), non-finite
    points are KEPT unless asked otherwise (`remove_nan_points` / `remove_infinite_points` default to False), so point counts
    and indices match what the reference sees.  The engine itself rejects nothing: feed it finite clouds."""
    with open(path, "rb") as f:
        fields, sizes, types, counts, npts, data = [], [], [], [], None, None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PCD header")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok or tok[0].startswith("#"):
                continue
            key = tok[0].upper()
            if key == "FIELDS":
                fields = tok[1:]
            elif key == "SIZE":
                sizes = [int(t) for t in tok[1:]]
            elif key == "TYPE":
                types = tok[1:]
            elif key == "COUNT":
                counts = [int(t) for t in tok[1:]]
            elif key == "POINTS":
                npts = int(tok[1])
            elif key == "DATA":
                data = tok[1].lower()
                break
        if not counts:
            counts = [1] * len(fields)
        if npts is None or not fields:
            raise ValueError(f"{path}: incomplete PCD header")
        for name in ("x", "y", "z"):
            if name not in fields:
                raise ValueError(f"{path}: field {name} missing")
        if data == "binary":
            dt = []
            for name, s, t, c in zip(fields, sizes, types, counts):
                base = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 4): "<u4", ("I", 4): "<i4", ("U", 1): "u1", ("U", 2): "<u2",
                        ("I", 2): "<i2", ("I", 1): "i1"}[(t.upper(), s)]
                dt.append((name, base, (c,)) if c != 1 else (name, base))
            rec = np.frombuffer(f.read(npts * np.dtype(dt).itemsize), dtype=np.dtype(dt), count=npts)
            xyz = np.stack([rec["x"], rec["y"], rec["z"]], axis=1)
        elif data == "ascii":
            arr = np.loadtxt(f, dtype=np.float64, ndmin=2)
            cols = [fields.index(n) for n in ("x", "y", "z")]
            xyz = arr[:npts, cols].astype(np.float32)
        else:
            raise ValueError(f"{path}: DATA {data} not supported")
    xyz = np.asarray(xyz)
    if remove_nan_points:
        xyz = xyz[~np.isnan(xyz).any(axis=1)]
    if remove_infinite_points:
        xyz = xyz[~np.isinf(xyz).any(axis=1)]
    return xyz.astype(np.float64)


def write_pcd_xyz(path: str, xyz: np.ndarray) -> None:
    """Write an N x 3 cloud as binary PCD v0.7 with float32 ``x y z`` (the reference's NCLT layout)."""
    a = np.ascontiguousarray(xyz, dtype="<f4").reshape(-1, 3)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n"
           f"COUNT 1 1 1\nWIDTH {a.shape[0]}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {a.shape[0]}\nDATA binary\n")
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(a.tobytes())


def read_pose(path: str) -> np.ndarray:
    return np.loadtxt(path, dtype=np.float64).reshape(4, 4)


def write_pose(path: str, T: np.ndarray, fmt: str = "%.18e") -> None:
    np.savetxt(path, np.asarray(T, dtype=np.float64).reshape(4, 4), fmt=fmt)
