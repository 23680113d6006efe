"""Synthetic Velodyne-HDL-32-like scan pairs of NCLT shape.

There is no network, so the benchmark and the parity tests run on procedurally generated scans:
32 beams between -30.67 and +10.67 degrees times ``azimuth_steps`` firings, Gaussian range noise,
a ground plane, an outer wall box and random boxes.  Coordinates are rounded to float32 and promoted
to float64, like clouds loaded from the reference's ``SIZE 4 TYPE F`` PCD files
(2_MGICP_refinement_in_NCLT_dataset.py:169).  The continuous noise makes exact distance ties
measure-zero, which the parity rules require.
"""
from __future__ import annotations

import numpy as np

_BEAMS = np.deg2rad(np.linspace(-30.67, 10.67, 32))


def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _rot_xyz(rx, ry, rz):
    cx, sx, cy, sy = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    return _rot_z(rz) @ Ry @ Rx


def make_pose(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


class Scene:
    """Ground plane z=0, an outer 120 x 100 x 12 m wall box and ``n_boxes`` random boxes."""

    def __init__(self, seed: int = 0, n_boxes: int = 40):
        rng = np.random.default_rng(seed)
        self.outer = np.array([[-60.0, -50.0, 0.0], [60.0, 50.0, 12.0]])
        c = np.stack([rng.uniform(-55, 55, n_boxes), rng.uniform(-45, 45, n_boxes)], axis=1)
        keep = np.hypot(c[:, 0] - 17.0, c[:, 1]) > 4.0  # keep the sensor's own circuit (radius ~17 m) free
        c = c[keep]
        sz = rng.uniform(1.0, 8.0, (c.shape[0], 2))
        h = rng.uniform(1.0, 9.0, c.shape[0])
        self.lo = np.concatenate([c - sz / 2, np.zeros((c.shape[0], 1))], axis=1)
        self.hi = np.concatenate([c + sz / 2, h[:, None]], axis=1)

    def cast(self, origin: np.ndarray, dirs: np.ndarray, rmax: float) -> np.ndarray:
        """Range along each unit ray (inf where nothing is hit within rmax)."""
        n = dirs.shape[0]
        best = np.full(n, np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / dirs
            # ground plane z = 0
            tg = -origin[2] * inv[:, 2]
            tg[~(tg > 0)] = np.inf
            best = np.minimum(best, tg)
            # outer box, seen from inside: exit distance
            t1 = (self.outer[0] - origin) * inv
            t2 = (self.outer[1] - origin) * inv
            texit = np.nanmin(np.maximum(t1, t2), axis=1)
            texit[~(texit > 0)] = np.inf
            best = np.minimum(best, texit)
            for lo, hi in zip(self.lo, self.hi):
                a = (lo - origin) * inv
                b = (hi - origin) * inv
                tn = np.nanmax(np.minimum(a, b), axis=1)
                tf = np.nanmin(np.maximum(a, b), axis=1)
                hit = (tn <= tf) & (tn > 0)
                best = np.where(hit & (tn < best), tn, best)
        best[best > rmax] = np.inf
        return best


def sensor_pose(k: int, step: float = 0.6, yaw_deg: float = 2.0, height: float = 1.9) -> np.ndarray:
    """World pose of the sensor at scan ``k``: a circuit with NCLT's median motion (0.61 m / 1.95 deg per scan)."""
    yaw = np.deg2rad(yaw_deg) * k
    rad = step / np.deg2rad(yaw_deg)
    t = np.array([rad * np.sin(yaw), rad * (1 - np.cos(yaw)), height])
    # small deterministic roll/pitch wobble so that pairs are not pure planar motions
    R = _rot_xyz(0.01 * np.sin(0.37 * k), 0.012 * np.cos(0.23 * k), yaw)
    return make_pose(R, t)


def make_scan(scene: Scene, pose: np.ndarray, azimuth_steps: int, seed: int, sigma: float = 0.02, rmin: float = 1.0,
              rmax: float = 100.0) -> np.ndarray:
    """One scan in the SENSOR frame, N x 3 float64 holding float32-representable values."""
    rng = np.random.default_rng(seed)
    az = np.linspace(0.0, 2 * np.pi, azimuth_steps, endpoint=False) + rng.uniform(0, 2 * np.pi / azimuth_steps)
    el, azg = np.meshgrid(_BEAMS, az, indexing="ij")
    d_s = np.stack([np.cos(el) * np.cos(azg), np.cos(el) * np.sin(azg), np.sin(el)], axis=-1).reshape(-1, 3)
    d_w = d_s @ pose[:3, :3].T
    rng_true = scene.cast(pose[:3, 3], d_w, rmax)
    ok = np.isfinite(rng_true) & (rng_true >= rmin)
    r = rng_true[ok] + rng.normal(0.0, sigma, ok.sum())
    pts = d_s[ok] * r[:, None]
    return pts.astype(np.float32).astype(np.float64)


def perturbation(rng: np.random.Generator, rot_deg: float = 0.74, trans: float = 0.15) -> np.ndarray:
    """An FGR-sized error (reference data: median 0.74 deg / 0.15 m between FGR and refined poses)."""
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rot_deg) * abs(rng.normal(1.0, 0.3))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    t = rng.normal(size=3)
    t *= trans * abs(rng.normal(1.0, 0.3)) / np.linalg.norm(t)
    return make_pose(R, t)


def make_pair(azimuth_steps: int = 1000, seed: int = 0, k: int = 0, gap: int = 1, scene: Scene | None = None,
              rot_deg: float = 0.74, trans: float = 0.15):
    """(source, target, T_init, T_true): source = scan k+gap, target = scan k (the reference registers
    clouds[i+1] onto clouds[i], S2:191); T maps source-frame points into the target frame."""
    scene = scene or Scene(seed=12345)
    Pt, Ps = sensor_pose(k), sensor_pose(k + gap)
    tgt = make_scan(scene, Pt, azimuth_steps, seed=1000003 * seed + 2 * k)
    src = make_scan(scene, Ps, azimuth_steps, seed=1000003 * seed + 2 * (k + gap) + 1)
    T_true = np.linalg.inv(Pt) @ Ps
    rng = np.random.default_rng(77 + 1000003 * seed + k)
    T_init = perturbation(rng, rot_deg, trans) @ T_true
    return src, tgt, T_init, T_true


def make_sequence(n_scans: int, azimuth_steps: int = 3125, seed: int = 0):
    """Consecutive scans + per-pair (T_init, T_true) for pairs (i+1 -> i), config 3 of BASELINE.json."""
    scene = Scene(seed=12345)
    scans = [make_scan(scene, sensor_pose(k), azimuth_steps, seed=1000003 * seed + k) for k in range(n_scans)]
    inits, truths = [], []
    for k in range(n_scans - 1):
        T_true = np.linalg.inv(sensor_pose(k)) @ sensor_pose(k + 1)
        rng = np.random.default_rng(77 + 1000003 * seed + k)
        inits.append(perturbation(rng) @ T_true)
        truths.append(T_true)
    return scans, inits, truths


def pose_error(T_a: np.ndarray, T_b: np.ndarray):
    """(rotation angle in rad, translation distance in m) between two 4x4 poses."""
    dR = T_a[:3, :3] @ T_b[:3, :3].T
    c = np.clip((np.trace(dR) - 1.0) / 2.0, -1.0, 1.0)
    # for tiny angles use the antisymmetric part (acos loses half the digits near 1)
    s = 0.5 * np.sqrt((dR[2, 1] - dR[1, 2]) ** 2 + (dR[0, 2] - dR[2, 0]) ** 2 + (dR[1, 0] - dR[0, 1]) ** 2)
    return float(np.arctan2(s, c)), float(np.linalg.norm(T_a[:3, 3] - T_b[:3, 3]))


# ---- config 4: dense TLS-like pair (Courtyard / Facade shape) -----------------------------------------------------------
def make_tls_pair(n_points: int = 2_000_000, seed: int = 0, extent=(50.0, 45.0, 16.0), sigma: float = 0.003,
                  rot_deg: float = 0.74, trans: float = 0.15):
    """(source, target, T_init, T_true) for a dense terrestrial-laser-scanner-like pair: a courtyard of size ``extent``
    (ground, four facades with window recesses, a few free-standing blocks) sampled independently by two scanner
    set-ups ~6 m apart with inverse-square density fall-off and millimetre noise; every cloud is expressed in its own
    scanner frame.  About ``n_points`` points per cloud, float32-representable coordinates."""
    rng = np.random.default_rng(1234567 + seed)
    ex, ey, ez = extent
    # planar patches: (origin, edge u, edge v); points = o + a u + b v with a, b in [0, 1)
    patches = [((0, 0, 0), (ex, 0, 0), (0, ey, 0))]                                   # ground
    for (o, u) in (((0, 0, 0), (ex, 0, 0)), ((0, ey, 0), (ex, 0, 0)), ((0, 0, 0), (0, ey, 0)), ((ex, 0, 0), (0, ey, 0))):
        patches.append((o, u, (0, 0, ez)))                                            # facades
    for _ in range(6):                                                                # free-standing blocks
        cx, cy = rng.uniform(0.2 * ex, 0.8 * ex), rng.uniform(0.2 * ey, 0.8 * ey)
        sx, sy, sz = rng.uniform(1.5, 5.0), rng.uniform(1.5, 5.0), rng.uniform(1.0, 6.0)
        patches += [((cx, cy, 0), (sx, 0, 0), (0, 0, sz)), ((cx, cy + sy, 0), (sx, 0, 0), (0, 0, sz)),
                    ((cx, cy, 0), (0, sy, 0), (0, 0, sz)), ((cx + sx, cy, 0), (0, sy, 0), (0, 0, sz)),
                    ((cx, cy, sz), (sx, 0, 0), (0, sy, 0))]
    P = np.array([[o, u, v] for o, u, v in patches], float)
    area = np.linalg.norm(np.cross(P[:, 1], P[:, 2]), axis=1)

    def surface(m, rs):
        which = rs.choice(len(P), size=m, p=area / area.sum())
        a, b = rs.random(m), rs.random(m)
        pts = P[which, 0] + a[:, None] * P[which, 1] + b[:, None] * P[which, 2]
        # window recesses on the facades: push a regular pattern of rectangles 0.3 m into the wall
        nrm = np.cross(P[which, 1], P[which, 2])
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        facade = (which >= 1) & (which <= 4)
        wa, wb = (a * 12.0) % 1.0, (b * 5.0) % 1.0
        recess = facade & (wa > 0.3) & (wa < 0.7) & (wb > 0.35) & (wb < 0.8)
        inward = np.sign(((np.array([ex / 2, ey / 2, ez / 2]) - pts) * nrm).sum(1))
        pts[recess] -= 0.3 * (inward[recess, None] * nrm[recess])
        return pts

    def scan(center, n, rs):
        # uniform per area, kept with probability min(1, (8 m / r)^2): closer surfaces are denser, like a TLS
        def keep_prob(pts):
            return np.minimum(1.0, 64.0 / ((pts - center) ** 2).sum(1))
        frac = keep_prob(surface(50_000, rs)).mean()
        out, have = [], 0
        while have < n:
            pts = surface(int(min(4_000_000, 1.1 * (n - have) / frac + 1000)), rs)
            pts = pts[rs.random(len(pts)) < keep_prob(pts)]
            out.append(pts)
            have += len(pts)
        pts = np.concatenate(out)[:n]
        return pts + rs.normal(0.0, sigma, pts.shape)

    c_t = np.array([0.45 * ex, 0.5 * ey, 1.7])
    c_s = c_t + np.array([5.0, 3.0, 0.05])
    Pt = make_pose(_rot_xyz(0.004, -0.006, 0.3), c_t)
    Ps = make_pose(_rot_xyz(-0.005, 0.003, 0.85), c_s)
    world_t, world_s = scan(c_t, n_points, rng), scan(c_s, n_points, rng)
    to_local = lambda W, Pose: (W - Pose[:3, 3]) @ Pose[:3, :3]
    tgt = to_local(world_t, Pt).astype(np.float32).astype(np.float64)
    src = to_local(world_s, Ps).astype(np.float32).astype(np.float64)
    T_true = np.linalg.inv(Pt) @ Ps
    T_init = perturbation(np.random.default_rng(99 + seed), rot_deg, trans) @ T_true
    return src, tgt, T_init, T_true
