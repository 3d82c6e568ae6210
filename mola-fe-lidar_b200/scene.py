"""Seeded synthetic KITTI-shaped LiDAR scans.

A street corridor of axis-aligned boxes (buildings, parked cars), vertical
cylinders (poles / trunks), a ground plane at z = -1.73 m and four far boundary
walls so that every ray returns a hit (=> exactly beams x azimuth_steps points
per scan: 64 x 1875 = 120,000 or 128 x 2032 = 260,096).  Analytic ray casting,
range noise N(0, sigma), output float32 in the SENSOR frame.

Only numpy; used by bench.py, tests/ and __graft_entry__.smoke() to build the
workloads BASELINE.json names.  There is no dataset in this container.
"""
import math

import numpy as np

GROUND_Z = -1.73


def _rot_ypr(yaw, pitch, roll):
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cr, sr = math.cos(roll), math.sin(roll)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])


def pose_matrix(x, y, z, yaw, pitch=0.0, roll=0.0):
    T = np.eye(4)
    T[:3, :3] = _rot_ypr(yaw, pitch, roll)
    T[:3, 3] = (x, y, z)
    return T


def matrix_to_pose6(T):
    R = T[:3, :3]
    cpitch = math.hypot(R[0, 0], R[1, 0])
    pitch = math.atan2(-R[2, 0], cpitch)
    if cpitch < 1e-12:
        roll, yaw = 0.0, math.atan2(-R[0, 1], R[1, 1])
    else:
        yaw, roll = math.atan2(R[1, 0], R[0, 0]), math.atan2(R[2, 1], R[2, 2])
    return np.array([T[0, 3], T[1, 3], T[2, 3], yaw, pitch, roll])


class World:
    """Boxes = (xmin,ymin,zmin,xmax,ymax,zmax); cylinders = (x,y,r,zmin,zmax)."""

    def __init__(self, seed=1, half_length=140.0):
        rng = np.random.default_rng(seed)
        boxes = []
        # buildings on both sides of a street running along +x
        for side in (-1.0, 1.0):
            x = -half_length + 5.0
            while x < half_length - 35.0:
                w = rng.uniform(8.0, 30.0)
                setback = rng.uniform(8.0, 20.0)
                depth = rng.uniform(10.0, 20.0)
                h = rng.uniform(6.0, 15.0)
                y0, y1 = sorted((side * setback, side * (setback + depth)))
                boxes.append((x, y0, GROUND_Z, x + w, y1, GROUND_Z + h))
                x += w + rng.uniform(0.0, 6.0)
        # parked cars
        for _ in range(10):
            cx = rng.uniform(-60.0, 90.0)
            cy = rng.choice((-1.0, 1.0)) * rng.uniform(2.8, 3.8)
            boxes.append((cx - 2.2, cy - 0.9, GROUND_Z, cx + 2.2, cy + 0.9, GROUND_Z + 1.5))
        # boundary walls: every ray hits something
        L, H, T = half_length, 100.0, 2.0  # tall enough for the +15 deg beams of the 128-beam sensor
        boxes += [(L, -L - T, GROUND_Z, L + T, L + T, H), (-L - T, -L - T, GROUND_Z, -L, L + T, H),
                  (-L, L, GROUND_Z, L, L + T, H), (-L, -L - T, GROUND_Z, L, -L, H)]
        self.boxes = np.array(boxes, dtype=np.float64)
        cyl = []
        for _ in range(30):
            cx = rng.uniform(-70.0, 100.0)
            cy = rng.choice((-1.0, 1.0)) * rng.uniform(4.5, 7.5)
            cyl.append((cx, cy, rng.uniform(0.15, 0.4), GROUND_Z, GROUND_Z + rng.uniform(3.0, 8.0)))
        self.cylinders = np.array(cyl, dtype=np.float64)

    def raycast(self, origin, dirs):
        """Nearest hit distance along unit rays `dirs` (N,3) from `origin`."""
        o = np.asarray(origin, dtype=np.float64)
        d = np.asarray(dirs, dtype=np.float64)
        n = len(d)
        best = np.full(n, np.inf)
        # ground
        with np.errstate(divide="ignore", invalid="ignore"):
            tg = (GROUND_Z - o[2]) / d[:, 2]
            tg = np.where((d[:, 2] < 0) & (tg > 0), tg, np.inf)
            best = np.minimum(best, tg)
            inv = 1.0 / d
            for b in self.boxes:
                t0 = (b[:3] - o) * inv
                t1 = (b[3:] - o) * inv
                tn = np.minimum(t0, t1).max(axis=1)
                tf = np.maximum(t0, t1).min(axis=1)
                hit = (tf >= tn) & (tf > 0)
                t = np.where(tn > 0, tn, tf)
                best = np.where(hit & (t < best), t, best)
            a = d[:, 0] ** 2 + d[:, 1] ** 2
            for c in self.cylinders:
                ox, oy = o[0] - c[0], o[1] - c[1]
                bq = 2.0 * (ox * d[:, 0] + oy * d[:, 1])
                cq = ox * ox + oy * oy - c[2] ** 2
                disc = bq * bq - 4.0 * a * cq
                t = (-bq - np.sqrt(np.maximum(disc, 0.0))) / (2.0 * a)
                z = o[2] + t * d[:, 2]
                hit = (disc > 0) & (t > 0) & (z >= c[3]) & (z <= c[4])
                best = np.where(hit & (t < best), t, best)
        return best


def beam_directions(n_beams=64, n_azimuth=1875, elev_top_deg=2.0, elev_bottom_deg=-24.8):
    """Unit ray directions in the sensor frame, azimuth-major (firing order)."""
    el = np.deg2rad(np.linspace(elev_top_deg, elev_bottom_deg, n_beams))
    az = np.linspace(0.0, 2.0 * np.pi, n_azimuth, endpoint=False)
    azg, elg = np.meshgrid(az, el, indexing="ij")
    d = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], axis=-1)
    return d.reshape(-1, 3)


def trajectory(n_scans, v=10.0, yaw_rate=0.05, hz=10.0, x0=-40.0):
    """Sensor poses (4x4, world frame). 1.0 m and ~0.29 deg per scan; the yaw
    rate flips sign every 20 scans so that long sequences stay in the street."""
    poses = []
    x, y, yaw = x0, 0.0, 0.0
    dt = 1.0 / hz
    for i in range(n_scans):
        poses.append(pose_matrix(x, y, 0.0, yaw))
        w = yaw_rate * (1.0 if ((i + 10) // 20) % 2 == 0 else -1.0)
        x += v * dt * math.cos(yaw)
        y += v * dt * math.sin(yaw)
        yaw += w * dt
    return poses


def make_scan(world, pose, rng, n_beams=64, n_azimuth=1875, sigma=0.02, elev=(2.0, -24.8)):
    """One scan (N,3) float32 in the sensor frame."""
    d_s = beam_directions(n_beams, n_azimuth, elev[0], elev[1])
    R, o = pose[:3, :3], pose[:3, 3]
    d_w = d_s @ R.T
    rng_m = world.raycast(o, d_w)
    assert np.isfinite(rng_m).all(), "every ray must hit (boundary walls)"
    rng_m = rng_m + rng.normal(0.0, sigma, size=rng_m.shape) if sigma > 0 else rng_m
    return (d_s * rng_m[:, None]).astype(np.float32)


def make_sequence(n_scans, seed=1, beams=64, sigma=0.02):
    """(scans, poses): KITTI-shaped 64-beam (120,000 pts) or 128-beam
    (260,096 pts) scans along the default trajectory."""
    world = World(seed)
    rng = np.random.default_rng(seed + 1000)
    poses = trajectory(n_scans)
    if beams == 64:
        kw = dict(n_beams=64, n_azimuth=1875, elev=(2.0, -24.8))
    elif beams == 128:
        kw = dict(n_beams=128, n_azimuth=2032, elev=(15.0, -25.0))
    else:
        raise ValueError("beams must be 64 or 128")
    scans = [make_scan(world, T, rng, sigma=sigma, **kw) for T in poses]
    return scans, poses


def relative_pose6(T_from, T_to):
    """Pose of `to` with respect to `from` (the quantity ICP estimates,
    LidarOdometry.cpp:272-283) as (x,y,z,yaw,pitch,roll)."""
    return matrix_to_pose6(np.linalg.inv(T_from) @ T_to)


def make_pair_c1(seed=1, n=20000, sigma=0.0):
    """BASELINE config C1: A = uniform subsample of a 64-beam scan; B = T^-1 A
    with T = (0.30,-0.20,0.05 m; yaw 2, pitch 0.5, roll -0.3 deg).
    Returns (A_global, B_local, T_pose6)."""
    world = World(seed)
    rng = np.random.default_rng(seed + 2000)
    scan = make_scan(world, trajectory(1)[0], rng, sigma=0.0)
    sel = np.sort(rng.choice(len(scan), size=n, replace=False))
    A = scan[sel].astype(np.float64)
    pose6 = np.array([0.30, -0.20, 0.05, math.radians(2.0), math.radians(0.5), math.radians(-0.3)])
    T = pose_matrix(*pose6)
    Ti = np.linalg.inv(T)
    B = A @ Ti[:3, :3].T + Ti[:3, 3]
    if sigma > 0:
        B = B + rng.normal(0.0, sigma, size=B.shape)
    return A.astype(np.float32), B.astype(np.float32), pose6
