"""Seeded synthetic OS1-128 scan stream -- the workload of bench.py and tests.

Analytic ray-cast of an axis-aligned room with a few boxes and cylinders:
  * H beams uniformly over +-vfov/2 elevation, W azimuth columns over 360 deg, row-major [H][W];
  * room centred on the map: half-extents 0.45*side in x,y and min(0.30*side, 3 m) in z;
  * 8 seeded obstacles (5 boxes, 3 vertical cylinders) kept clear of the sensor track;
  * range noise N(0, 10 mm), rng = default_rng(1234 + frame);
  * trajectory: 0.10 m/frame along +x, 0.5 deg/frame yaw;
  * points are emitted as int32 millimetres; `points_map` are in the map frame, produced by the
    reference's own fixed-point transform (include/util/util.h:13-18, as src/warpsense/app.cpp:143).
No dedup: every one of the H*W returns is kept (all rays hit the room).
"""
import math

import numpy as np

from .fixedpoint import to_int_mat, transform_points

SCENE_SEED = 1234


def _rotz(yaw):
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def make_scene(map_side_m, seed=SCENE_SEED):
    """Room half extents + obstacle list (metres, map frame)."""
    hx = hy = 0.45 * map_side_m
    hz = min(0.30 * map_side_m, 3.0)
    rng = np.random.default_rng(seed)
    boxes, cyls = [], []
    for i in range(8):
        # keep the corridor |y| < 1.5 m (the sensor track) free
        side = 1.0 if i % 2 == 0 else -1.0
        cx = rng.uniform(-0.8 * hx, 0.8 * hx)
        cy = side * rng.uniform(2.5, 0.8 * hy)
        if i < 5:
            sx, sy = rng.uniform(0.4, 1.5, size=2)
            h = rng.uniform(0.8, 2.0 * hz)
            boxes.append((cx - sx, cy - sy, -hz, cx + sx, cy + sy, -hz + h))
        else:
            r = rng.uniform(0.3, 0.9)
            h = rng.uniform(1.0, 2.0 * hz)
            cyls.append((cx, cy, r, -hz, -hz + h))
    return {"half": (hx, hy, hz), "boxes": boxes, "cyls": cyls}


def beam_directions(H=128, W=1024, vfov_deg=45.0):
    """Unit ray directions in the sensor frame, row-major [H][W] -> [H*W, 3]."""
    el = np.deg2rad(np.linspace(-vfov_deg / 2.0, vfov_deg / 2.0, H))
    az = 2.0 * np.pi * np.arange(W) / W
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (H, W))], axis=-1)
    return d.reshape(-1, 3)


def raycast(scene, origin, dirs):
    """Range (m) along each unit direction from `origin` to the first surface."""
    o = np.asarray(origin, dtype=np.float64)
    d = np.asarray(dirs, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        half = np.array(scene["half"])
        # room: exit distance of the enclosing box
        t_exit = np.where(d > 0, (half - o) * inv, np.where(d < 0, (-half - o) * inv, np.inf))
        t = t_exit.min(axis=1)
        for (x0, y0, z0, x1, y1, z1) in scene["boxes"]:
            lo = (np.array([x0, y0, z0]) - o) * inv
            hi = (np.array([x1, y1, z1]) - o) * inv
            tmin = np.minimum(lo, hi).max(axis=1)
            tmax = np.maximum(lo, hi).min(axis=1)
            hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.0)
            t = np.where(hit & (tmin < t), tmin, t)
        for (cx, cy, r, z0, z1) in scene["cyls"]:
            a = d[:, 0] ** 2 + d[:, 1] ** 2
            ox, oy = o[0] - cx, o[1] - cy
            b = 2.0 * (ox * d[:, 0] + oy * d[:, 1])
            c = ox * ox + oy * oy - r * r
            disc = b * b - 4.0 * a * c
            ts = (-b - np.sqrt(np.maximum(disc, 0.0))) / (2.0 * a)
            z = o[2] + ts * d[:, 2]
            hit = (disc > 0) & (ts > 0) & (z >= z0) & (z <= z1)
            t = np.where(hit & (ts < t), ts, t)
            # top cap
            tc = (z1 - o[2]) * inv[:, 2]
            px, py = o[0] + tc * d[:, 0] - cx, o[1] + tc * d[:, 1] - cy
            hitc = (tc > 0) & (px * px + py * py <= r * r) & (d[:, 2] < 0)
            t = np.where(hitc & (tc < t), tc, t)
    return t


def frame_pose(frame, step_m=0.10, yaw_deg=0.5):
    """Ground-truth sensor pose of a frame: 4x4 float32, translation in millimetres.  The 100-frame trajectory
     is walked back and forth for longer runs, so the sensor never leaves the room."""
    frame = frame % 200
    if frame > 100:
        frame = 200 - frame
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = _rotz(math.radians(yaw_deg * frame))
    T[0, 3] = step_m * frame * 1000.0
    return T.astype(np.float32)


class ScanStream:
    """Deterministic scan generator.  `frame(k)` -> dict(points_sensor, pose, points_map, ...)."""

    def __init__(self, H=128, W=1024, grid_side=512, resolution=50, vfov_deg=45.0,
                 noise_mm=10.0, step_m=0.10, yaw_deg=0.5):
        self.H, self.W = H, W
        self.resolution = int(resolution)
        self.map_side_m = grid_side * resolution / 1000.0
        self.scene = make_scene(self.map_side_m)
        self.dirs = beam_directions(H, W, vfov_deg)
        self.noise_mm = noise_mm
        self.step_m, self.yaw_deg = step_m, yaw_deg

    def pose(self, k):
        return frame_pose(k, self.step_m, self.yaw_deg)

    def frame(self, k, prior_pose=None):
        """Scan k.  `points_map` uses the ground-truth pose; `points_prior` (if prior_pose is given)
        places the same returns with another pose (the odometry prior a registration starts from)."""
        pose = self.pose(k)
        R = pose[:3, :3].astype(np.float64)
        origin_m = pose[:3, 3].astype(np.float64) / 1000.0
        rng = np.random.default_rng(1234 + k)
        dw = self.dirs @ R.T
        rng_m = raycast(self.scene, origin_m, dw) + rng.normal(0.0, self.noise_mm / 1000.0, len(dw))
        keep = rng_m >= 0.3
        pts_sensor_m = (self.dirs[keep] * rng_m[keep, None]).astype(np.float32)
        # iter_x[i] * 1000 -> int (src/cpu/fastsense.cpp:152-155)
        pts_sensor = np.trunc(pts_sensor_m * np.float32(1000.0)).astype(np.int32)
        out = {
            "frame": k,
            "pose": pose,
            "points_sensor": pts_sensor,
            "points_sensor_m": pts_sensor_m,
            "points_map": transform_points(pts_sensor, to_int_mat(pose)),
        }
        if prior_pose is not None:
            out["points_prior"] = transform_points(pts_sensor, to_int_mat(prior_pose))
        return out
