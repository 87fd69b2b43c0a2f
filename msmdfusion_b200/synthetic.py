"""Seeded synthetic nuScenes-shaped inputs (host side, numpy) -- SURVEY.md section 8(d).

No dataset is reachable (no network); these generators reproduce the SHAPE of the inputs
the reference pipeline produces (``configs/MSMDFusion_nusc_voxel_LC.py:24-57``,
``mmdet3d/datasets/pipelines/my_loading_multi_proj.py``): a 32-beam LiDAR scan in the
``[-54,54]^2 x [-5,3]`` range, optionally 10 jittered sweeps, six cameras of virtual points.
"""
import numpy as np

POINT_CLOUD_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
VOXEL_SIZE = [0.075, 0.075, 0.2]


def lidar_sweep(rng, beams=32, azimuths=1090, sensor_height=1.84):
    """One 32-beam revolution: each ray hits the ground plane or an obstacle at U(4,54) m."""
    elev = np.deg2rad(np.linspace(-30.0, 10.0, beams))[:, None]
    azim = np.linspace(0, 2 * np.pi, azimuths, endpoint=False)[None, :]
    azim = azim + rng.uniform(0, 2 * np.pi / azimuths)
    obstacle = rng.uniform(4.0, 54.0, size=(beams, azimuths))
    with np.errstate(divide='ignore'):
        ground = np.where(elev < 0, sensor_height / np.tan(-elev), np.inf)
    ground = np.broadcast_to(ground, obstacle.shape)
    hits_obstacle = rng.random(obstacle.shape) < 0.35
    rng_m = np.where(hits_obstacle, np.minimum(obstacle, ground), ground)
    keep = np.isfinite(rng_m) & (rng_m < 75.0)
    r = rng_m + rng.normal(0, 0.02, size=rng_m.shape)
    x = r * np.cos(elev) * np.cos(azim)
    y = r * np.cos(elev) * np.sin(azim)
    z = r * np.sin(elev) + (sensor_height - 1.84)
    pts = np.stack([x, y, z], -1)[keep]
    return pts.astype(np.float32)


def lidar_scene(seed=0, sweeps=1, shuffle=True):
    """(N,5) f32 points = x,y,z,intensity,dt.  sweeps=1 -> ~30 k points (profile S);
    sweeps=10 -> ~300 k points (profile L, ``LoadPointsFromMultiSweeps sweeps_num=10``)."""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(sweeps):
        p = lidar_sweep(rng)
        if s > 0:  # ego-motion jitter between sweeps
            yaw = rng.normal(0, 0.01)
            c, sn = np.cos(yaw), np.sin(yaw)
            p = p @ np.array([[c, -sn, 0], [sn, c, 0], [0, 0, 1]], np.float32).T
            p = p + rng.normal(0, 0.15, size=(1, 3)).astype(np.float32) * np.array([1, 1, 0.1], np.float32)
        inten = rng.uniform(0, 255, size=(p.shape[0], 1)).astype(np.float32)
        dt = np.full((p.shape[0], 1), 0.05 * s, np.float32)
        out.append(np.concatenate([p, inten, dt], 1))
    pts = np.concatenate(out, 0).astype(np.float32)
    if shuffle:  # PointShuffle, configs/MSMDFusion_nusc_voxel_LC.py:50
        pts = pts[rng.permutation(pts.shape[0])]
    return np.ascontiguousarray(pts)


def random_points(n, c, seed=0, pc_range=POINT_CLOUD_RANGE, margin=1.05):
    """Uniform points slightly over-filling the range (so some fall outside), config-1 style."""
    rng = np.random.default_rng(seed)
    lo = np.array(pc_range[:3], np.float32)
    hi = np.array(pc_range[3:], np.float32)
    mid, half = (lo + hi) / 2, (hi - lo) / 2 * margin
    xyz = (mid + (rng.random((n, 3), dtype=np.float32) * 2 - 1) * half).astype(np.float32)
    extra = rng.random((n, max(c - 3, 0)), dtype=np.float32)
    return np.ascontiguousarray(np.concatenate([xyz, extra], 1)[:, :c])
