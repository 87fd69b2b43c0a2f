"""Seeded synthetic nuScenes-shaped inputs (host side, numpy) -- SURVEY.md section 8(d).

No dataset is reachable (no network); these generators reproduce the SHAPE of the inputs
the reference pipeline produces (``configs/MSMDFusion_nusc_voxel_LC.py:24-57``,
``mmdet3d/datasets/pipelines/my_loading_multi_proj.py``): a 32-beam LiDAR scan in the
``[-54,54]^2 x [-5,3]`` range, optionally 10 jittered sweeps, six cameras of virtual points.
"""
import os

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


# --------------------------------------------------------------------------------------
# cameras + virtual points (SURVEY.md section 8(d) config 3, Appendix B input schema)
# --------------------------------------------------------------------------------------
INPUT_SHAPE = (448, 800)          # network-input (H, W), configs/MSMDFusion_nusc_voxel_LC.py:17,55-57
CAMERA_YAWS_DEG = (0.0, -55.0, 55.0, 180.0, 110.0, -110.0)  # front, front-right/left, back, back-left/right


def camera_matrices(seed=0):
    """Six synthetic ``lidar2img`` (4,4) float64 matrices for an 800x448 input image
    (nuScenes intrinsics scaled by 0.5; camera x right, y down, z forward)."""
    rng = np.random.default_rng(seed + 1000)
    H, W = INPUT_SHAPE
    f = 633.0
    K = np.array([[f, 0, W / 2.0, 0], [0, f, H / 2.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    mats = []
    for yaw_deg in CAMERA_YAWS_DEG:
        yaw = np.deg2rad(yaw_deg + rng.normal(0, 0.5))
        fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0])
        right = np.array([np.sin(yaw), -np.cos(yaw), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        R = np.stack([right, down, fwd], 0)                  # lidar -> camera axes
        t = -R @ (np.array([0.0, 0.0, 1.5]) + 0.8 * fwd)     # camera 1.5 m up, 0.8 m ahead
        E = np.eye(4)
        E[:3, :3], E[:3, 3] = R, t
        mats.append(K @ E)
    return mats


def project(points_xyz, lidar2img):
    """(N,3) lidar points -> (u, v, depth) in network-input pixels + in-image mask."""
    H, W = INPUT_SHAPE
    p = np.concatenate([points_xyz.astype(np.float64), np.ones((points_xyz.shape[0], 1))], 1) @ lidar2img.T
    depth = p[:, 2]
    with np.errstate(divide='ignore', invalid='ignore'):
        u, v = p[:, 0] / depth, p[:, 1] / depth
    ok = (depth > 0.5) & (u >= 0) & (u < W - 1e-3) & (v >= 0) & (v < H - 1e-3)
    return np.stack([u, v, depth], 1), ok


def camera_scene(seed=0, lidar_points=None, virtual_per_camera=10000, real_per_camera=2000,
                 boxes_per_camera=4, empty_cameras=()):
    """One sample's ``img_metas`` entry: ``lidar2img``, ``input_shape``, ``pad_shape`` and
    ``foreground2D_info`` = {fg_pixels, fg_points, fg_real_pixels} per camera (Appendix B).

    Virtual points are sampled inside random object boxes in each camera's frustum (15 dims =
    xyz + 10 one-hot class + score + dt); real pixels are projected LiDAR returns.
    """
    rng = np.random.default_rng(seed + 2000)
    mats = camera_matrices(seed)
    fg_pixels, fg_points, fg_real = [], [], []
    for cam, (mat, yaw_deg) in enumerate(zip(mats, CAMERA_YAWS_DEG)):
        if cam in empty_cameras:
            fg_pixels.append(np.zeros((0, 3), np.float32))
            fg_points.append(np.zeros((0, 15), np.float32))
            fg_real.append(np.zeros((0, 3), np.float32))
            continue
        yaw = np.deg2rad(yaw_deg)
        pts = []
        for _ in range(boxes_per_camera):
            dist = rng.uniform(6.0, 45.0)
            ang = yaw + np.deg2rad(rng.uniform(-25, 25))
            centre = np.array([dist * np.cos(ang), dist * np.sin(ang), rng.uniform(-1.5, 0.0)])
            size = np.array([rng.uniform(1.5, 5.0), rng.uniform(1.5, 2.5), rng.uniform(1.4, 2.5)])
            n = 2 * virtual_per_camera // boxes_per_camera
            local = (rng.random((n, 3)) - 0.5) * size
            byaw = rng.uniform(0, np.pi)
            c, s = np.cos(byaw), np.sin(byaw)
            local[:, :2] = local[:, :2] @ np.array([[c, -s], [s, c]]).T
            cls = rng.integers(0, 10)
            onehot = np.zeros((n, 10), np.float32)
            onehot[:, cls] = 1.0
            score = np.full((n, 1), rng.uniform(0.3, 1.0), np.float32)
            pts.append(np.concatenate([(local + centre).astype(np.float32), onehot, score,
                                       np.zeros((n, 1), np.float32)], 1))
        pts = np.concatenate(pts, 0)
        pix, ok = project(pts[:, :3], mat)
        inside = (np.abs(pts[:, 0]) < 53.9) & (np.abs(pts[:, 1]) < 53.9) & (pts[:, 2] > -4.9) & (pts[:, 2] < 2.9)
        keep = np.nonzero(ok & inside)[0][:virtual_per_camera]
        fg_pixels.append(np.ascontiguousarray(pix[keep], np.float32))
        fg_points.append(np.ascontiguousarray(pts[keep], np.float32))
        if lidar_points is not None and lidar_points.shape[0]:
            rpix, rok = project(lidar_points[:, :3], mat)
            ridx = np.nonzero(rok)[0]
            if ridx.shape[0] > real_per_camera:
                ridx = rng.choice(ridx, real_per_camera, replace=False)
            fg_real.append(np.ascontiguousarray(rpix[ridx], np.float32))
        else:
            fg_real.append(np.zeros((0, 3), np.float32))
    return dict(lidar2img=mats, input_shape=INPUT_SHAPE, pad_shape=(INPUT_SHAPE[0], INPUT_SHAPE[1], 3),
                foreground2D_info=dict(fg_pixels=fg_pixels, fg_points=fg_points, fg_real_pixels=fg_real))


def fpn_features(seed=0, batch=1, ncam=6, channels=256):
    """Random stand-ins for the three FPN levels the lift reads (strides 4, 8, 16 of 448x800)."""
    rng = np.random.default_rng(seed + 3000)
    H, W = INPUT_SHAPE
    return [rng.standard_normal((batch * ncam, channels, H // s, W // s), dtype=np.float32) for s in (4, 8, 16)]


# --------------------------------------------------------------------------------------------
# Virtual-point WIRE FORMAT: the per-sweep `.pkl.npy` files the reference's
# LoadForeground2D / LoadForeground2DFromMultiSweeps read
# (mmdet3d/datasets/pipelines/my_loading_multi_proj.py:17-35,126-131,307-311)
# --------------------------------------------------------------------------------------------
FOREGROUND_DIR = 'FOREGROUND_MIXED_6NN_WITH_DEPTH'


def foreground_wire_dict(rng, virtual_per_camera=2000, real_per_camera=300, ncam=6, empty_cameras=(),
                         outside_fraction=0.05):
    """One sweep's saved dict: four keys, each a list of `ncam` float32 arrays.  `*_pixel_indices`:
    (n, 14) = u, v in ORIGINAL image pixels (1600x900), depth, 10 one-hot class dims + score;
    `*_points`: (n, 3) xyz in that sweep's LiDAR frame.  A fraction of the points lies outside the
    point-cloud range so that the range filter has work to do."""
    out = {k: [] for k in ('virtual_pixel_indices', 'real_pixel_indices', 'virtual_points', 'real_points')}
    for cam in range(ncam):
        for kind, n in (('virtual', virtual_per_camera), ('real', real_per_camera)):
            n = 0 if cam in empty_cameras else int(n * rng.uniform(0.6, 1.0))
            u = rng.uniform(0, 1599, n) if kind == 'real' else np.floor(rng.uniform(0, 1600, n))
            v = rng.uniform(0, 899, n) if kind == 'real' else np.floor(rng.uniform(0, 900, n))
            depth = rng.uniform(1.0, 60.0, n)
            onehot = np.zeros((n, 10))
            if n:
                onehot[np.arange(n), rng.integers(0, 10, n)] = 1.0
            score = rng.uniform(0.3, 1.0, (n, 1))
            out[kind + '_pixel_indices'].append(
                np.concatenate([u[:, None], v[:, None], depth[:, None], onehot, score], 1).astype(np.float32))
            xyz = rng.uniform([-50, -50, -4.5], [50, 50, 2.5], (n, 3))
            far = rng.random(n) < outside_fraction
            xyz[far] *= 1.3
            out[kind + '_points'].append(xyz.astype(np.float32))
    return out


def write_foreground_wire(root, seed=0, sweeps=10, missing_sweeps=(), **kw):
    """Writes `<root>/samples/<FOREGROUND_DIR>/*.pkl.npy` for one sample and its sweeps the way the
    reference expects to find them next to `<root>/samples/LIDAR_TOP/<file>` (path logic of :126-128) and
    returns the `results` dict the pipeline stages receive (keys of nuscenes_dataset.py:get_data_info that
    these stages read): pts_filename, timestamp [s], sweeps[{data_path, timestamp [us], sensor2lidar_*}]."""
    rng = np.random.default_rng(seed + 7000)
    fg_dir = os.path.join(root, 'samples', FOREGROUND_DIR)
    os.makedirs(fg_dir, exist_ok=True)

    def save(name, d):
        np.save(os.path.join(fg_dir, name + '.pkl.npy'), d, allow_pickle=True)
    sample_name = 'n015-sample-%04d.pcd.bin' % seed
    save(sample_name, foreground_wire_dict(rng, **kw))
    ts_us = 1533151603547590 + 500000 * seed
    results = dict(pts_filename=os.path.join(root, 'samples', 'LIDAR_TOP', sample_name), timestamp=ts_us / 1e6,
                   sweeps=[])
    for s in range(sweeps):
        name = 'n015-sweep-%04d-%02d.pcd.bin' % (seed, s)
        if s not in missing_sweeps:   # the reference skips sweeps whose file does not exist (:311,324)
            save(name, foreground_wire_dict(rng, **kw))
        ang = rng.normal(0, 0.02)
        c, sn = np.cos(ang), np.sin(ang)
        results['sweeps'].append(dict(
            data_path=os.path.join(root, 'samples', 'LIDAR_TOP', name), timestamp=ts_us - 50000 * (s + 1),
            sensor2lidar_rotation=np.array([[c, -sn, 0.0], [sn, c, 0.0], [0.0, 0.0, 1.0]]),
            sensor2lidar_translation=rng.normal(0, 0.5, 3) * np.array([1.0, 1.0, 0.05])))
    return results
