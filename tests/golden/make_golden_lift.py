"""Generates tests/golden/lift_reference.npz: outputs of the REFERENCE's `get_foreground2D` and of the
sparse depth canvas of `depth_aware_channel_compression` (MSMDFusion.py:169-238, 335-356), their method
bodies compiled from /root/reference in place (oracle/ref_lift.py).  Inputs are the seeded synthetic
scene `lift_inputs()` below rebuilds identically on the GPU box (a CRC of them is stored); the outputs
are stored in full for the 15 point columns' checksum, every 8th row of the gated features, and the
depth canvas as (linear index, value) pairs.  Runs only in the build container.

    python tests/golden/make_golden_lift.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from msmdfusion_b200 import synthetic  # noqa: E402  (input generator only)

HERE = os.path.dirname(os.path.abspath(__file__))
ROW_STEP = 8


def lift_inputs():
    """Two samples (one with an empty camera), level-1 compressed features (49 ch, stride 8), a score
    gate open for roughly half of the points.  Duplicate real pixels are forced in camera 0."""
    pts = synthetic.lidar_scene(seed=3, sweeps=1)
    metas = [synthetic.camera_scene(seed=3, lidar_points=pts, virtual_per_camera=1500, real_per_camera=2500,
                                    empty_cameras=(2,)),
             synthetic.camera_scene(seed=4, lidar_points=pts, virtual_per_camera=900, real_per_camera=2500)]
    for meta in metas:
        r = meta['foreground2D_info']['fg_real_pixels'][0]
        r[-50:, :2] = r[:50, :2]
    rng = np.random.default_rng(77)
    H, W = synthetic.INPUT_SHAPE
    feat = rng.standard_normal((12, 49, H // 8, W // 8)).astype(np.float32)
    score_w = (rng.standard_normal(49 + 17) * 0.02).astype(np.float32)
    score_b = np.float32(0.05)
    return metas, feat, score_w, score_b


def inputs_crc(metas, feat, score_w, score_b):
    crc = zlib.crc32(feat.tobytes())
    crc = zlib.crc32(score_w.tobytes(), crc)
    for m in metas:
        for k in ('fg_pixels', 'fg_points', 'fg_real_pixels'):
            for a in m['foreground2D_info'][k]:
                crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
        crc = zlib.crc32(np.ascontiguousarray(np.stack(m['lidar2img'])).tobytes(), crc)
    return crc


def main():
    from oracle import ref_lift
    metas, feat, score_w, score_b = lift_inputs()
    fg = ref_lift.get_foreground2d(feat, metas, score_w, score_b)
    H, W = synthetic.INPUT_SHAPE
    canvas = ref_lift.depth_maps([(H, W)] * 3, metas)[0].reshape(-1)   # same-size bilinear = identity
    nz = np.nonzero(canvas)[0]
    out = dict(inputs_crc=np.array([inputs_crc(metas, feat, score_w, score_b)], np.int64),
               canvas_index=nz.astype(np.int64), canvas_value=canvas[nz],
               canvas_size=np.array([canvas.shape[0]], np.int64))
    for b, a in enumerate(fg):
        out['count%d' % b] = np.array([a.shape[0]], np.int64)
        out['points_crc%d' % b] = np.array([zlib.crc32(np.ascontiguousarray(a[:, :15]).tobytes())], np.int64)
        out['rows%d' % b] = a[::ROW_STEP]
        print('sample', b, a.shape, 'gate open', float((a[:, 15:] != 0).any(1).mean()))
    np.savez_compressed(os.path.join(HERE, 'lift_reference.npz'), **out)
    print('canvas nonzeros', nz.shape[0])


if __name__ == '__main__':
    main()
