"""Generates tests/golden/*.npz by running the REFERENCE's own CPU ``hard_voxelize``
(compiled from /root/reference by oracle/build.py into oracle/_ref) on seeded synthetic
inputs.  Runs only in the build container (needs /root/reference); the fixtures it writes
are committed so that the GPU box never needs the reference.

    python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from msmdfusion_b200 import synthetic  # noqa: E402  (input generator only)
from oracle import cpu  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (points generator, voxel_size, range, max_points, max_voxels)
    'lidar_s': (lambda: synthetic.lidar_scene(seed=3, sweeps=1)[:12000], synthetic.VOXEL_SIZE,
                synthetic.POINT_CLOUD_RANGE, 10, 160000),
    'lidar_overflow': (lambda: synthetic.lidar_scene(seed=4, sweeps=1)[:12000], synthetic.VOXEL_SIZE,
                       synthetic.POINT_CLOUD_RANGE, 3, 2500),
    'coarse_c64': (lambda: synthetic.random_points(6000, 64, seed=5), [0.6, 0.6, 1.6],
                   synthetic.POINT_CLOUD_RANGE, 10, 160000),
}


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def main():
    for name, (gen, vs, rng, mp, mv) in CASES.items():
        pts = gen()
        voxels, coors, num = cpu.hard_voxelize_ref(pts, vs, rng, mp, mv)
        np.savez_compressed(
            os.path.join(HERE, f'voxelize_{name}.npz'), points_crc=crc(pts), coors=coors.astype(np.int16),
            num_points=num.astype(np.int8), voxels_crc=crc(voxels), voxel_num=np.int32(coors.shape[0]),
            voxel_size=np.float64(vs), coors_range=np.float64(rng), max_points=np.int32(mp),
            max_voxels=np.int32(mv))
        print(name, pts.shape, coors.shape, int(num.sum()))


if __name__ == '__main__':
    main()
