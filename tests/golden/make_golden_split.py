"""Generates tests/golden/split_*.npz: inputs and outputs of the REFERENCE's voxel_modality_split
method and its numba `type_assign`, both compiled from /root/reference in place and run on CPU tensors
(oracle/ref_split.py).  Runs only in the build container; the fixtures are committed
so the GPU box never needs the reference.

    python tests/golden/make_golden_split.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from msmdfusion_b200 import synthetic  # noqa: E402  (input generator only)
from oracle import cpu, ref_split  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rand_coors(rng, shape, n):
    D, H, W = shape
    lin = rng.choice(D * H * W, size=n, replace=False)
    return np.stack([lin // (H * W), lin % (H * W) // W, lin % W], 1).astype(np.int32)


def dense_overlap(rng):
    """Small grid, heavy overlap, duplicated 2D rows, float-key collisions (z >= 17: x and x+1 share a key)."""
    i3 = rand_coors(rng, [41, 200, 200], 6000)
    i2 = rand_coors(rng, [41, 200, 200], 5000)
    i2[:2500] = i3[rng.choice(6000, 2500, replace=False)]
    high = np.nonzero(i3[:, 0] >= 17)[0][:200]
    i2[2500:2700] = i3[high] + np.array([0, 0, 1], np.int32)
    i2[2700:2720] = i2[2600:2620]
    return i3, i2[rng.permutation(i2.shape[0])]


def lidar_grid(rng):
    """Real hot-path geometry: LiDAR voxels of a synthetic sweep vs a jittered copy (the virtual points)."""
    pts = synthetic.lidar_scene(seed=11, sweeps=1)[:60000]
    _, c3, _ = cpu.hard_voxelize(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    jit = pts[rng.choice(pts.shape[0], 20000, replace=False)].copy()
    jit[:, :3] += rng.normal(0, 0.05, (jit.shape[0], 3)).astype(np.float32)
    _, c2, _ = cpu.hard_voxelize(jit, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    return c3.astype(np.int32), c2.astype(np.int32)


def disjoint(rng):
    i3 = rand_coors(rng, [8, 64, 64], 900)
    i3 = i3[i3[:, 2] % 2 == 0]
    i2 = rand_coors(rng, [8, 64, 64], 700)
    return i3, i2[i2[:, 2] % 2 == 1]


CASES = {'split_dense_overlap': dense_overlap, 'split_lidar_grid': lidar_grid, 'split_disjoint': disjoint}


def main():
    for k, (name, fn) in enumerate(CASES.items()):
        c3, c2 = fn(np.random.default_rng(100 + k))
        mix3, mix2, syn3, syn2 = ref_split.split_single(c3, c2)
        z = lambda c: np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], 1)  # noqa: E731
        np.savez_compressed(os.path.join(HERE, name + '.npz'), indices3=z(c3), indices2=z(c2),
                            mix3=mix3, mix2=mix2, syn3=syn3.astype(np.int64), syn2=syn2.astype(np.int64))
        print(name, c3.shape[0], c2.shape[0], 'mixed', int(mix3.sum()), int(mix2.sum()))


if __name__ == '__main__':
    main()
