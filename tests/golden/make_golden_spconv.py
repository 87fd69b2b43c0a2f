"""Generates tests/golden/spconv1x_*.npz by running the REFERENCE's own vendored spconv-1.x CPU ops
(mmdet3d/ops/spconv, compiled unmodified from /root/reference by oracle/ref_spconv.py) on seeded
sparse tensors.  Runs only in the build container; the fixtures are committed so the GPU box never
needs the reference.  Inputs are stored (seeded but explicit) together with the reference outputs
(strided-conv rows sorted by linear index; see oracle/ref_spconv.py).

    python tests/golden/make_golden_spconv.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from msmdfusion_b200 import synthetic  # noqa: E402  (input generator only)
from oracle import cpu, ref_spconv  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rand_sparse(rng, batch, shape, n, c):
    D, H, W = shape
    cells = batch * D * H * W
    lin = rng.choice(cells, size=min(n, cells), replace=False)
    b, r = lin // (D * H * W), lin % (D * H * W)
    z, r = r // (H * W), r % (H * W)
    idx = np.stack([b, z, r // W, r % W], 1).astype(np.int32)
    return idx, rng.standard_normal((idx.shape[0], c)).astype(np.float32)


def lidar_voxels(seed, n):
    """Real hot-path geometry: voxel coordinates of a synthetic sweep in the [41,1440,1440] grid."""
    pts = synthetic.lidar_scene(seed=seed, sweeps=1)[:n]
    _, coors, _ = cpu.hard_voxelize(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    return np.concatenate([np.zeros((coors.shape[0], 1), np.int32), coors], 1).astype(np.int32)


# name: (indices/features builder, cin, cout, ksize, stride, padding, subm)
def cases():
    rng = np.random.default_rng(7)
    out = {}
    i, f = rand_sparse(rng, 2, [9, 20, 20], 900, 5)
    out['subm_5_16'] = (i, f, [9, 20, 20], 2, 16, 3, 1, 1, True)
    i, f = rand_sparse(rng, 2, [9, 20, 20], 900, 16)
    out['conv_s2_16_32'] = (i, f, [9, 20, 20], 2, 32, 3, 2, 1, False)
    i, f = rand_sparse(rng, 1, [11, 24, 24], 700, 32)
    out['conv_s2_pad011_32_64'] = (i, f, [11, 24, 24], 1, 64, 3, 2, [0, 1, 1], False)
    i, f = rand_sparse(rng, 1, [5, 16, 16], 300, 128)
    out['conv_out_311_128_128'] = (i, f, [5, 16, 16], 1, 128, [3, 1, 1], [2, 1, 1], 0, False)
    idx = lidar_voxels(9, 2500)
    out['lidar_subm_16_16'] = (idx, rng.standard_normal((idx.shape[0], 16)).astype(np.float32),
                               [41, 1440, 1440], 1, 16, 3, 1, 1, True)
    out['lidar_conv_s2_16_32'] = (idx, rng.standard_normal((idx.shape[0], 16)).astype(np.float32),
                                  [41, 1440, 1440], 1, 32, 3, 2, 1, False)
    return out, rng


def main():
    cs, rng = cases()
    for name, (idx, feat, shape, batch, cout, ksize, stride, padding, subm) in cs.items():
        ks = ksize if isinstance(ksize, list) else [ksize] * 3
        cin = feat.shape[1]
        w = (rng.standard_normal((cout, *ks, cin)) / np.sqrt(cin * np.prod(ks) * 0.2)).astype(np.float32)
        oi, of, oshape = ref_spconv.conv(idx, feat, w, shape, batch, ksize, stride, padding, 1, subm)
        np.savez_compressed(
            os.path.join(HERE, f'spconv1x_{name}.npz'), indices=idx.astype(np.int16), features=feat.astype(np.float32),
            weight_krsc=w, spatial_shape=np.int32(shape), batch_size=np.int32(batch),
            ksize=np.int32(ks), stride=np.int32(stride if isinstance(stride, list) else [stride] * 3),
            padding=np.int32(padding if isinstance(padding, list) else [padding] * 3), subm=np.int32(subm),
            out_indices=oi.astype(np.int16), out_features=of.astype(np.float32), out_shape=np.int32(oshape))
        print(name, idx.shape, '->', oi.shape, 'max|out|', float(np.abs(of).max()))


if __name__ == '__main__':
    main()
