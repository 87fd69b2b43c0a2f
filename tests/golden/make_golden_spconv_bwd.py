"""Generates tests/golden/spconv1xbwd_*.npz: the REFERENCE's own vendored spconv-1.x backward
(``indice_conv_backward_fp32``, mmdet3d/ops/spconv/include/spconv/spconv_ops.h:364-457; compiled
unmodified from /root/reference by oracle/ref_spconv.py) on the inputs of the forward fixtures
``spconv1x_<name>.npz``.  Only the seeded output gradient (as float16-representable values, so it can
be stored compactly) and the reference's two results are stored; inputs and weights are read from the
forward fixture of the same name.

    python tests/golden/make_golden_spconv_bwd.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_spconv  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ('subm_5_16', 'conv_s2_16_32', 'conv_out_311_128_128', 'lidar_subm_16_16')


def grad_out_for(name, n_out, cout):
    """Seeded output gradient, exactly representable in float16 (stored as such)."""
    rng = np.random.default_rng(abs(hash_name(name)))
    return rng.standard_normal((n_out, cout)).astype(np.float16)


def hash_name(name):
    import zlib
    return zlib.crc32(name.encode())


def main():
    for name in CASES:
        g = np.load(os.path.join(HERE, f'spconv1x_{name}.npz'))
        idx = g['indices'].astype(np.int32)
        ks, st, pd = ([int(x) for x in g[k]] for k in ('ksize', 'stride', 'padding'))
        subm = bool(int(g['subm']))
        go = grad_out_for(name, g['out_features'].shape[0], g['out_features'].shape[1])
        gi, gw = ref_spconv.conv_backward(idx, g['features'], g['weight_krsc'], go.astype(np.float32),
                                          [int(s) for s in g['spatial_shape']], int(g['batch_size']), ks, st, pd,
                                          1, subm)
        np.savez_compressed(os.path.join(HERE, f'spconv1xbwd_{name}.npz'), grad_out=go,
                            grad_features=gi.astype(np.float32), grad_weight=gw.astype(np.float32))
        print(name, go.shape, 'max|gi|', float(np.abs(gi).max()), 'max|gw|', float(np.abs(gw).max()))


if __name__ == '__main__':
    main()
