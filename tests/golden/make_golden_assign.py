"""Generates tests/golden/assign_*.npz: inputs and outputs of the REFERENCE's `fps_NN_fast`
(sparse_multimodal_encoder_painting.py:276-323), its method body compiled from /root/reference in
place and run by torch on the CPU with the two CUDA-only ops bound to the pinned C restatements
(oracle/ref_assign.py).  Parameters are the four scales of configs/MSMDFusion_nusc_voxel_LC.py:146-149.
Runs only in the build container; the fixtures are committed so the GPU box never needs the reference.

    python tests/golden/make_golden_assign.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_assign  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def voxel_cloud(rng, n, shape, clusters, spread):
    """Unique integer (b=0,z,y,x) voxel coordinates clustered like foreground objects."""
    D, H, W = shape
    centres = np.stack([rng.integers(0, D, clusters), rng.integers(0, H, clusters), rng.integers(0, W, clusters)], 1)
    pts = centres[rng.integers(0, clusters, 4 * n)] + \
        np.round(rng.normal(0, 1, (4 * n, 3)) * np.array([1.5, spread, spread])).astype(np.int64)
    ok = np.all((pts >= 0) & (pts < np.array(shape)), 1)
    pts = pts[ok]
    _, first = np.unique(pts, axis=0, return_index=True)
    pts = pts[np.sort(first)][:n]
    return np.concatenate([np.zeros((pts.shape[0], 1), np.int64), pts], 1).astype(np.int32)


# name: (grid, queries, keys, fps_num, radius, max_cluster_samples, dist_thresh)
CASES = {
    'assign_scale0': ([41, 1440, 1440], 12000, 9000, 2048, 6, 200, 13.3),
    'assign_scale1': ([21, 720, 720], 7000, 6000, 2048, 3, 100, 6.6),
    'assign_scale2': ([11, 360, 360], 4000, 3000, 2048, 2, 50, 3.3),
    'assign_scale3': ([5, 180, 180], 2500, 2000, 2048, 1, 25, 1.6),
    'assign_direct': ([11, 360, 360], 1500, 2500, 2048, 2, 50, 3.3),   # Q <= fps_num: brute-force branch
}


def main():
    for k, (name, (shape, nq, nk, fps, rad, ns, th)) in enumerate(CASES.items()):
        rng = np.random.default_rng(500 + k)
        both = voxel_cloud(rng, nq + nk, shape, 14, 9.0)   # same objects seen by both modalities
        both = both[rng.permutation(both.shape[0])]
        q, key = both[:nq], both[nq:]
        out = ref_assign.fps_nn_fast(q, key, fps, rad, ns, th)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), query=q, key=key, assign=out,
                            params=np.array([fps, rad, ns, th], np.float64))
        print(name, q.shape[0], key.shape[0], 'assigned', float((out >= 0).mean()))


if __name__ == '__main__':
    main()
