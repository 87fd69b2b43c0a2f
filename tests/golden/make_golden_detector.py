"""Generates tests/golden/detector_reference.npz: the output of the REFERENCE's own
`MSMDFusionDetector.extract_pts_feat` up to the tensor `bev_fusion` consumes (MSMDFusion.py:421-445), run
in place by oracle/ref_detector.py -- the reference's methods and classes compiled from /root/reference,
its own C++ CPU voxelizer, the oracle's conv restatement inside the two sparse encoders -- on the seeded
batch-2 scene and the seeded random-init detector of tests/_fixtures.py.  The GPU box rebuilds the same
weights and inputs (CRCs stored) and checks the CUDA path against what is stored here: per stage the
voxel count, a CRC of the index tensor and every 128th feature row; of the (2, 640, 180, 180) BEV tensor
40 000 seeded positions plus its non-zero count.  Runs only in the build container (≈2 min).

    python tests/golden/make_golden_detector.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import _fixtures  # noqa: E402

BATCH, DET_SEED, DUMMY_SEED, ROW_STEP, BEV_SAMPLES = 2, 1, 77, 128, 40000


def bev_positions(size):
    return np.random.default_rng(4242).choice(size, BEV_SAMPLES, replace=False)


def inputs_crc(scenes, metas, fpn):
    crc = 0
    for a in list(scenes) + list(fpn):
        crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
    for m in metas:
        for k in ('fg_pixels', 'fg_points', 'fg_real_pixels'):
            for a in m['foreground2D_info'][k]:
                crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
    return crc


def main():
    from oracle import ref_detector
    det, cfg = _fixtures.build_msmd_detector(DET_SEED)
    sd = det.state_dict()
    scenes, metas, fpn = _fixtures.lc_scene(BATCH)
    bev, outs = ref_detector.extract_voxel_space(sd, cfg, scenes, fpn, metas, DUMMY_SEED)
    flat = bev.reshape(-1)
    pos = bev_positions(flat.shape[0])
    out = dict(weights_crc=np.array([_fixtures.state_dict_crc(sd)], np.int64),
               inputs_crc=np.array([inputs_crc(scenes, metas, fpn)], np.int64),
               bev_shape=np.array(bev.shape, np.int64), bev_values=flat[pos],
               bev_nonzero=np.array([np.count_nonzero(flat)], np.int64),
               bev_absmax=np.array([np.abs(flat).max()], np.float32))
    for i, o in enumerate(outs):
        out['count%d' % i] = np.array([o.indices.shape[0]], np.int64)
        out['shape%d' % i] = np.array(o.spatial_shape, np.int64)
        out['indices_crc%d' % i] = np.array([zlib.crc32(np.ascontiguousarray(o.indices, np.int32).tobytes())], np.int64)
        out['rows%d' % i] = o.features[::ROW_STEP]
        print('stage', i, o.features.shape, o.spatial_shape)
    np.savez_compressed(os.path.join(HERE, 'detector_reference.npz'), **out)
    print('bev', bev.shape, 'nonzero', int(out['bev_nonzero'][0]), 'absmax', float(out['bev_absmax'][0]))


if __name__ == '__main__':
    main()
