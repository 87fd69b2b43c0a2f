"""SURVEY §8(f) rank 3 measurement: the foreground-2D test pipeline (4 stages,
configs/MSMDFusion_nusc_voxel_LC.py:97-103) on one sample + 10 sweeps of synthetic wire files at
config-3 size (about 60 k virtual points per sample after the merge) -- the reference's own classes run
in place (oracle/ref_loading.py) beside msmdfusion_b200.loading, same files (page cache warm), one host
thread, median of `--reps` runs.  CPU only; needs /root/reference for the reference leg.

    python tests/tools/bench_loader.py [--reps 20] [--virtual 900] [--real 180]
"""
import argparse
import copy
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from msmdfusion_b200 import loading, synthetic  # noqa: E402


def median_ms(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t)), float(np.min(t))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--virtual', type=int, default=1100)
    ap.add_argument('--real', type=int, default=220)
    args = ap.parse_args()
    torch.set_num_threads(1)
    os.chdir(tempfile.mkdtemp())
    res = synthetic.write_foreground_wire('data', seed=0, sweeps=10, virtual_per_camera=args.virtual,
                                          real_per_camera=args.real)
    res.update(scale_factor=np.array([0.5, 0.49777778, 0.5, 0.49777778], np.float32), img_shape=(448, 800, 3),
               transformation_3d_flow=['R', 'S', 'T'], pcd_rotation=torch.eye(3), pcd_scale_factor=1.0,
               pcd_trans=np.zeros(3), flip=False)
    ours = loading.build_pipeline([
        dict(type='LoadForeground2D'), dict(type='LoadForeground2DFromMultiSweeps', sweeps_num=10, test_mode=True),
        dict(type='GlobalRotTransFilterForeground2D', point_cloud_range=synthetic.POINT_CLOUD_RANGE),
        dict(type='ImgScaleCropFlipForeground2D')])
    out = loading.run_pipeline(ours, copy.copy(res))
    scene = out['foreground2D_info']['packed']
    paths = [loading.foreground_path(res['pts_filename'])] + [loading.foreground_path(s['data_path']) for s in res['sweeps']]
    line = {'workload': 'foreground-2D test pipeline, 1 key frame + 10 sweeps, 6 cameras',
            'rows_after_merge_and_filter': int(scene.offsets[-1]), 'real_rows': int(scene.real_offsets[-1]),
            'wire_bytes': int(sum(os.path.getsize(p) for p in paths)), 'host_threads': 1, 'reps': args.reps}
    line['read_files_ms'] = median_ms(lambda: [loading.read_wire(p) for p in paths], args.reps)[0]
    line['ours_ms'], line['ours_min_ms'] = median_ms(lambda: loading.run_pipeline(ours, copy.copy(res)), args.reps)
    try:
        from oracle import ref_loading
        have_ref = ref_loading.available()
    except Exception:
        have_ref = False
    if have_ref:
        stages = ref_loading.test_pipeline(synthetic.POINT_CLOUD_RANGE)
        ref = ref_loading.run(stages, copy.copy(res))
        same = all(np.array_equal(a, b) for a, b in zip(ref['foreground2D_info']['fg_pixels'],
                                                        out['foreground2D_info']['fg_pixels']))
        same = same and all(torch.equal(a.tensor, b.tensor) for a, b in zip(ref['foreground2D_info']['fg_points'],
                                                                            out['foreground2D_info']['fg_points']))
        line['identical_to_reference'] = bool(same)
        line['reference_ms'], line['reference_min_ms'] = median_ms(lambda: ref_loading.run(stages, copy.copy(res)),
                                                                   args.reps)
        line['speedup'] = round(line['reference_ms'] / line['ours_ms'], 2)
        line['speedup_excluding_file_read'] = round((line['reference_ms'] - line['read_files_ms'])
                                                    / max(1e-9, line['ours_ms'] - line['read_files_ms']), 2)
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in line.items()}))


if __name__ == '__main__':
    main()
