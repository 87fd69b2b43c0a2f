"""GPU tool: time furthest point sampling (m = 2048) on voxel-coordinate clouds of the LC scene's sizes with the
bucketed exact algorithm (default) and the brute-force kernels, and check that they agree.
    python tools/fps_bench.py [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    from msmdfusion_b200 import ops
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(0)

    def cloud(n):
        # clustered integer voxel coordinates (the only-2D voxels of a scene: objects + ground patches)
        centres = np.stack([rng.integers(0, 41, 200), rng.integers(0, 1440, 200), rng.integers(0, 1440, 200)], 1)
        pts = centres[rng.integers(0, 200, 3 * n)] + np.round(rng.normal(0, 1, (3 * n, 3)) * np.array([1.5, 25, 25])).astype(np.int64)
        pts = pts[(pts[:, 0] >= 0) & (pts[:, 0] < 41) & (pts[:, 1] >= 0) & (pts[:, 1] < 1440) & (pts[:, 2] >= 0) & (pts[:, 2] < 1440)]
        _, first = np.unique(pts, axis=0, return_index=True)
        return pts[np.sort(first)][:n].astype(np.float32)

    def timed(fn, it=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it, out

    rows = []
    for n in (1500, 3249, 8000, 17567, 46831, 65000):
        xyz = torch.from_numpy(cloud(n)).to(dev)
        n = xyz.shape[0]
        m = min(2048, n)
        ops.check(ops.lib().msmd_fps_set_algorithm(1), 'set')
        t_brute, a = timed(lambda: ops.furthest_point_sample_single(xyz, m))
        ops.check(ops.lib().msmd_fps_set_algorithm(0), 'set')
        t_bucket, b = timed(lambda: ops.furthest_point_sample_single(xyz, m))
        same = bool(torch.equal(a, b))
        print('n=%6d m=%4d  brute %.3f ms (%.2f us/round)   bucketed %.3f ms (%.2f us/round)  x%.2f  %s' % (
            n, m, t_brute, 1e3 * t_brute / (m - 1), t_bucket, 1e3 * t_bucket / (m - 1), t_brute / t_bucket,
            'identical' if same else 'MISMATCH'))
        rows.append(dict(n=n, m=m, brute_ms=t_brute, bucketed_ms=t_bucket, identical=same))
    if args.json:
        json.dump(rows, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
