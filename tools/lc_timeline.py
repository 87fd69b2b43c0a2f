"""GPU tool: poor man's timeline of one LC step (no nsys in the image).  Wraps the phases of
MSMDFusionDetector.extract_voxel_space and records, per phase, the HOST interval (perf_counter) and the GPU interval
(CUDA events on the stream the phase was issued on), all relative to the start of the step.  Shows whether the step
is bound by the host issuing work or by a GPU dependency chain, and what overlaps with what.

    python tools/lc_timeline.py [--precision bf16x3c] [--steps 3] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='bf16x3c')
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    import torch
    import bench
    from msmdfusion_b200 import fusion_encoder as fe
    from msmdfusion_b200 import ops, spconv
    spconv.CONV_PRECISION = args.precision
    dev = torch.device('cuda:0')
    cfg, det, pts_np, meta, fpn = bench.build_lc_pipeline(dev, 0, 'S')
    pts = torch.from_numpy(pts_np).to(dev)
    metas = [meta]
    LOG = []
    T0 = [0.0, None]

    def wrap(obj, name, label=None):
        fn = getattr(obj, name)
        lab = label or name

        def inner(*a, **k):
            st = torch.cuda.current_stream(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0 = time.perf_counter()
            e0.record(st)
            out = fn(*a, **k)
            e1.record(st)
            LOG.append((lab, h0, time.perf_counter(), e0, e1, st.cuda_stream))
            return out
        setattr(obj, name, inner)

    wrap(det, 'depth_aware_channel_compression', 'compression (cuDNN)')
    wrap(det, 'voxelize_mean', 'voxelize+mean')
    wrap(det.pts_middle_encoder, 'forward', 'LiDAR SparseEncoder (executor)')
    wrap(det, 'fetch_2D_voxels', 'lift+voxelize (scale)')
    wrap(det, 'voxel_modality_split', 'modality split (scale)')
    wrap(det.multimodal_middle_encoder, '_assign_b1', 'FPS/NN/ball-query chain (stage)')
    wrap(det.multimodal_middle_encoder, '_grouped_sparse_conv_b1', 'GMA stage: gates+only3D+aggregation')
    wrap(det.multimodal_middle_encoder, 'forward', 'GMA encoder (whole)')
    wrap(fe.Fsp, 'sparse_add', 'sparse_add')
    wrap(ops, 'furthest_point_sample_single', '  fps')
    wrap(ops, 'ball_query_single', '  ball_query')
    wrap(ops, 'nn_search', '  nn_search')

    def step():
        with torch.no_grad():
            return det.extract_voxel_space([pts], fpn, metas)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    report = []
    for it in range(args.steps):
        LOG.clear()
        torch.cuda.synchronize()
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        s0.record()
        bev, _ = step()
        s1.record()
        h_issue = time.perf_counter()
        torch.cuda.synchronize()
        h_done = time.perf_counter()
        rows = []
        for lab, a, b, e0, e1, st in LOG:
            rows.append(dict(phase=lab, stream=hex(st)[-5:], host_start=(a - h0) * 1e3, host_end=(b - h0) * 1e3,
                             gpu_start=s0.elapsed_time(e0), gpu_end=s0.elapsed_time(e1)))
        rows.sort(key=lambda r: r['host_start'])
        total = s0.elapsed_time(s1)
        print('--- step %d: GPU %.3f ms, host issue %.3f ms, host until idle %.3f ms' % (
            it, total, (h_issue - h0) * 1e3, (h_done - h0) * 1e3))
        print('%-44s %6s | %8s %8s | %8s %8s %8s' % ('phase', 'stream', 'host beg', 'host end', 'gpu beg', 'gpu end', 'gpu dur'))
        for r in rows:
            print('%-44s %6s | %8.3f %8.3f | %8.3f %8.3f %8.3f' % (r['phase'], r['stream'], r['host_start'], r['host_end'],
                                                               r['gpu_start'], r['gpu_end'], r['gpu_end'] - r['gpu_start']))
        report.append(dict(gpu_ms=total, host_issue_ms=(h_issue - h0) * 1e3, phases=rows))
    if args.json:
        json.dump(report, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
