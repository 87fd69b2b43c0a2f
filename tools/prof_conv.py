"""GPU tool: run the LiDAR SparseEncoder module by module (one C-ABI call per convolution) so that ncu can
capture the conv launches of ONE scene:   ncu --set full --import-source on -k regex:spconv_fwd -s 21 -c 21 ...
    python tools/prof_conv.py [--precision tf32x3|bf16x3] [--sweeps 1] [--passes 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default=None)
    ap.add_argument('--sweeps', type=int, default=1)
    ap.add_argument('--passes', type=int, default=2)
    ap.add_argument('--lc', action='store_true', help='the LC voxel-space path (37 conv launches per pass) instead of the LiDAR encoder (21)')
    args = ap.parse_args()
    import torch
    import bench
    from msmdfusion_b200 import spconv, synthetic
    from msmdfusion_b200 import sparse_encoder as se
    if args.precision:
        spconv.CONV_PRECISION = args.precision
    dev = torch.device('cuda:0')
    se.SparseEncoder.use_executor = False
    if args.lc:
        from msmdfusion_b200 import fusion_encoder as fe
        fe.SparseMultiModalEncoderPaint.use_executor = False
        cfg, det, pts_np, meta, fpn = bench.build_lc_pipeline(dev, 0, 'S' if args.sweeps == 1 else 'L')
        pts = torch.from_numpy(pts_np).to(dev)
        with torch.no_grad():
            for _ in range(args.passes):
                bev, outs = det.extract_voxel_space([pts], fpn, [meta])
        torch.cuda.synchronize()
        print('ok', int(outs[0].indices.shape[0]), 'voxels at stage 1')
        return
    cfg, layer, enc = bench.build_pipeline(dev)
    pts = torch.from_numpy(synthetic.lidar_scene(0, args.sweeps)).to(dev)
    with torch.no_grad():
        for _ in range(args.passes):
            mean, coors, _ = layer.forward_mean(pts, 5, batch_idx=0)
            enc(mean, coors, 1)
    torch.cuda.synchronize()
    print('ok', int(coors.shape[0]), 'voxels')


if __name__ == '__main__':
    main()
