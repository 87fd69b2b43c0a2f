"""GPU tool: cProfile of the HOST side of LC steps (where does the Python time go?).
    python tools/lc_hostprofile.py [--steps 20] [--top 45]"""
import argparse
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--top', type=int, default=45)
    ap.add_argument('--precision', default='bf16x3c')
    args = ap.parse_args()
    import torch
    import bench
    from msmdfusion_b200 import spconv
    spconv.CONV_PRECISION = args.precision
    dev = torch.device('cuda:0')
    cfg, det, pts_np, meta, fpn = bench.build_lc_pipeline(dev, 0, 'S')
    pts = torch.from_numpy(pts_np).to(dev)
    metas = [meta]

    def step():
        with torch.no_grad():
            return det.extract_voxel_space([pts], fpn, metas)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('tottime').print_stats(args.top)
    st.sort_stats('cumulative').print_stats(args.top)


if __name__ == '__main__':
    main()
