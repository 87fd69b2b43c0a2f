"""Debug tool (GPU): host-side cProfile of the LC voxel-space step (where do the CPU milliseconds go?)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    cfg, det, pts_np, meta, fpn = bench.build_lc_pipeline(dev, 0, 'S')
    pts = torch.from_numpy(pts_np).to(dev)

    def step():
        with torch.no_grad():
            return det.extract_voxel_space([pts], fpn, [meta])
    for _ in range(5):
        out = step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(10):
        out = step()
    torch.cuda.synchronize()
    print('ms/step (wall, 10 steps)', (time.perf_counter() - t0) * 100)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        out = step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('cumulative').print_stats(45)
    st.sort_stats('tottime').print_stats(30)


if __name__ == '__main__':
    main()
