"""Debug probe (GPU): per-launch CUDA-event times of the first sparse convs of the LiDAR encoder
over several steps, to localise launch-side stalls."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from msmdfusion_b200 import ops, synthetic  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    cfg, layer, enc = bench.build_pipeline(dev)
    pts = torch.from_numpy(synthetic.lidar_scene(0, 1)).to(dev)

    def step():
        with torch.no_grad():
            mean, coors, _ = layer.forward_mean(pts, 5, batch_idx=0)
            return enc(mean, coors, 1)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    for mode in ('plain', 'sync_before_each_op'):
        ops.PROFILE_SYNC = mode != 'plain'
        for it in range(4):
            ops.PROFILE = []
            step()
            torch.cuda.synchronize()
            recs = ops.PROFILE
            ops.PROFILE = None
            line = []
            for r in recs:
                ms = r['start'].elapsed_time(r['end'])
                if r['op'] == 'spconv_fwd':
                    line.append('%d>%d:%.3f' % (r['cin'], r['cout'], ms))
                elif ms > 0.2:
                    line.append('%s:%.3f' % (r['op'], ms))
            print(mode, it, ' '.join(line[:12]))
    ops.PROFILE_SYNC = False
    # whole-step wall/device time without instrumentation
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print('uninstrumented ms/step', e0.elapsed_time(e1) / 10)


if __name__ == '__main__':
    main()
