"""Diagnostic: depth_aware_channel_compression on the GPU (cuDNN default / TF32 off / cuDNN off) against
the same torch modules on the CPU, per FPN level.  Prints max|diff| / max|ref|."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import _fixtures  # noqa: E402


def main():
    det_cpu, _ = _fixtures.build_msmd_detector(1)
    det, _ = _fixtures.build_msmd_detector(1, torch.device('cuda:0'))
    scenes, metas, fpn = _fixtures.lc_scene(2)
    from oracle import model as omodel
    ref = omodel.depth_aware_channel_compression(det_cpu.state_dict(), fpn, metas)
    feats = [torch.from_numpy(f).cuda() for f in fpn]

    def run(tag):
        with torch.no_grad():
            out = det.depth_aware_channel_compression(feats, metas)
        torch.cuda.synchronize()
        errs = [float(np.abs(o.cpu().numpy() - r).max() / max(1.0, np.abs(r).max())) for o, r in zip(out, ref)]
        print(tag, ' '.join('%.3e' % e for e in errs), flush=True)
    run('cudnn default       ')
    torch.backends.cudnn.allow_tf32 = False
    run('cudnn allow_tf32=0  ')
    torch.backends.cudnn.deterministic = True
    run('cudnn deterministic ')
    torch.backends.cudnn.deterministic = False
    with torch.backends.cudnn.flags(enabled=False):
        run('cudnn disabled      ')
    print('matmul tf32', torch.backends.cuda.matmul.allow_tf32)


if __name__ == '__main__':
    main()
