"""Debug tool (GPU): per-layer timing of the tensor-core conv variants (2: A via smem, 3: A via TMEM)
on the real rulebooks of a synthetic scene -- picks the crossover used by the auto rule."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from msmdfusion_b200 import ops, synthetic  # noqa: E402
from msmdfusion_b200 import sparse_encoder as se  # noqa: E402


def time_call(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device('cuda:0')
    cfg, layer, enc = bench.build_pipeline(dev)
    se.SparseEncoder.use_executor = False
    for sweeps in (1, 10):
        pts = torch.from_numpy(synthetic.lidar_scene(0, sweeps)).to(dev)
        ops.PROFILE = []
        with torch.no_grad():
            mean, coors, _ = layer.forward_mean(pts, 5, batch_idx=0)
            enc(mean, coors, 1)
        torch.cuda.synchronize()
        recs = [r for r in ops.PROFILE if r['op'] == 'spconv_fwd']
        ops.PROFILE = None
        seen = set()
        print('--- sweeps', sweeps)
        for r in recs:
            key = (r['cin'], r['cout'], r['kvol'], r['n_out'], r['residual'])
            if key in seen:
                continue
            seen.add(key)
            pair = r['pair']
            feat = torch.randn(r['n_in'], r['cin'], device=dev)
            w = torch.randn(r['cout'], r['kvol'], 1, 1, r['cin'], device=dev) * 0.05
            tcw = ops.pack_weight_tc(w)
            res = torch.randn(r['n_out'], r['cout'], device=dev) if r['residual'] else None
            sc = torch.ones(r['cout'], device=dev)
            sh = torch.zeros(r['cout'], device=dev)
            out = {}
            for v in (2, 3):
                ops.set_tc_variant(v)
                out[v] = time_call(lambda: ops.spconv_fwd_tc(feat, tcw, pair, sc, sh, res, True))
            ops.set_tc_variant(0)
            auto = time_call(lambda: ops.spconv_fwd_tc(feat, tcw, pair, sc, sh, res, True))
            print('%3d->%3d k=%2d n_out=%6d res=%d  v2=%.4f v3=%.4f auto=%.4f ms' %
                  (r['cin'], r['cout'], r['kvol'], r['n_out'], int(r['residual']), out[2], out[3], auto))


if __name__ == '__main__':
    main()
