"""GPU tool: per-layer timing of every prepared configuration of the tensor-core sparse convolution on the real
rulebooks of the bench scenes -- the table round 2 needs to decide defaults (precision mode, kernel variant, chunk
blocks per stage, mask-sorted tiles, occupancy) per layer class.

    python tools/autotune_conv.py [--sweeps 1,10] [--iters 30] [--json gpurun_out/autotune.json]

For each distinct (Cin, Cout, kernel volume, N_out, residual) of the LiDAR SparseEncoder it times, with CUDA events
after warm-up (L2 warm: the layer's own working set, as inside the network where the previous layer just wrote it):

    tf32x3            the default (variant chosen by N, split-K pairs where the rule says so)
    tf32x3 v2 / v3    forced variants
    bf16x3            16-bit kernel, variant 2 (A in shared memory)          [+ cps=2: two chunk blocks per stage]
    bf16x3 v3         16-bit kernel, variant 3 (A in tensor memory)          [only meaningful if quick_gpu_check says ok]
    ... each of them on the plain and on the mask-sorted table (3x3x3 SubM layers)
    occ=1 / occ=2     one / two CTAs per SM forced (tf32x3 and bf16x3)

and prints one line per layer with the best configuration and its gain over the default.  Results are checked
against the default kernel's output (<= 2e-5 of the tensor's scale) so that a configuration that is fast because it
is wrong cannot win.
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sweeps', default='1,10')
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--json', default=None)
    ap.add_argument('--with-v3-16bit', action='store_true', help='include variant 3 of the 16-bit kernel')
    args = ap.parse_args()
    import torch
    import bench
    from msmdfusion_b200 import ops, synthetic
    from msmdfusion_b200 import sparse_encoder as se
    dev = torch.device('cuda:0')
    L = ops.lib()

    def tune(**kw):
        for key, name in enumerate(('occ', 'stages', 'split', 'cps')):
            ops.check(L.msmd_spconv_tc_set_tuning(key, int(kw.get(name, 0))), 'msmd_spconv_tc_set_tuning')

    def time_call(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters * 1e3   # microseconds

    cfg, layer, enc = bench.build_pipeline(dev)
    se.SparseEncoder.use_executor = False
    report = []
    for sweeps in [int(x) for x in args.sweeps.split(',')]:
        pts = torch.from_numpy(synthetic.lidar_scene(0, sweeps)).to(dev)
        ops.PROFILE = []
        with torch.no_grad():
            mean, coors, _ = layer.forward_mean(pts, 5, batch_idx=0)
            enc(mean, coors, 1)
        torch.cuda.synchronize()
        recs = [r for r in ops.PROFILE if r['op'] == 'spconv_fwd']
        ops.PROFILE = None
        seen = set()
        print('=== %d sweep(s)' % sweeps)
        for r in recs:
            key = (r['cin'], r['cout'], r['kvol'], r['n_out'], r['residual'])
            if key in seen:
                continue
            seen.add(key)
            pair = r['pair']
            feat = torch.randn(r['n_in'], r['cin'], device=dev)
            w = torch.randn(r['cout'], r['kvol'], 1, 1, r['cin'], device=dev) * 0.05
            res = torch.randn(r['n_out'], r['cout'], device=dev) if r['residual'] else None
            sc, sh = torch.ones(r['cout'], device=dev), torch.zeros(r['cout'], device=dev)
            sortable = r['kvol'] == 27 and r['n_in'] == r['n_out']
            perm, pair_s = ops.rulebook_mask_sort(pair) if sortable else (None, None)
            packed = {m: ops.pack_weight_tc(w, ops.TC_MODES[m]) for m in ('tf32x3', 'bf16x3')}
            ref = ops.spconv_fwd_tc(feat, packed['tf32x3'], pair, sc, sh, res, True)
            scale = float(ref.abs().max().clamp(min=1.0))
            rows = {}
            configs = [('tf32x3', dict()), ('tf32x3 v2', dict(v=2)), ('tf32x3 v3', dict(v=3)),
                       ('tf32x3 occ=1', dict(occ=1)), ('tf32x3 occ=2', dict(occ=2)),
                       ('bf16x3', dict(p='bf16x3')), ('bf16x3 cps=2', dict(p='bf16x3', cps=2)),
                       ('bf16x3 occ=1', dict(p='bf16x3', occ=1)), ('bf16x3 occ=2', dict(p='bf16x3', occ=2)),
                       ('bf16x3 occ=1 cps=2', dict(p='bf16x3', occ=1, cps=2))]
            if args.with_v3_16bit:
                configs += [('bf16x3 v3', dict(p='bf16x3', v16=3))]
            for (name, c), srt in itertools.product(configs, (False, True) if sortable else (False,)):
                tcw = packed[c.get('p', 'tf32x3')]
                try:
                    ops.set_tc_variant(c.get('v', 0))
                    ops.set_tc16_variant(c.get('v16', 2))
                    tune(occ=c.get('occ', 0), cps=c.get('cps', 0))
                    if srt:
                        fn = lambda: ops.spconv_fwd_tc(feat, tcw, pair_s, sc, sh, res, True, row_perm=perm)  # noqa: E731
                    else:
                        fn = lambda: ops.spconv_fwd_tc(feat, tcw, pair, sc, sh, res, True)  # noqa: E731
                    out = fn()
                    err = float((out - ref).abs().max()) / scale
                    us = time_call(fn)
                    rows[name + (' +sort' if srt else '')] = dict(us=round(us, 2), err=err, ok=bool(err < 2e-5))
                except RuntimeError as e:   # a configuration the kernel refuses (does not fit, ...)
                    rows[name + (' +sort' if srt else '')] = dict(us=None, err=None, ok=False, error=str(e)[:120])
                finally:
                    ops.set_tc_variant(0)
                    ops.set_tc16_variant(2)
                    tune()
            base = rows['tf32x3']['us']
            good = {k: v for k, v in rows.items() if v['ok'] and v['us']}
            best = min(good, key=lambda k: good[k]['us'])
            best32 = min((k for k in good if k.startswith('tf32x3')), key=lambda k: good[k]['us'])
            lname = '%d->%d k%d n=%d%s' % (r['cin'], r['cout'], r['kvol'], r['n_out'], '+res' if r['residual'] else '')
            print('%-28s default %7.1f us | best %-24s %7.1f us (x%.2f) | best 3xTF32 %-18s %7.1f us (x%.2f)' % (
                lname, base, best, good[best]['us'], base / good[best]['us'], best32, good[best32]['us'],
                base / good[best32]['us']))
            report.append(dict(layer=lname, sweeps=sweeps, configs=rows, best=best, best_tf32x3=best32))
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(report, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
