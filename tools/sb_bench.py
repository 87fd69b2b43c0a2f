"""GPU tool: per-layer timing of the split-operand kernel (csrc/spconv_sb.cu, precision 'bf16x3c') against the
3xTF32 and bf16x3 kernels on the real rulebooks of the bench scene, with a correctness guard.

    python tools/sb_bench.py [--sweeps 1] [--iters 30] [--lc] [--json out.json]

For every distinct conv shape of the LiDAR SparseEncoder (and, with --lc, of the GMA fusion encoder) it prints the
CUDA-event time of: tf32x3 (default kernel), bf16x3 (spconv_tc16.cu), bf16x3c with the activations already split
(steady state inside a chain: the producing layer wrote the image) for several stage / occupancy settings, and the
stand-alone split kernel.  L2 is warm with the layer's own working set, as inside the network.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sweeps', type=int, default=1)
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--lc', action='store_true')
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    import torch
    import bench
    from msmdfusion_b200 import ops, synthetic
    from msmdfusion_b200 import fusion_encoder as fe
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

    se.SparseEncoder.use_executor = False
    fe.SparseMultiModalEncoderPaint.use_executor = False
    ops.PROFILE = []
    if args.lc:
        cfg, det, pts_np, meta, fpn = bench.build_lc_pipeline(dev, 0, 'S' if args.sweeps == 1 else 'L')
        with torch.no_grad():
            det.extract_voxel_space([torch.from_numpy(pts_np).to(dev)], fpn, [meta])
    else:
        cfg, layer, enc = bench.build_pipeline(dev)
        pts = torch.from_numpy(synthetic.lidar_scene(0, args.sweeps)).to(dev)
        with torch.no_grad():
            mean, coors, _ = layer.forward_mean(pts, 5, batch_idx=0)
            enc(mean, coors, 1)
    torch.cuda.synchronize()
    recs = [r for r in ops.PROFILE if r['op'] == 'spconv_fwd']
    ops.PROFILE = None
    seen, report = set(), []
    tot = dict(tf32x3=0.0, bf16x3=0.0, sb=0.0)
    tot['sb_tile'] = 0.0
    tot_ms = {}
    print('%-30s %9s %9s | %9s | %9s %9s %9s %9s | %7s %9s %9s' % ('layer', 'tf32x3', 'bf16x3', 'sb tile', 'sbp auto',
                                                                    'sbp occ1', 'sbp occ2', 'sbp st=2', 'split', 'err',
                                                                    'vs tile'))
    for r in recs:
        key = (r['cin'], r['cout'], r['kvol'], r['n_out'], r['residual'])
        mult = sum(1 for q in recs if (q['cin'], q['cout'], q['kvol'], q['n_out'], q['residual']) == key)
        if key in seen:
            continue
        seen.add(key)
        pair = r['pair']
        feat = torch.randn(r['n_in'], r['cin'], device=dev)
        w = torch.randn(r['cout'], r['kvol'], 1, 1, r['cin'], device=dev) * 0.05
        res = torch.randn(r['n_out'], r['cout'], device=dev) if r['residual'] else None
        sc, sh = torch.rand(r['cout'], device=dev) + 0.5, torch.randn(r['cout'], device=dev)
        packed = {m: ops.pack_weight_tc(w, ops.TC_MODES[m]) for m in ('tf32x3', 'bf16x3', 'bf16x3c')}
        ref = ops.spconv_fwd_tc(feat, packed['tf32x3'], pair, sc, sh, res, True)
        scale = float(ref.abs().max().clamp(min=1.0))
        row = dict(layer='%d->%d k%d n=%d%s' % (r['cin'], r['cout'], r['kvol'], r['n_out'], '+res' if r['residual'] else ''),
                   count=mult)
        row['tf32x3'] = time_call(lambda: ops.spconv_fwd_tc(feat, packed['tf32x3'], pair, sc, sh, res, True))
        row['bf16x3'] = time_call(lambda: ops.spconv_fwd_tc(feat, packed['bf16x3'], pair, sc, sh, res, True))
        xs = ops.split_bf16(feat)          # cached on `feat`: the timed calls below gather the existing image
        ops.check(L.msmd_spconv_sb_set_variant(1), 'msmd_spconv_sb_set_variant')   # one tile per CTA (r02c-k kernel)
        out_tile = ops.spconv_fwd_sb(feat, packed['bf16x3c'], pair, sc, sh, res, True)
        row['sb_tile'] = time_call(lambda: ops.spconv_fwd_sb(feat, packed['bf16x3c'], pair, sc, sh, res, True))
        ops.check(L.msmd_spconv_sb_set_variant(0), 'msmd_spconv_sb_set_variant')   # persistent, work-balanced
        out = ops.spconv_fwd_sb(feat, packed['bf16x3c'], pair, sc, sh, res, True)
        row['err'] = float((out - ref).abs().max()) / scale
        row['vs_tile'] = float((out - out_tile).abs().max()) / scale
        again = ops.spconv_fwd_sb(feat, packed['bf16x3c'], pair, sc, sh, res, True)
        row['deterministic'] = bool(torch.equal(out, again))
        img = out._msmd_split[1]
        chk = ops.split_bf16(out.clone())
        row['split_image_ok'] = bool(torch.equal(img, chk))
        for name, kw in (('sb', dict()), ('sb_occ1', dict(occ=1)), ('sb_occ2', dict(occ=2)), ('sb_st2', dict(stages=2))):
            tune(**kw)
            try:
                row[name] = time_call(lambda: ops.spconv_fwd_sb(feat, packed['bf16x3c'], pair, sc, sh, res, True))
            except RuntimeError as e:
                row[name] = float('nan')
                row[name + '_error'] = str(e)[:100]
            tune()
        row['sb_msort'] = float('nan')
        row['msort_err'] = float('nan')
        if r['kvol'] == 27 and r['n_in'] == r['n_out']:   # SubM 3x3x3: mask-sorted tiles (row permutation in the epilogue)
            row_perm, pair_sorted = ops.rulebook_mask_sort(pair)
            o2 = ops.spconv_fwd_sb(feat, packed['bf16x3c'], pair_sorted, sc, sh, res, True, row_perm=row_perm)
            row['msort_err'] = float((o2 - ref).abs().max()) / scale
            row['sb_msort'] = time_call(lambda: ops.spconv_fwd_sb(feat, packed['bf16x3c'], pair_sorted, sc, sh, res, True,
                                                                  row_perm=row_perm))
            row['msort_build'] = time_call(lambda: ops.rulebook_mask_sort(pair))
        pc = pair.clone()
        row['tile_masks_build'] = time_call(lambda: (pc.__dict__.pop('_msmd_tile_mask', None), ops.rulebook_tile_masks(pc)))
        f2 = feat.clone()
        row['split'] = time_call(lambda: (f2.__dict__.pop('_msmd_split', None), ops.split_bf16(f2)))
        for k in tot:
            tot[k] += mult * row[k]
        tot_ms['sb_msort'] = tot_ms.get('sb_msort', 0.0) + mult * (row['sb_msort'] if row['sb_msort'] == row['sb_msort'] else row['sb'])
        print('%-30s mask-sorted %8.1fu (err %.2e; sort %.1fu per rulebook), tile masks %.1fu per rulebook' % (
            '', row['sb_msort'], row['msort_err'], row.get('msort_build', float('nan')), row['tile_masks_build']))
        print('%-30s %8.1fu %8.1fu | %8.1fu | %8.1fu %8.1fu %8.1fu %8.1fu | %6.1fu %9.2e %9.2e %s %s x%d' % (
            row['layer'], row['tf32x3'], row['bf16x3'], row['sb_tile'], row['sb'], row['sb_occ1'], row['sb_occ2'],
            row['sb_st2'], row['split'], row['err'], row['vs_tile'],
            'img-ok' if row['split_image_ok'] else 'IMG-MISMATCH', 'det' if row['deterministic'] else 'NONDET', mult))
        report.append(row)
    print('sum over the %d conv launches of one scene: tf32x3 %.1f us, bf16x3 %.1f us, bf16x3c tile-per-CTA %.1f us, '
          'bf16x3c persistent %.1f us, persistent + mask-sorted SubM layers %.1f us' % (
              len(recs), tot['tf32x3'], tot['bf16x3'], tot['sb_tile'], tot['sb'], tot_ms.get('sb_msort', float('nan'))))
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(dict(layers=report, totals_us=tot, launches=len(recs)), open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
