"""Debug tool (GPU): per-role timeline of the warp-specialised tensor-core conv kernels, per layer of a
synthetic scene -- answers WHICH chain sets the per-chunk period (gather latency, stage hand-off, the
weight bulk copy, MMA issue), which ncu's per-kernel stall sums cannot separate for a producer /
consumer pipeline.

    python tools/tc_trace.py [--sweeps 1] [--precision tf32x3|bf16x3|bf16] [--mask-sort] [--json out.json]

Uses the -DMSMD_TC_TRACE build (csrc/tc_trace.cuh; built on first use into _C/libmsmd_b200_trace.so and
selected through MSMD_LIB).  Numbers are SM cycles of the traced CTAs (16 per launch, evenly spaced over
the grid) converted with the SM clock; tracing adds a few global stores per chunk, so read ratios, not
absolutes.  Per layer it prints

    period      mean time between consecutive MMA batches of a CTA (the per-chunk cost)
    mma_wait    share of the period the MMA issuer spends waiting for a full stage  (starved: producers
                or the weight copy are the limit)  -- for variant 3, split into weights / A operand
    prod_wait   share of the period a gather warp waits for a free stage (back-pressure: MMA is the limit)
    prod_fill   share spent between 'stage free' and 'arrived' (waiting for its own loads + convert + store)
    b_lead      how long before the MMA needs it the weight copy of a chunk was issued
    setup / epilogue / total   CTA phases in microseconds
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HEAD, ROLES, ITS, PHASES, CTAS = 16, 7, 128, 4, 16


def summarize(rec, mhz):
    """rec: (CTAS, words) uint64 numpy array of one launch."""
    import numpy as np
    out = []
    for r in rec:
        if r[0] == 0 or r[5] == 0:
            continue
        n_act = int(r[7])
        ev = r[HEAD:].reshape(ROLES, ITS, PHASES).astype(np.float64)
        # pipeline steps actually traced: one per chunk, or per group of chunks (MSMD_TC_TUNE cps=2), at most ITS
        n = int(np.count_nonzero(ev[3, :, 2]))
        cyc = 1.0 / mhz  # microseconds per cycle
        d = dict(cta=int(r[8]), sm=int(r[6]), n_act=n_act, steps=n, setup_us=(r[1] - r[0]) * cyc,
                 total_us=(r[5] - r[0]) * cyc, wall_us=(int(r[10]) - int(r[9])) * 1e-3)
        if r[2] and r[4]:
            d['epilogue_us'] = (r[4] - r[2]) * cyc
            d['main_us'] = (r[2] - r[1]) * cyc
        if n >= 3:
            mma = ev[3, :n]
            period = (mma[-1, 2] - mma[0, 2]) / (n - 1)
            d['period_us'] = period * cyc
            d['mma_wait'] = float(np.mean(mma[1:, 1] - mma[1:, 0]) / period)
            if mma[1:, 3].all():   # variant 3: phase 3 = weights in
                d['mma_wait_weights'] = float(np.mean(mma[1:, 3] - mma[1:, 0]) / period)
            d['mma_issue'] = float(np.mean(mma[:, 2] - mma[:, 1]) / period)
            for role, name in ((0, 'prod0'), (1, 'prod7')):
                p = ev[role, :n]
                if p[:, 2].all():
                    d[name + '_wait'] = float(np.mean(p[1:, 1] - p[1:, 0]) / period)
                    d[name + '_fill'] = float(np.mean(p[1:, 2] - p[1:, 1]) / period)
            b = ev[2, :n]
            if b[:, 1].all():
                d['b_wait'] = float(np.mean(b[1:, 1] - b[1:, 0]) / period)
                d['b_lead_us'] = float(np.mean(mma[1:, 0] - b[1:, 1]) * cyc)  # copy issued -> MMA starts waiting
            d['mma_gaps_us'] = [round(float(x) * cyc, 3) for x in np.diff(mma[:, 2])]
        # persistent kernel: one record per segment from epilogue warp 0 (role 4) and the segment loader (role 5)
        nseg = int(np.count_nonzero(ev[4, :, 3]))
        if nseg:
            e, sl = ev[4, :nseg], ev[5, :nseg]
            t0 = float(r[0])
            d['segments'] = [dict(acc_wait_from=round((e[i, 0] - t0) * cyc, 2), acc_full=round((e[i, 1] - t0) * cyc, 2),
                                  flags_seen=round((e[i, 2] - t0) * cyc, 2), rows_written=round((e[i, 3] - t0) * cyc, 2),
                                  loader_wait_from=round((sl[i, 0] - t0) * cyc, 2), loader_free=round((sl[i, 1] - t0) * cyc, 2),
                                  pairs_landed=round((sl[i, 2] - t0) * cyc, 2), published=round((sl[i, 3] - t0) * cyc, 2))
                             for i in range(nseg)]
            st6 = ev[6]
            d['epi_steps'] = [[i // 8, i % 8] + [round((st6[i, ph] - t0) * cyc, 2) if st6[i, ph] else None for ph in (0, 1, 3)]
                              for i in range(ITS) if st6[i, 0]]
            d['first_mma_us'] = round((ev[3, 0, 1] - t0) * cyc, 2) if n else None
            d['last_mma_us'] = round((ev[3, n - 1, 2] - t0) * cyc, 2) if n else None
        out.append(d)
    return out


def mean_of(rows, key):
    v = [r[key] for r in rows if key in r]
    return sum(v) / len(v) if v else float('nan')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sweeps', type=int, default=1)
    ap.add_argument('--precision', default='tf32x3', choices=['tf32x3', 'bf16x3', 'bf16x3c', 'bf16'])
    ap.add_argument('--mask-sort', action='store_true')
    ap.add_argument('--json', default=None)
    ap.add_argument('--sb-variant', type=int, default=0, help='bf16x3c: 1 = tile per CTA, 2 / 0 = persistent')
    ap.add_argument('--only', default=None, help='comma list of layer names to trace, e.g. "128->128 k27"')
    ap.add_argument('--dump-cta', action='store_true', help='print the per-segment timeline of the first traced CTAs')
    args = ap.parse_args()

    # the library path is read when msmdfusion_b200._cabi is imported, i.e. by ANY import of the package: load the
    # build script by path, set MSMD_LIB, and only then import the package
    import importlib.util
    spec = importlib.util.spec_from_file_location('_msmd_build', os.path.join(ROOT, 'msmdfusion_b200', 'build.py'))
    _build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(_build)
    os.environ['MSMD_LIB'] = _build.build_trace()
    import numpy as np
    import torch
    import bench
    from msmdfusion_b200 import _cabi, ops, synthetic
    from msmdfusion_b200 import sparse_encoder as se
    L = _cabi.lib()
    for name in ('msmd_tc_trace_set', 'msmd_tc16_trace_set', 'msmd_sb_trace_set'):
        getattr(L, name).restype = ctypes.c_int
        getattr(L, name).argtypes = [ctypes.c_void_p]
    L.msmd_spconv_sb_set_variant.argtypes = [ctypes.c_int]
    assert L.msmd_spconv_sb_set_variant(args.sb_variant) == 0
    L.msmd_tc_trace_record_words.restype = ctypes.c_int
    words = L.msmd_tc_trace_record_words()
    assert words == HEAD + ROLES * ITS * PHASES
    dev = torch.device('cuda:0')
    mhz = float(os.environ.get('MSMD_SM_MHZ', 1965.0))   # B200 SM clock under load in every bench run of round 1
    buf = torch.zeros(CTAS * words, dtype=torch.int64, device=dev)

    cfg, layer, enc = bench.build_pipeline(dev)
    se.SparseEncoder.use_executor = False
    pts = torch.from_numpy(synthetic.lidar_scene(0, args.sweeps)).to(dev)
    ops.PROFILE = []
    with torch.no_grad():
        mean, coors, _ = layer.forward_mean(pts, 5, batch_idx=0)
        enc(mean, coors, 1)
    torch.cuda.synchronize()
    recs = [r for r in ops.PROFILE if r['op'] == 'spconv_fwd']
    ops.PROFILE = None
    mode = ops.TC_MODES[args.precision]
    seen, report = set(), []
    print('%-22s %7s %6s | %7s %8s %8s %9s %9s %7s %8s | %6s %6s %6s %7s' % (
        'layer', 'n_out', 'chunks', 'period', 'mma_wait', '(weights)', 'prod_wait', 'prod_fill', 'b_wait', 'b_lead',
        'setup', 'main', 'epi', 'total'))
    for r in recs:
        key = (r['cin'], r['cout'], r['kvol'], r['n_out'], r['residual'])
        if key in seen:
            continue
        seen.add(key)
        if args.only and ('%d->%d k%d' % (r['cin'], r['cout'], r['kvol'])) not in args.only.split(','):
            continue
        pair = r['pair']
        feat = torch.randn(r['n_in'], r['cin'], device=dev)
        w = torch.randn(r['cout'], r['kvol'], 1, 1, r['cin'], device=dev) * 0.05
        tcw = ops.pack_weight_tc(w, mode)
        res = torch.randn(r['n_out'], r['cout'], device=dev) if r['residual'] else None
        sc, sh = torch.ones(r['cout'], device=dev), torch.zeros(r['cout'], device=dev)
        row_perm = None
        if args.mask_sort and r['kvol'] == 27 and r['n_in'] == r['n_out']:
            row_perm, pair = ops.rulebook_mask_sort(pair)
        for _ in range(3):   # warm-up untraced
            ops.spconv_fwd_tc(feat, tcw, pair, sc, sh, res, True, row_perm=row_perm)
        torch.cuda.synchronize()
        buf.zero_()
        L.msmd_tc_trace_set(buf.data_ptr())
        L.msmd_tc16_trace_set(buf.data_ptr())
        L.msmd_sb_trace_set(buf.data_ptr())
        ops.spconv_fwd_tc(feat, tcw, pair, sc, sh, res, True, row_perm=row_perm)
        torch.cuda.synchronize()
        L.msmd_tc_trace_set(None)
        L.msmd_tc16_trace_set(None)
        L.msmd_sb_trace_set(None)
        rows = summarize(buf.cpu().numpy().astype(np.uint64).reshape(CTAS, words), mhz)
        name = '%d->%d k%d%s' % (r['cin'], r['cout'], r['kvol'], '+res' if r['residual'] else '')
        print('%-22s %7d %6.0f | %6.2fus %7.0f%% %8.0f%% %8.0f%% %8.0f%% %6.0f%% %6.2fus | %6.1f %6.1f %6.1f %6.1fus' % (
            name, r['n_out'], mean_of(rows, 'n_act'), mean_of(rows, 'period_us'), 100 * mean_of(rows, 'mma_wait'),
            100 * mean_of(rows, 'mma_wait_weights'), 100 * mean_of(rows, 'prod0_wait'), 100 * mean_of(rows, 'prod0_fill'),
            100 * mean_of(rows, 'b_wait'), mean_of(rows, 'b_lead_us'), mean_of(rows, 'setup_us'),
            mean_of(rows, 'main_us'), mean_of(rows, 'epilogue_us'), mean_of(rows, 'total_us')))
        if args.dump_cta:
            for row in rows[:3] + rows[-1:]:
                print('   cta %d sm %d: chunks %d, setup %.1f us, first mma %s, last mma %s, total %.1f us' % (
                    row['cta'], row['sm'], row['n_act'], row['setup_us'], row.get('first_mma_us'), row.get('last_mma_us'),
                    row['total_us']))
                for i, sg in enumerate(row.get('segments', [])):
                    print('      seg %d: loader wait %.1f free %.1f landed %.1f published %.1f | acc wait from %.1f full %.1f '
                          'flags %.1f written %.1f' % (i, sg['loader_wait_from'], sg['loader_free'], sg['pairs_landed'],
                                                       sg['published'], sg['acc_wait_from'], sg['acc_full'],
                                                       sg['flags_seen'], sg['rows_written']))
                for e in row.get('epi_steps', []):
                    print('         epilogue seg %d step %d: tmem read %s staged %s stored %s' % tuple(e))
                gaps = row.get('mma_gaps_us', [])
                print('      mma batch gaps (us):', ' '.join('%.2f' % g for g in gaps[:64]))
        report.append(dict(layer=name, n_out=r['n_out'], ctas=rows))
    if args.json:
        json.dump(dict(precision=args.precision, mask_sort=args.mask_sort, sweeps=args.sweeps, sm_mhz=mhz,
                       layers=report), open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
