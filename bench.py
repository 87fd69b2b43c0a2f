#!/usr/bin/env python
"""bench.py -- scenes/s of the MSMDFusion voxel-space hot path on B200 (contract: see DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--profile S|L]

A "step" is one pass of the hot path over one synthetic nuScenes-shaped scene:
  hard_voxelize (+fused HardSimpleVFE mean) -> SparseEncoder (21 sparse convs, BN, ReLU)
  -> dense()  ==  the LiDAR branch of transfusion_nusc_voxel_L / MSMDFusion_nusc_voxel_LC
(BASELINE.json configs[1]).  One scene per GPU, no data-path collective (weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 and libgomp reads it once, when the first OpenMP runtime is
# loaded (import torch): the CPU arms (--impl reference / cpu_baseline) are specified to use all
# the host threads they can, so the variable is fixed BEFORE numpy / torch are imported
if os.environ.get('OMP_NUM_THREADS', '') in ('', '1'):
    os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'nuScenes-shape scenes/sec forward (voxel hot path: hard_voxelize+VFE+SparseEncoder+dense; --workload LC adds lift+split+GMA encoder)'
UNIT = 'scenes/s'
SETTLE_MS = 200.0  # untimed pipeline run before the warm-up steps (see run_ours)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    # defaults (flags absent): 100 / 30 for this implementation (a fresh box needs ~25 steps to reach steady clocks,
    # profiles/r01f_bench_S.json); 3 / 1 for --impl reference, whose step is one whole scene on the host cores (5-7 s)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--profile', default='S', choices=['S', 'L'],
                    help='S: single 32-beam sweep (~30 k points); L: 10 sweeps (~285 k points)')
    ap.add_argument('--workload', default='LC', choices=['L', 'LC', 'train'],
                    help='LC: MSMDFusion_nusc_voxel_LC voxel-space fusion path (BASELINE configs[2]/[3], the workload '
                         'north_star quotes the metric on; the default); '
                         'L: transfusion_nusc_voxel_L LiDAR hot path (configs[1]); '
                         'train: LC train step (configs[4]): forward in train mode, backward through the GMA '
                         'encoder, one NCCL gradient all-reduce, clip, AdamW')
    ap.add_argument('--precision', default=None, choices=['tf32x3', 'bf16x3', 'bf16x3c', 'bf16'],
                    help='operand precision of the tensor-core sparse convolutions (default: MSMD_CONV_PRECISION or '
                         'bf16x3c = the library default: bf16 hi/lo split with the split cached by the producing layer, ~5e-6 per '
                         'layer vs fp32).  tf32x3 / bf16x3: the round-1 kernels with the same error class.  bf16: operands '
                         'rounded to bf16 (the train-step arithmetic of BASELINE configs[4]; not a parity mode)')
    ap.add_argument('--scenes-per-step', type=int, default=1,
                    help='workload L only: scenes batched into one step on each GPU (default 1 = the workload '
                         'BASELINE configs[1] names; the reference config trains with samples_per_gpu=2).  Reported in '
                         'config.scenes_per_gpu_per_step; value counts scenes, not steps')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cuda-baseline', action='store_true',
                    help='skip the `cuda_baseline` leg (the reference\'s own CUDA kernels on the same scene)')
    ap.add_argument('--breakdown', default=None, help='write a per-op CUDA-event breakdown (json) here')
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.impl == 'reference' else 100
    if args.warmup is None:
        args.warmup = 1 if args.impl == 'reference' else 30
    return args


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('bf16_tflops', 1590.0)), 'measured'
    return 6650.0, 1590.0, 'fallback'


ARITHMETIC = {
    'tf32x3': 'fp32 in/out; contraction = 3xTF32 tensor-core split with fp32 accumulate (<=1e-4 vs fp32 oracle)',
    'bf16x3': 'fp32 in/out; contraction = bf16 hi/lo split (3 kind::f16 MMAs per product) with fp32 accumulate '
              '(~5e-6 per layer vs fp32; opt-in)',
    'bf16x3c': 'fp32 in/out; contraction = bf16 hi/lo split (3 kind::f16 MMAs per product, fp32 accumulate, ~5e-6 per '
               'layer vs fp32) with the activations\' split cached by the producing layer (csrc/spconv_sb.cu)',
    'bf16': 'fp32 storage; contraction operands rounded to bf16 (1 kind::f16 MMA per product), fp32 accumulate -- '
            'NOT a parity mode (2e-3 per layer): the train-step arithmetic of BASELINE configs[4]',
}


def workload_name(profile, workload='L'):
    scene = '30k-pt single-sweep' if profile == 'S' else '285k-pt 10-sweep'
    if workload == 'train':
        return ('MSMDFusion_nusc_voxel_LC TRAIN step on the voxel-space path (frozen LiDAR encoder, lift x4 scales, '
                'GMA encoder forward + backward, quadratic loss on the BEV tensor in place of TransFusionHead.loss, '
                'flat-buffer gradient all-reduce, clip 10, AdamW), synthetic %s scene + 60k virtual points, one '
                'scene per GPU' % scene)
    if workload == 'LC':
        return ('MSMDFusion_nusc_voxel_LC voxel-space fusion path (LiDAR encoder + 6-camera virtual-point '
                'lift x4 scales + modality split + GMA encoder + dense), synthetic %s scene + 60k virtual '
                'points' % scene)
    return 'transfusion_nusc_voxel_L LiDAR hot path, synthetic %s scene' % scene


def hotpath_cfg():
    from msmdfusion_b200 import Config
    return Config.fromfile(os.path.join(ROOT, 'configs', 'msmd_lc_hotpath.py')).hotpath


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# ------------------------------------------------------------------------------------------
class Clocks:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        if os.environ.get('MSMD_BENCH_NOSMI'):  # diagnostic switch
            return
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Index of the next sample: brackets the timed region so that only samples taken DURING
        it are summarised."""
        return len(self.lines)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        window = self.lines[first:last] if len(self.lines[first:last]) >= 2 else self.lines
        for ln in window:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def build_pipeline(device):
    import msmdfusion_b200 as m
    from msmdfusion_b200 import registry
    cfg = hotpath_cfg()
    torch.manual_seed(0)
    enc = registry.build_middle_encoder(cfg.pts_middle_encoder).to(device).eval()
    layer = m.Voxelization(**cfg.pts_voxel_layer).eval()
    return cfg, layer, enc


def build_lc_pipeline(device, rank, profile):
    """MSMDFusionDetector voxel-space path + its synthetic inputs (device-resident FPN features,
    host-side virtual points exactly as the reference's img_metas carry them)."""
    import msmdfusion_b200 as m
    from msmdfusion_b200 import synthetic
    cfg = hotpath_cfg()
    torch.manual_seed(0)
    det = m.MSMDFusionDetector(**{k: cfg[k] for k in (
        'pts_voxel_layer', 'pts_voxel_encoder', 'pts_middle_encoder', 'multimodal_middle_encoder',
        'spatial_shapes', 'downscale_factors', 'fps_num_list', 'radius_list', 'max_cluster_samples_list',
        'dist_thresh_list')}).to(device).eval()
    with torch.no_grad():
        det.score_net[0].weight.mul_(0.2)
        det.score_net[0].bias.fill_(0.05)
    pts_np = synthetic.lidar_scene(seed=rank, sweeps=1 if profile == 'S' else 10)
    meta = synthetic.camera_scene(rank, pts_np)
    fpn = [torch.from_numpy(f).to(device) for f in synthetic.fpn_features(rank, batch=1)]
    return cfg, det, pts_np, meta, fpn


def conv_layer_bytes_flops(records):
    """Algorithmic bytes / flops of each sparse-conv launch:
    bytes = 4*(N_in*Cin + N_out*Cout + K*Cin*Cout) + 4*K*N_out ; flops = 2*P*Cin*Cout."""
    tot_b = tot_f = 0.0
    for r in records:
        tot_b += 4.0 * (r['n_in'] * r['cin'] + r['n_out'] * r['cout'] + r['kvol'] * r['cin'] * r['cout']) \
            + 4.0 * r['kvol'] * r['n_out'] + (4.0 * r['n_out'] * r['cout'] if r['residual'] else 0.0)
        tot_f += 2.0 * r['pairs'] * r['cin'] * r['cout']
    return tot_b, tot_f


def run_ours(args, rank, world, device):
    import torch.distributed as dist
    from msmdfusion_b200 import _cabi, ops, synthetic
    from msmdfusion_b200 import spconv as _spconv
    if args.precision:
        _spconv.CONV_PRECISION = args.precision
    args.precision = _spconv.CONV_PRECISION
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    trainer = None
    if args.workload == 'train':
        from msmdfusion_b200 import train as _train
        cfg, det, pts_np, meta, fpn = build_lc_pipeline(device, rank, args.profile)
        pts_host = torch.from_numpy(pts_np).pin_memory()
        pts_dev = pts_host.to(device)
        h2d_extra = [0]
        metas_resident = [meta]
        target = torch.randn((1, 640, 180, 180), generator=torch.Generator().manual_seed(1 + rank)).to(device)
        trainer = _train.VoxelSpaceTrainStep(det, lambda bev: ((bev - target) ** 2).mean())

        def step(points, fresh_upload=False):
            metas = [dict(meta)] if fresh_upload else metas_resident
            loss = trainer([points], fpn, metas)
            if fresh_upload:
                h2d_extra[0] = det.packed_foreground(metas, device).h2d_bytes
            return loss, trainer.last_stage_outs
    elif args.workload == 'LC':
        cfg, det, pts_np, meta, fpn = build_lc_pipeline(device, rank, args.profile)
        pts_host = torch.from_numpy(pts_np).pin_memory()
        pts_dev = pts_host.to(device)
        h2d_extra = [0]

        metas_resident = [meta]  # same list object every step -> the packed upload stays on the device

        def step(points, fresh_upload=False):
            metas = [dict(meta)] if fresh_upload else metas_resident  # new dict -> pack + upload again
            with torch.no_grad():
                bev, stage_outs = det.extract_voxel_space([points], fpn, metas)
            if fresh_upload:
                h2d_extra[0] = det.packed_foreground(metas, device).h2d_bytes
            return bev, stage_outs
    else:
        cfg, layer, enc = build_pipeline(device)
        B = max(1, args.scenes_per_step)
        sweeps = 1 if args.profile == 'S' else 10
        scenes = [synthetic.lidar_scene(seed=rank * B + b, sweeps=sweeps) for b in range(B)]
        pts_np = np.concatenate(scenes) if B > 1 else scenes[0]
        bounds = np.cumsum([0] + [len(sc) for sc in scenes])
        pts_host = torch.from_numpy(pts_np).pin_memory()
        pts_dev = pts_host.to(device)
        h2d_extra = [0]

        def step(points, fresh_upload=False):
            with torch.no_grad():
                if B == 1:
                    mean, coors, _ = layer.forward_mean(points, 5, batch_idx=0)
                else:   # per-sample voxelisation, batch index in front (mvx_two_stage.py voxelize)
                    parts = [layer.forward_mean(points[bounds[b]:bounds[b + 1]], 5, batch_idx=b) for b in range(B)]
                    mean = torch.cat([q[0] for q in parts])
                    coors = torch.cat([q[1] for q in parts])
                spatial, feats = enc(mean, coors, B)
            return spatial, feats

    def sync_all():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    clocks = Clocks(torch.cuda.current_device())
    clocks.start()  # nvidia-smi needs ~100 ms to produce its first sample: start it before warm-up
    # settle: a freshly acquired box runs its first ~50 ms of work below steady speed with the SM clock
    # already reading max (profiles/r01f_bench_S.json), so the pipeline is driven for SETTLE_MS of wall
    # time before the W warm-up steps the caller asked for; untimed, reported in config.settle_ms
    t_settle = None
    settle_steps = 0
    def keep_settling():
        more = t_settle is None or (time.perf_counter() - t_settle) * 1e3 < SETTLE_MS
        if world > 1 and trainer is not None:
            # the train step holds a collective (the gradient all-reduce): every rank must run the SAME number of
            # steps, so the wall-clock decision is taken jointly (r02l: rank-local counts dead-locked the N=2 run)
            flag = torch.tensor([int(more)], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            more = bool(int(flag.item()))
        return more

    while keep_settling():
        flush.zero_()
        spatial, feats = step(pts_dev)
        torch.cuda.synchronize(device)
        settle_steps += 1
        if t_settle is None:  # the very first step carries lazy initialisation: the clock starts after it
            t_settle = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        # results are kept across iterations exactly as in the timed loops, so the caching
        # allocator reaches its steady state (two live result arenas) during warm-up
        flush.zero_()
        spatial, feats = step(pts_dev)
    sync_all()
    # serving-loop hygiene: move everything allocated so far out of the cyclic GC's reach so a
    # generation-2 collection cannot stall a timed step (the model graph is static from here on)
    gc.collect()
    gc.freeze()
    if os.environ.get('MSMD_BENCH_NOGC'):  # diagnostic switch
        gc.disable()

    # ---- device-resident timing: K steps, L2 flushed between steps, CUDA events per step ----
    mark0 = clocks.mark()
    ms0 = torch.cuda.memory_stats(device)
    launches0 = _cabi.lib().msmd_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    if trainer is not None:
        trainer.exchange_events = []
    for s, e in ev:
        flush.zero_()
        s.record()
        spatial, feats = step(pts_dev)
        e.record()
    sync_all()
    exchange = None
    if trainer is not None:
        ex_ms, ex_n = trainer.exchange_ms()
        trainer.exchange_events = None
        exchange = {'all_reduce_ms_per_step': round(ex_ms / max(ex_n, 1), 4), 'steps': ex_n,
                    'gradient_bytes': int(trainer.grads.flat.numel() * 4),
                    'note': 'one NCCL all-reduce of the flat fp32 gradient buffer, issued on the compute stream after '
                            'the backward pass: its duration is its exposed time (no-op at 1 GPU)'}
    launches = _cabi.lib().msmd_launch_count() - launches0
    ms1 = torch.cuda.memory_stats(device)
    alloc_diag = {k: int(ms1.get(k, 0) - ms0.get(k, 0)) for k in
                  ('num_device_alloc', 'num_device_free', 'num_alloc_retries', 'num_sync_all_streams')}
    alloc_diag['reserved_gb'] = round(ms1.get('reserved_bytes.all.current', 0) / 1e9, 2)
    step_raw = [s.elapsed_time(e) for s, e in ev]
    step_ms = sorted(step_raw)
    dev_ms = sum(step_ms)

    # ---- end to end: pinned host points -> H2D -> hot path -> D2H of a result checksum ----
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    checksum = 0.0
    for _ in range(3):  # warm the e2e path itself (first .sum() loads its reduction kernel lazily)
        spatial, feats = step(pts_host.to(device, non_blocking=True), fresh_upload=True)
        checksum = float(spatial.sum().item())
    sync_all()
    for s, e in ev2:
        flush.zero_()
        s.record()
        p = pts_host.to(device, non_blocking=True)
        spatial, feats = step(p, fresh_upload=True)
        checksum = float(spatial.sum().item())  # device -> host read of the step result
        e.record()
    sync_all()
    clk = clocks.stop(mark0, clocks.mark())
    e2e_raw = [s.elapsed_time(e) for s, e in ev2]
    e2e_ms = sum(e2e_raw)

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    # ---- per-kernel timing of the dominant kernel (sparse conv), live, CUDA events ----
    roof = None
    if rank != 0 and trainer is not None and world > 1:
        # the train step holds a collective: the 4 instrumented steps rank 0 runs below need their partners
        # (r02l / r02u: without them rank 0 waited in the gradient all-reduce for ranks that had already left)
        for _ in range(4):
            flush.zero_()
            step(pts_dev)
        torch.cuda.synchronize(device)
    if rank == 0:
        # per-kernel timing needs one C-ABI call per convolution: run the module-by-module path
        # (same kernels, same operands) instead of the one-call native executor for these 4 steps
        from msmdfusion_b200 import fusion_encoder as _fe
        from msmdfusion_b200 import sparse_encoder as _se
        _se.SparseEncoder.use_executor = False
        _fe.SparseMultiModalEncoderPaint.use_executor = False
        ops.PROFILE = []
        flush.zero_()
        step(pts_dev)                       # first instrumented step warms the event pool: discarded
        torch.cuda.synchronize(device)
        ops.PROFILE = []
        for _ in range(3):
            flush.zero_()
            step(pts_dev)
        torch.cuda.synchronize(device)
        recs = ops.PROFILE
        ops.PROFILE = None
        _se.SparseEncoder.use_executor = True
        _fe.SparseMultiModalEncoderPaint.use_executor = True
        # Kernel durations.  The events around a single C-ABI call of the module-by-module path also span the
        # host's gap to the next launch (~10-20 us of Python per call, more than a 6 us kernel).  So every conv call
        # that carries a `replay` closure (the identical launch: same operands, same stream) is re-issued REPLAY
        # times back to back between two events -- the launch queue stays full and elapsed / REPLAY is the kernel's
        # own duration (inputs L2-warm, as inside the chain where the previous layer just wrote them).
        REPLAY = 8
        replayed = 0
        for r in recs:
            r['ms'] = r['start'].elapsed_time(r['end'])
            rp = r.pop('replay', None)
            if rp is not None and r['op'] == 'spconv_fwd':
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                rp()
                e0.record()
                for _ in range(REPLAY):
                    rp()
                e1.record()
                e1.synchronize()
                r['ms_call'] = r['ms']
                r['ms'] = e0.elapsed_time(e1) / REPLAY
                replayed += 1
            if 'pair' in r:
                r['pairs'] = int((r.pop('pair') >= 0).sum().item())
        conv = [r for r in recs if r['op'] in ('spconv_fwd', 'spconv_bwd_data', 'spconv_bwd_weight')]
        for r in conv:
            r.setdefault('residual', False)
        if args.breakdown:
            by_op = {}
            for r in recs:
                by_op.setdefault(r['op'], [0, 0.0])
                by_op[r['op']][0] += 1
                by_op[r['op']][1] += r['ms']
            layers = [{k: v for k, v in r.items() if k not in ('start', 'end')} for r in recs[:len(recs) // 3]]
            json.dump({'per_op_calls_ms_over_3_steps': by_op, 'first_step_calls': layers},
                      open(args.breakdown, 'w'), indent=1)
        b, f = conv_layer_bytes_flops(conv)
        conv_ms = sum(r['ms'] for r in conv)
        hbm, bf16, src = peaks()
        paths = sorted({r.get('path', 'simt') for r in conv})
        tc = paths == ['tc'] or (args.workload == 'train' and 'tc' in paths)
        gbs = b / (conv_ms * 1e-3) / 1e9 if conv_ms else 0.0
        tfl = f / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0
        # tcgen05 kind::tf32 issues at half the bf16 rate; the 16-bit modes run kind::f16 at the bf16 rate
        tf32_peak = bf16 / 2.0 if args.precision == 'tf32x3' else bf16
        mma_per_product = 1 if args.precision == 'bf16' else 3
        hbm_time = b / (hbm * 1e9)
        tensor_time = (mma_per_product * f) / (tf32_peak * 1e12) if tc else 0.0
        # DRAM traffic cannot be measured inside an un-profiled run: it is taken from a committed
        # `ncu --set full` capture ONLY if one exists for exactly this workload / profile / precision
        # (profiles/traffic_<workload>_<profile>_<precision>.json, written by tools/ncu_traffic.py from
        # the capture named inside it); otherwise null
        traffic, traffic_src, ncu_tensor_pct = None, None, None
        tp = os.path.join(ROOT, 'profiles', 'traffic_%s_%s_%s.json' % (args.workload, args.profile, args.precision))
        if os.path.exists(tp):
            t = json.load(open(tp))
            traffic = float(t['dram_bytes_per_step'])
            ncu_tensor_pct = t.get('tensor_active_pct_time_weighted')
            traffic_src = t.get('source')
        common = {'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': src,
                  'kernel': ('sparse conv forward + data gradient (%s) + weight gradient '
                             '(spconv_wgrad_simt_kernel, FFMA)' % {
                                 'tf32x3': 'tcgen05 kind::tf32, 3xTF32', 'bf16x3c': 'tcgen05 kind::f16, split-bf16 operand '
                                 'cache', 'bf16x3': 'tcgen05 kind::f16, bf16 x3', 'bf16': 'tcgen05 kind::f16, bf16 operands'
                             }[args.precision] if args.workload == 'train' else
                             'spconv_fwd_tc_kernel (tcgen05 kind::tf32, 3xTF32)' if tc and args.precision == 'tf32x3' else
                             'spconv_fwd_sbp_kernel / spconv_fwd_sb_kernel (tcgen05 kind::f16 on cached split-bf16 operands, '
                             'cp.async gather; persistent work-balanced schedule for N >= 64 and >= 16 K chunks per tile, one '
                             'tile per CTA for the narrow layers)'
                             if tc and args.precision == 'bf16x3c' else
                             'spconv_fwd_tc16_kernel (tcgen05 kind::f16 on bf16 operands, %s)' % args.precision if tc else
                             'spconv_fwd_simt_kernel') + ' (%d launches/scene, summed)' % (len(conv) // 3),
                  'kernel_ms_per_step': round(conv_ms / 3, 4),
                  'kernel_timing': ('%d of %d conv launches timed by replaying the identical launch %d x back to back '
                                    'between two CUDA events (kernel duration without the host gap of the '
                                    'module-by-module path); the rest by events around the single call' %
                                    (replayed, len(conv), REPLAY)),
                  'algorithmic_bytes_per_step': b / 3, 'algorithmic_flops_per_step': f / 3,
                  'hbm': {'achieved_gbs': round(gbs, 2), 'peak_gbs': hbm, 'frac': round(gbs / hbm, 4)},
                  'tensor': {'achieved_tflops': round(tfl, 3), 'peak_tf32_tflops': round(tf32_peak, 1),
                             'frac': round(tfl / tf32_peak, 4), 'mma_per_product': mma_per_product if tc else 0,
                             'operand_precision': args.precision,
                             'ncu_tensor_pipe_active_pct_time_weighted': ncu_tensor_pct,
                             'note': 'algorithmic flops 2*P*Cin*Cout; the fp32-parity modes issue 3 MMAs per '
                                     'product (hi*hi + hi*lo + lo*hi), so the attainable ceiling is peak/3; `peak` is '
                                     'the measured dense bf16 rate (kind::f16 modes) or half of it (kind::tf32)'}}
        if tc and tensor_time > hbm_time:
            roof = dict(bound='tensor', achieved=round(tfl, 3), peak=round(tf32_peak, 1), unit='TFLOP/s',
                        frac=round(tfl / tf32_peak, 4), **common)
        else:
            roof = dict(bound='hbm', achieved=round(gbs, 2), peak=hbm, unit='GB/s', frac=round(gbs / hbm, 4),
                        **common)
    # ---- reference CUDA kernels on the same scene (rank 0, N=1): see oracle/cuda_baseline.py ----
    cuda_base = None
    if rank == 0 and world == 1 and args.workload in ('L', 'LC') and not args.no_cuda_baseline:
        try:
            from oracle import cuda_baseline as _cb
            if _cb.available():
                n_cb = max(3, min(args.steps, 10))
                cb_ms, _ = _cb.time_steps(step, pts_dev, device, n_cb, 3, flush)
                cb_ms = sorted(cb_ms)
                cuda_base = {'value': 1e3 / (sum(cb_ms) / len(cb_ms)), 'unit': UNIT, 'ms_per_step': sum(cb_ms) / len(cb_ms),
                             'step_ms': {'min': round(cb_ms[0], 3), 'median': round(cb_ms[len(cb_ms) // 2], 3),
                                         'max': round(cb_ms[-1], 3)},
                             'steps': n_cb, 'warmup': 3, 'kind': _cb.KIND,
                             'timing': 'same scene, device-resident inputs, L2 flushed between steps, CUDA events '
                                       '(compare with `value`, not with `e2e`)'}
            else:
                cuda_base = {'unavailable': 'oracle/_ref reference CUDA libraries are not built (they are compiled '
                                            'from /root/reference by __graft_entry__.build())'}
        except Exception as e:  # noqa: BLE001 - a reported baseline must not take the bench line down
            cuda_base = {'unavailable': 'reference CUDA baseline failed: %s' % (str(e)[-300:],)}
    n_vox = int(feats[0].indices.shape[0])
    return dict(cuda_baseline=cuda_base, dev_ms=dev_ms, e2e_ms=e2e_ms, launches=int(launches), exchange=exchange, clocks=clk, roofline=roof, settle_steps=settle_steps,
                e2e_step_ms=dict(min=round(min(e2e_raw), 4), median=round(sorted(e2e_raw)[len(e2e_raw) // 2], 4),
                                 max=round(max(e2e_raw), 4), all=[round(x, 2) for x in e2e_raw[:32]]),
                alloc=alloc_diag,
                step_ms=dict(min=round(step_ms[0], 4), median=round(step_ms[len(step_ms) // 2], 4),
                             max=round(step_ms[-1], 4), all=[round(x, 3) for x in step_raw[:32]]),
                points=int(pts_np.shape[0]), voxels=n_vox, checksum=checksum,
                virtual_points=(int(sum(len(x) for x in meta['foreground2D_info']['fg_pixels'])) if args.workload != 'L' else 0),
                h2d=int(pts_np.nbytes) + int(h2d_extra[0]), d2h=4)


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU hard_voxelize (oracle/_ref, compiled from the reference
# sources) + the oracle port of the un-vendored spconv-2.x arithmetic, all host threads.
# ------------------------------------------------------------------------------------------
def cpu_pass(profile, workload='L', seed=0):
    """One whole-scene pass of the workload on the host cores -> (callable, threads, kind, description,
    points).  L: voxelize + VFE + SparseEncoder.  LC / train: the whole voxel-space fusion forward
    (oracle.model.extract_voxel_space = the reference's extract_pts_feat up to the BEV tensor,
    MSMDFusion.py:421-445; pinned against the reference method run in place, tests/test_oracle.py)."""
    from msmdfusion_b200 import registry, synthetic
    from oracle import build as obuild
    from oracle import cpu
    from oracle import model as omodel
    cfg = hotpath_cfg()
    torch.manual_seed(0)
    pts = synthetic.lidar_scene(seed=seed, sweeps=1 if profile == 'S' else 10)
    use_ref = obuild.ref_so_path() is not None
    vox = cpu.hard_voxelize_ref if use_ref else cpu.hard_voxelize
    kind = 'reference' if use_ref else 'port'

    if workload == 'L':
        enc = registry.build_middle_encoder(cfg.pts_middle_encoder).eval()  # weights only (host tensors)
        sd = {k: v.numpy() for k, v in enc.state_dict().items()}

        def one():
            t0 = time.perf_counter()
            v, c, n = vox(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
            t1 = time.perf_counter()
            mean = cpu.hard_simple_vfe(v, n, 5)
            coors = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], 1)
            omodel.sparse_encoder(sd, dict(cfg.pts_middle_encoder), mean, coors, 1)
            return time.perf_counter() - t0, t1 - t0
        desc = ('hard_voxelize = %s; spconv-2.x SparseEncoder arithmetic = oracle/c port (spconv is not '
                'vendored in the reference), OpenMP' % ('reference CPU op (oracle/_ref)' if use_ref else 'oracle port'))
        return one, cpu.num_threads(), kind, desc, int(pts.shape[0])

    import msmdfusion_b200 as m
    det = m.MSMDFusionDetector(**{k: cfg[k] for k in (
        'pts_voxel_layer', 'pts_voxel_encoder', 'pts_middle_encoder', 'multimodal_middle_encoder',
        'spatial_shapes', 'downscale_factors', 'fps_num_list', 'radius_list', 'max_cluster_samples_list',
        'dist_thresh_list')}).eval()
    with torch.no_grad():
        det.score_net[0].weight.mul_(0.2)
        det.score_net[0].bias.fill_(0.05)
    sd = {k: v.numpy() for k, v in det.state_dict().items()}
    meta = synthetic.camera_scene(seed, pts)
    fpn = synthetic.fpn_features(seed, batch=1)
    rng = np.random.RandomState(77)
    dummies = [rng.rand(1, c).astype(np.float32) for c in cfg.multimodal_middle_encoder['in_channels_3D']]

    def one():
        t0 = time.perf_counter()
        vox(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)   # the reference's own CPU op, timed alone
        t1 = time.perf_counter()
        omodel.extract_voxel_space(sd, cfg, [pts], fpn, [meta], dummies)
        t2 = time.perf_counter()
        # the fusion pass voxelises the LiDAR points itself (oracle port); the separately timed leg is
        # reported, not added
        return t2 - t1, t1 - t0
    desc = ('the whole voxel-space fusion forward on the CPU (oracle.model.extract_voxel_space: hard_voxelize, '
            'SparseEncoder, depth-aware compression [torch CPU conv2d], 4x lift + voxelize + modality split, GMA '
            'encoder with FPS / ball query, dense): C port with OpenMP + numpy; spconv-2.x is not vendored in the '
            'reference, so its arithmetic is the port\'s')
    return one, cpu.num_threads(), 'port', desc, int(pts.shape[0])


def bench_config(args, world, points, virtual_points):
    """The `config` object both arms print (identical for the same command line): only what the
    command line and the seeded synthetic scene determine.  Run-dependent facts (voxel count, settle
    steps) are top-level keys of our arm's line."""
    from msmdfusion_b200 import spconv as _spconv
    precision = args.precision or _spconv.CONV_PRECISION
    SPS = max(1, args.scenes_per_step) if args.workload == 'L' else 1
    return {'workload': workload_name(args.profile, args.workload), 'profile': args.profile,
            'arithmetic': ARITHMETIC[precision], 'points_per_scene': points,
            'virtual_points_per_scene': virtual_points, 'scenes_per_gpu_per_step': SPS,
            'parallelism': 'dp%d' % world,
            'l2': 'flushed between steps (256 MiB memset, outside the per-step events)',
            'weights': ('random init (spconv default); LiDAR encoder frozen, GMA encoder BN in '
                        'training mode' if args.workload == 'train' else 'random init (spconv default), BN eval')}


def scene_counts(args, seed=0):
    from msmdfusion_b200 import synthetic
    B = max(1, args.scenes_per_step) if args.workload == 'L' else 1
    pts = synthetic.lidar_scene(seed=seed * B, sweeps=1 if args.profile == 'S' else 10)
    if args.workload == 'L':
        return int(pts.shape[0]), 0
    meta = synthetic.camera_scene(seed, pts)
    fg = meta['foreground2D_info']
    return int(pts.shape[0]), int(sum(len(x) for x in fg['fg_pixels']))


def main():
    args = parse()
    if os.environ.get('MSMD_BENCH_HANG_DUMP'):   # diagnostic: Python stacks of every thread after that many seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['MSMD_BENCH_HANG_DUMP']), exit=False)
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        # the reference's CPU implementation of the path on the box's host cores; under torchrun
        # rank 0 alone runs and prints, the other ranks exit without work (contract)
        if rank != 0:
            return 0
        one, cores, kind, desc, npts = cpu_pass(args.profile, args.workload)
        npts, nvirt = scene_counts(args)
        for _ in range(args.warmup):
            one()
        times = [one()[0] for _ in range(args.steps)]   # exactly K timed steps, one whole scene each
        val = 1.0 / float(np.mean(times))
        line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 / val,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic',
                'config': bench_config(args, args.gpus, npts, nvirt),
                'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': kind,
                                 'sample': '%d whole-scene passes on rank 0 (one scene per step whatever N is: the '
                                           'host cores are shared by the N ranks); %s' % (len(times), desc)},
                'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        print(json.dumps({'error': 'bench.py needs a CUDA device; msmdfusion_b200 has no CPU fallback'}))
        return 1
    import torch.distributed as dist
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    res = run_ours(args, rank, world, device)

    # the other ranks are released BEFORE rank 0 spends seconds on the CPU baseline (they used to spin
    # in a barrier, which read as GPU activity on GPUs 1..N-1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:   # contract: on rank 0 at N=1 only
        one, cores, kind, desc, _ = cpu_pass(args.profile, args.workload)
        if args.workload == 'L':
            one()
        tt = [one() for _ in range(2 if (args.profile == 'S' and args.workload == 'L') else 1)]
        cpu_base = {'value': 1.0 / float(np.mean([t[0] for t in tt])), 'unit': UNIT, 'cores': cores,
                    'kind': kind, 'sample': '%d whole-scene passes of the same workload; %s; '
                    'reference CPU hard_voxelize leg alone %.1f ms on 1 core' % (len(tt), desc, 1e3 * np.mean([t[1] for t in tt]))}
    K = args.steps
    SPS = max(1, args.scenes_per_step) if args.workload == 'L' else 1
    value = world * K * SPS / (res['dev_ms'] * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': max(args.warmup, 3),
        'ms_per_step': res['dev_ms'] / K, 'step_ms': res['step_ms'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': bench_config(args, world, res['points'] // SPS, res['virtual_points']),
        'run': {'voxels_per_scene': res['voxels'] // SPS,
                'settle': '%d untimed steps (the first + %.0f ms of wall time) before the %d warm-up steps' % (res['settle_steps'], SETTLE_MS, max(args.warmup, 3))},
        'e2e': {'value': world * K * SPS / (res['e2e_ms'] * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': res['h2d'],
                'd2h_bytes_per_step': res['d2h'], 'step_ms': res['e2e_step_ms'],
                'note': ('pinned host points (+ packed virtual points for LC) -> H2D -> public modules '
                         '(Voxelization.forward_mean/SparseEncoder or MSMDFusionDetector.extract_voxel_space) '
                         '-> checksum of the BEV tensor read back.  Limitation: the BEV tensor itself '
                         '(B x 640 x 180 x 180 f32 = 83 MB for LC) stays on the device, where its consumer '
                         '(SPPModule / SECOND / TransFusionHead) runs; the FPN image features are device-resident '
                         'inputs (produced on the device by the image branch in the reference too)')},
        'gpu_launches': res['launches'],
        'allocator_during_timed_steps': res['alloc'],
        'clocks': res['clocks'],
        'roofline': res['roofline'],
        'cpu_baseline': cpu_base,
        'cuda_baseline': res.get('cuda_baseline'),
    }
    if res.get('exchange') is not None:   # --workload train: the exchange step of SURVEY 8(e)
        line['gradient_exchange'] = res['exchange']
    print(json.dumps(line))
    return 0


if __name__ == '__main__':
    sys.exit(main())
