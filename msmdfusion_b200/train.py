"""Train step of the voxel-space path (BASELINE.json configs[4]: MSMDFusion_nusc_voxel_LC train step,
one scene per GPU, NCCL gradient all-reduce).

Mirrors the reference's step as far as the hot path goes: ``tools/train.py:185-211`` freezes
the LiDAR components (parameters and BatchNorm statistics), ``MMDistributedDataParallel(
find_unused_parameters=True)`` all-reduces the gradients of the parameters that took part in the step
(``configs/MSMDFusion_nusc_voxel_LC.py:309``), the optimiser is AdamW with gradient clipping at 10
(``:282-286``).  ``TransFusionHead.loss`` is outside the hot path; the step takes the
loss as a callable on the tensor the head consumes.

B200-first: the gradients of all participating parameters are VIEWS of one flat fp32 buffer, so the
exchange step is ONE NCCL all-reduce over NVLink (no bucketing, no copies: ~10 M parameters = 38 MB, far
below the size where splitting would buy overlap), the clip is one norm over that buffer and the update
is the fused multi-tensor AdamW.  The buffer covers a STATIC parameter set (``trainable_parameters``): the
sub-modules of the GMA encoder that its ``forward`` reaches.  The blocks the reference builds but never calls,
``score_net`` and ``conv1x1_blocks`` (the virtual-point features are constants, ``MSMDFusion.py:462-464``)
are left out of the buffer and of the optimiser, which is what ``find_unused_parameters`` amounts to.
"""
import torch
import torch.distributed as dist


def freeze_lidar_components(detector):
    """tools/train.py:185-211: the parameters of the LiDAR voxel layer / voxel encoder / middle encoder stop
    requiring gradients and their BatchNorms get ``track_running_stats = False`` (``fix_bn``).  The model stays in
    TRAIN mode, so -- exactly as in the reference -- those BatchNorms normalise with the statistics of the current
    batch and never touch their running estimates (torch: ``bn_training`` is True in train mode, and the running
    buffers are not passed when ``track_running_stats`` is False).  Such a BatchNorm cannot be folded into the
    convolution (``spconv._bn_foldable``), so the frozen encoder runs conv -> torch BatchNorm -> ReLU module by
    module under ``no_grad``."""
    from torch import nn

    def fix_bn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.track_running_stats = False

    for name in ('pts_voxel_layer', 'pts_voxel_encoder', 'pts_middle_encoder'):
        mod = getattr(detector, name, None)
        if mod is None:
            continue
        for p in mod.parameters():
            p.requires_grad_(False)
        mod.apply(fix_bn)
    return detector


# Sub-modules of SparseMultiModalEncoderPaint that ``forward`` reaches (sparse_multimodal_encoder_painting.py:
# 325-459).  ``grouped_sp_conv_blocks_2D`` / ``_mix`` are built and never called (:142-156); ``score_net`` and
# ``conv1x1_blocks`` feed the virtual-point lift, which has no backward (MSMDFusion.py:462-464).
TRAINED_SUBMODULES = ('grouped_sp_conv_blocks_3D', 'aggregation_blocks', 'downscale_blocks', 'gate_control',
                      'cross_gate_control')


def trainable_parameters(detector):
    """The STATIC, rank-independent set of parameters the step optimises: every ``requires_grad`` parameter of
    the sub-modules of the GMA encoder that are on the call path.  Which of them actually receive a gradient in
    a given step is data dependent (a stage with no mixed voxel skips its ``gate_control``): those keep a zero
    gradient in the flat buffer, so all ranks exchange buffers of the same layout.  The reference gets the same
    effect from ``DDP(find_unused_parameters=True)``."""
    enc = getattr(detector, 'multimodal_middle_encoder', None)
    assert enc is not None, 'the train step optimises the GMA encoder (multimodal_middle_encoder)'
    named = [getattr(enc, name) for name in TRAINED_SUBMODULES if hasattr(enc, name)]
    if not named:   # an encoder with another layout (tests' stand-ins): everything it owns
        named = [enc]
    out = []
    for mod in named:
        out += [p for p in mod.parameters() if p.requires_grad]
    return out


class FlatGradients:
    """One flat fp32 buffer holding the gradients of ``params`` (each ``p.grad`` is a view of it)."""

    def __init__(self, params):
        self.params = [p for p in params]
        assert self.params, 'no parameters take part in the step'
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)   # autograd accumulates in place into the view
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """The exchange step: average over ranks with ONE collective (no-op for a single rank)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())

    def clip_(self, max_norm):
        """``clip_grad_norm_`` (mmcv ``grad_clip=dict(max_norm=10, norm_type=2)``) on the flat buffer,
        without a host synchronisation."""
        norm = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm


class VoxelSpaceTrainStep:
    """forward (train mode) -> loss -> backward -> gradient all-reduce -> clip -> AdamW."""

    def __init__(self, detector, loss_fn, lr=1e-4, weight_decay=0.01, grad_clip=10.0, precision=None):
        # precision: operand precision of the sparse convolutions during the step ('tf32x3' | 'bf16x3' | 'bf16',
        # spconv.CONV_PRECISION); None keeps the process-wide setting.  'bf16' is what BASELINE configs[4] names:
        # bf16 operands on the tensor cores, fp32 accumulation, fp32 master weights and optimiser state.
        self.precision = precision
        detector.train()                       # batch-statistics BatchNorm, max_voxels[0] (voxelize.py:103-114)
        self.det = freeze_lidar_components(detector)
        self.loss_fn = loss_fn
        self.last_stage_outs = None
        self.lr, self.weight_decay, self.grad_clip = lr, weight_decay, grad_clip
        self.grads = None
        self.opt = None
        # when a list, every step appends (start, end) CUDA events around the gradient exchange: the all-reduce is
        # issued on the compute stream after the backward pass, so its duration IS its exposed time
        self.exchange_events = None

    def _layout(self):
        """Lay out the flat gradient buffer over the static parameter set BEFORE the first backward, so that
        every rank owns the same layout whatever its first scene looks like."""
        used = trainable_parameters(self.det)
        self.grads = FlatGradients(used)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            n = torch.tensor([self.grads.flat.numel(), -self.grads.flat.numel()], dtype=torch.int64,
                             device=self.grads.flat.device)
            dist.all_reduce(n, op=dist.ReduceOp.MAX)
            assert int(n[0]) == -int(n[1]) == self.grads.flat.numel(), \
                'ranks disagree on the gradient buffer layout'
        self.opt = torch.optim.AdamW(used, lr=self.lr, weight_decay=self.weight_decay,
                                     fused=used[0].is_cuda)

    def exchange_ms(self):
        """Sum of the recorded gradient-exchange durations (call after a synchronize); clears the list."""
        ev, self.exchange_events = self.exchange_events or [], []
        return sum(a.elapsed_time(b) for a, b in ev), len(ev)

    def forward_loss(self, points, img_feats, img_metas):
        bev, self.last_stage_outs = self.det.extract_voxel_space(points, img_feats, img_metas)
        return self.loss_fn(bev)

    def __call__(self, points, img_feats, img_metas):
        from . import spconv
        if self.grads is None:
            self._layout()
        self.grads.zero()
        saved = spconv.CONV_PRECISION
        if self.precision:
            spconv.CONV_PRECISION = self.precision
        try:
            loss = self.forward_loss(points, img_feats, img_metas)
            loss.backward()
        finally:
            spconv.CONV_PRECISION = saved
        timed = self.exchange_events is not None and self.grads.flat.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.grads.all_reduce_mean()
        if timed:
            e1.record()
            self.exchange_events.append((e0, e1))
        self.grads.clip_(self.grad_clip)
        self.opt.step()
        return loss.detach()
