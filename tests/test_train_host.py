"""CPU tests of the train step's HOST logic (config 5): the autograd functions of
``msmdfusion_b200/autograd.py`` and the module plumbing around them, with the C-ABI calls replaced by
contract-level stand-ins built on the oracle (tests/_cpu_ops.py), against torch-native autograd in
float64.  The CUDA kernels behind the same calls are checked by the ``-m gpu`` tests."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import msmdfusion_b200 as m  # noqa: F401  (registers the modules)
from msmdfusion_b200 import functional as Fsp
from msmdfusion_b200 import spconv
from msmdfusion_b200.sparse_block import SparseBasicBlock, make_sparse_convmodule

import _cpu_ops


@pytest.fixture()
def cpu_ops(monkeypatch):
    _cpu_ops.install(monkeypatch)


def random_sparse(seed, batch, shape, n, c):
    rng = np.random.default_rng(seed)
    D, H, W = shape
    lin = rng.choice(batch * D * H * W, size=n, replace=False)
    idx = np.stack([lin // (D * H * W), (lin // (H * W)) % D, (lin // W) % H, lin % W], 1).astype(np.int32)
    return torch.from_numpy(idx), torch.from_numpy(rng.standard_normal((n, c)).astype(np.float32))


def dense_of(features, indices, shape, batch):
    """(B,C,D,H,W) float64, differentiable w.r.t. features."""
    i = indices.long()
    x = torch.zeros(batch, *shape, features.shape[1], dtype=torch.float64)
    return x.index_put((i[:, 0], i[:, 1], i[:, 2], i[:, 3]), features.double()).permute(0, 4, 1, 2, 3)


def rows_of(dense, indices):
    i = indices.long()
    return dense[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]]


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


@pytest.mark.parametrize('kind,ksize,stride,padding', [('subm', 3, 1, 1), ('subm', (3, 1, 1), 1, 0),
                                                       ('conv', 3, 2, 1), ('conv', (3, 1, 1), (2, 1, 1), 0),
                                                       ('conv', 3, 2, (0, 1, 1))])
def test_conv_module_gradients_vs_dense_autograd(cpu_ops, kind, ksize, stride, padding):
    shape, batch, cin, cout = [7, 12, 10], 2, 5, 6
    idx, feat = random_sparse(0, batch, shape, 300, cin)
    torch.manual_seed(0)
    cls = spconv.SubMConv3d if kind == 'subm' else spconv.SparseConv3d
    conv = cls(cin, cout, ksize, stride=stride, padding=padding, bias=False)
    f = feat.clone().requires_grad_(True)
    out = conv(spconv.SparseConvTensor(f, idx, shape, batch))
    g = torch.randn(out.features.shape, generator=torch.Generator().manual_seed(1))
    (out.features * g).sum().backward()

    fr = feat.clone().double().requires_grad_(True)
    wr = conv.weight.detach().double().requires_grad_(True)
    ks, st, pd = spconv.expand_nd(3, ksize), spconv.expand_nd(3, stride), spconv.expand_nd(3, padding)
    if kind == 'subm':
        st, pd = [1, 1, 1], [k // 2 for k in ks]
    y = F.conv3d(dense_of(fr, idx, shape, batch), wr.permute(0, 4, 1, 2, 3), stride=st, padding=pd)
    yr = rows_of(y, out.indices)
    assert rel(out.features, yr) < 1e-5
    (yr * g.double()).sum().backward()
    assert rel(f.grad, fr.grad) < 1e-5
    assert rel(conv.weight.grad, wr.grad) < 1e-5


def test_frozen_weight_or_constant_input(cpu_ops):
    """needs_input_grad is honoured: a frozen conv still passes the data gradient, a conv on constant
    features still gets its weight gradient."""
    shape, batch = [5, 8, 8], 1
    idx, feat = random_sparse(2, batch, shape, 120, 4)
    conv = spconv.SubMConv3d(4, 4, 3, padding=1, bias=False)
    conv.weight.requires_grad_(False)
    f = feat.clone().requires_grad_(True)
    conv(spconv.SparseConvTensor(f, idx, shape, batch)).features.sum().backward()
    assert f.grad is not None and conv.weight.grad is None
    conv.weight.requires_grad_(True)
    conv(spconv.SparseConvTensor(feat, idx, shape, batch)).features.sum().backward()
    assert conv.weight.grad is not None
    with torch.no_grad():  # inference keeps the fused path
        out = conv(spconv.SparseConvTensor(feat, idx, shape, batch))
    assert not out.features.requires_grad


def torch_ref_conv(features, weight, pair):
    """Torch-native (differentiable) restatement: sum_k x[pair[k]] @ W_k^T with missing rows zero."""
    cout, cin = weight.shape[0], weight.shape[-1]
    w3 = weight.reshape(cout, -1, cin)
    xz = torch.cat([features, features.new_zeros(1, cin)], 0)
    out = features.new_zeros(pair.shape[1], cout)
    for k in range(pair.shape[0]):
        p = pair[k].long()
        p = torch.where(p < 0, torch.full_like(p, features.shape[0]), p)
        out = out + xz[p] @ w3[:, k].T
    return out


def test_basic_block_train_mode_gradients(cpu_ops):
    """SparseBasicBlock in training mode (batch-statistics BatchNorm1d over the active rows, in-place
    ReLUs, residual): every parameter gradient and the input gradient against a float64 torch-native
    restatement of the block."""
    shape, batch, c = [5, 10, 10], 2, 8
    idx, feat = random_sparse(3, batch, shape, 260, c)
    torch.manual_seed(3)
    blk = SparseBasicBlock(c, c, norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01),
                           conv_cfg=dict(type='SubMConv3d')).train()
    f = feat.clone().requires_grad_(True)
    out = blk(spconv.SparseConvTensor(f, idx, shape, batch))
    g = torch.randn(out.features.shape, generator=torch.Generator().manual_seed(4))
    (out.features * g).sum().backward()

    pair = _cpu_ops.rulebook_subm(idx, _cpu_ops.CpuGrid(idx, shape, batch), [3, 3, 3])
    fr = feat.clone().double().requires_grad_(True)
    p = {k: v.detach().double().requires_grad_(True) for k, v in blk.named_parameters()}

    def bn(x, w, b):
        return F.batch_norm(x, None, None, w, b, True, 0.0, 1e-3)
    y = torch.relu(bn(torch_ref_conv(fr, p['conv1.weight'], pair), p['bn1.weight'], p['bn1.bias']))
    y = torch.relu(bn(torch_ref_conv(y, p['conv2.weight'], pair), p['bn2.weight'], p['bn2.bias']) + fr)
    assert rel(out.features, y) < 1e-5
    (y * g.double()).sum().backward()
    assert rel(f.grad, fr.grad) < 2e-5
    for k, v in blk.named_parameters():
        assert v.grad is not None, k
        assert rel(v.grad, p[k].grad) < 2e-5, k


def test_convmodule_sparse_add_dense_chain_gradients(cpu_ops):
    """downscale SparseConv3d module -> Fsp.sparse_add -> second strided conv -> dense(): the tail of
    SparseMultiModalEncoderPaint.forward (:452-456) + MSMDFusion.py:441, gradients vs float64."""
    shape, batch = [9, 12, 12], 1
    ia, fa = random_sparse(5, batch, shape, 200, 6)
    ib, fb = random_sparse(6, batch, [5, 6, 6], 90, 10)
    torch.manual_seed(5)
    norm = dict(type='BN1d', eps=1e-3, momentum=0.01)
    down1 = make_sparse_convmodule(6, 10, 3, 'ds1', stride=2, padding=1, conv_type='SparseConv3d', norm_cfg=norm).train()
    down2 = make_sparse_convmodule(10, 4, [3, 1, 1], 'ds2', stride=[2, 1, 1], padding=0, conv_type='SparseConv3d',
                                   norm_cfg=norm).train()
    a = fa.clone().requires_grad_(True)
    b = fb.clone().requires_grad_(True)
    x1 = down1(spconv.SparseConvTensor(a, ia, shape, batch))
    assert x1.spatial_shape == [5, 6, 6]
    s = Fsp.sparse_add(x1, spconv.SparseConvTensor(b, ib, [5, 6, 6], batch))
    x2 = down2(s)
    d = x2.dense()
    g = torch.randn(d.shape, generator=torch.Generator().manual_seed(7))
    (d * g).sum().backward()

    ar, br = fa.clone().double().requires_grad_(True), fb.clone().double().requires_grad_(True)
    p1 = {k: v.detach().double().requires_grad_(True) for k, v in down1.named_parameters()}
    p2 = {k: v.detach().double().requires_grad_(True) for k, v in down2.named_parameters()}
    grid = _cpu_ops.CpuGrid(ia, shape, batch)
    oi1, pair1, g1 = _cpu_ops.rulebook_conv(ia, grid, [3, 3, 3], [2, 2, 2], [1, 1, 1])
    y1 = torch.relu(F.batch_norm(torch_ref_conv(ar, p1['0.weight'], pair1), None, None, p1['1.weight'],
                                 p1['1.bias'], True, 0.0, 1e-3))
    # sparse_add restated: scatter both operands into the union's ascending rows
    ra, rb = _cpu_ops.grid_rows(oi1, _cpu_ops.CpuGrid(s.indices, [5, 6, 6], batch)), \
        _cpu_ops.grid_rows(ib, _cpu_ops.CpuGrid(s.indices, [5, 6, 6], batch))
    assert (ra >= 0).all() and (rb >= 0).all()
    u = torch.zeros(s.indices.shape[0], 10, dtype=torch.float64).index_add(0, ra, y1).index_add(0, rb, br)
    assert rel(s.features, u) < 1e-5
    oi2, pair2, g2 = _cpu_ops.rulebook_conv(s.indices, _cpu_ops.CpuGrid(s.indices, [5, 6, 6], batch), [3, 1, 1],
                                            [2, 1, 1], [0, 0, 0])
    y2 = torch.relu(F.batch_norm(torch_ref_conv(u, p2['0.weight'], pair2), None, None, p2['1.weight'],
                                 p2['1.bias'], True, 0.0, 1e-3))
    dr = dense_of(y2, oi2, g2.spatial_shape, batch)
    assert rel(d, dr) < 1e-5
    (dr * g.double()).sum().backward()
    assert rel(a.grad, ar.grad) < 2e-5 and rel(b.grad, br.grad) < 2e-5
    for mod, pr in ((down1, p1), (down2, p2)):
        for k, v in mod.named_parameters():
            assert rel(v.grad, pr[k].grad) < 2e-5, k


# --------------------------------------------------------------------------------------
# the whole GMA encoder in training mode
# --------------------------------------------------------------------------------------
def _encoder_inputs(seed, batch):
    """Four scales of (3-D voxels, 2-D voxels) with deliberate overlap, modality-split by the oracle."""
    from oracle import cpu
    shapes = ([41, 24, 24], [21, 12, 12], [11, 6, 6], [5, 3, 3])
    c3s = (4, 8, 8, 8)
    rng = np.random.default_rng(seed)
    v3l, v2l, s3l, s2l = [], [], [], []
    for shape, c3 in zip(shapes, c3s):
        cells = int(np.prod(shape))
        n3, n2 = max(12, cells // 14), max(16, cells // 9)
        per_b = []
        for kind, n, c in (('3', n3, c3), ('2', n2, 64)):
            idx = []
            for b in range(batch):
                lin = np.sort(rng.choice(cells, size=min(n, cells), replace=False))
                D, H, W = shape
                idx.append(np.stack([np.full_like(lin, b), lin // (H * W), (lin // W) % H, lin % W], 1))
            idx = np.concatenate(idx, 0).astype(np.int32)
            per_b.append((idx, rng.standard_normal((idx.shape[0], c)).astype(np.float32)))
        (i3, f3), (i2, f2) = per_b
        m3, m2, s3, s2 = cpu.voxel_modality_split(i3, i2, batch)
        assert s3.shape[0] > 0
        v3l.append((torch.from_numpy(m3), torch.from_numpy(f3), shape))
        v2l.append((torch.from_numpy(m2), torch.from_numpy(f2), shape))
        s3l.append(torch.from_numpy(s3)); s2l.append(torch.from_numpy(s2))
    return v3l, v2l, s3l, s2l


def _run_encoder(enc, inputs, batch, seed):
    v3l, v2l, s3l, s2l = inputs
    v3 = [spconv.SparseConvTensor(f.clone(), i.clone(), s, batch) for i, f, s in v3l]
    v2 = [spconv.SparseConvTensor(f.clone(), i.clone(), s, batch) for i, f, s in v2l]
    torch.manual_seed(seed)  # the dummy embeddings of :372 come from the CPU generator
    outs = enc(v3, v2, s3l, s2l, [6, 6, 6, 6], [6, 3, 2, 1], [20, 10, 5, 3], [13.3, 6.6, 3.3, 1.6])
    return outs, outs[-1].dense()


@pytest.mark.parametrize('batch', [1, 2])
def test_gma_encoder_train_mode_gradients_vs_torch_native(cpu_ops, monkeypatch, batch):
    """SparseMultiModalEncoderPaint.forward in training mode (config 5), gradients of the dense BEV
    output w.r.t. every parameter: the package's autograd functions (C-ABI contract stand-ins) against
    the same module graph with torch-native differentiable restatements of the three ops.  Also pins the
    gradient scope of SURVEY 3.3: the blocks the reference builds but never calls get no gradient."""
    from msmdfusion_b200 import autograd as ag
    from msmdfusion_b200 import fusion_encoder as fe
    from oracle import cpu

    def fps_nn_fast(query, key, fps_num, radius, nsample, thresh, base=0):
        out = cpu.fps_nn_fast(query.numpy(), key.numpy(), fps_num, radius, nsample, thresh)
        out = np.where(out >= 0, out + base, out)
        return torch.from_numpy(out)
    monkeypatch.setattr(fe, 'fps_nn_fast', fps_nn_fast)

    torch.manual_seed(11)
    enc = fe.SparseMultiModalEncoderPaint(in_channels_3D=(4, 8, 8, 8), in_channels_2D=(64,) * 4,
                                          out_channels=(8, 8, 8, 8), padding=(1, 1, [0, 1, 1], 0)).train()
    inputs = _encoder_inputs(20 + batch, batch)

    outs, d = _run_encoder(enc, inputs, batch, 5)
    assert [o.spatial_shape for o in outs] == [[21, 12, 12], [11, 6, 6], [5, 3, 3], [2, 3, 3]]
    g = torch.randn(d.shape, generator=torch.Generator().manual_seed(9))
    (d * g).sum().backward()
    got = {k: (None if v.grad is None else v.grad.clone()) for k, v in enc.named_parameters()}
    enc.zero_grad(set_to_none=True)

    # torch-native restatements of the three custom functions
    class NativeConv:
        @staticmethod
        def apply(features, weight, packed, rb):
            return torch_ref_conv(features, weight, rb['pair_fwd'])

    class NativeDense:
        @staticmethod
        def apply(features, indices, shape, bsz):
            x = torch.zeros(bsz, *shape, features.shape[1])
            i = indices.long()
            return x.index_put((i[:, 0], i[:, 1], i[:, 2], i[:, 3]), features).permute(0, 4, 1, 2, 3).contiguous()

    class NativeAdd:
        @staticmethod
        def apply(fa, fb, ia, ib, shape, bsz, holder):
            oi, _, grid = _cpu_ops.sparse_add(ia, fa, ib, fb, shape, bsz)
            holder['out_idx'], holder['grid'] = oi, grid
            ra, rb_ = _cpu_ops.grid_rows(ia, grid), _cpu_ops.grid_rows(ib, grid)
            return torch.zeros(oi.shape[0], fa.shape[1]).index_add(0, ra, fa).index_add(0, rb_, fb)
    monkeypatch.setattr(ag, 'SparseConvFunction', NativeConv)
    monkeypatch.setattr(ag, 'ToDenseFunction', NativeDense)
    monkeypatch.setattr(ag, 'SparseAddFunction', NativeAdd)
    outs_r, d_r = _run_encoder(enc, inputs, batch, 5)
    assert rel(d.detach(), d_r.detach()) < 1e-5
    (d_r * g).sum().backward()

    with_grad, without, errs = [], [], {}
    for k, v in enc.named_parameters():
        if v.grad is None:
            assert got[k] is None, k
            without.append(k)
            continue
        assert got[k] is not None, k
        errs[k] = rel(got[k], v.grad)
        with_grad.append(k)
    # gradient scope: 16 convs (+ their BNs) and the 8 gate Linears train; the _2D / _mix blocks do not
    assert all(k.startswith(('grouped_sp_conv_blocks_2D', 'grouped_sp_conv_blocks_mix')) for k in without), without
    assert sum(k.endswith('.weight') and 'conv' in k.split('.')[-2] or k.endswith('.0.weight') and
               ('downscale' in k or 'grouped_sp_conv_blocks_3D' in k) for k in with_grad) >= 16
    assert len([k for k in with_grad if 'gate_control' in k]) == 16
    print({k: round(v, 6) for k, v in errs.items()})
    assert max(errs.values()) < 1e-4, max(errs, key=errs.get)


def test_subm_gradient_with_duplicate_coordinates(cpu_ops):
    """An index list that repeats coordinates (the GMA conv's unified voxel list does when the float32
    keys of voxel_modality_split collide): neighbours read the LARGEST row of a coordinate, every copy
    is an output.  The mirrored-offset data gradient must fold the copies' output gradients onto that
    row and give the unread copies zero."""
    shape, batch, c = [5, 8, 8], 1, 4
    idx, feat = random_sparse(8, batch, shape, 100, c)
    idx = torch.cat([idx, idx[:15], idx[5:10]], 0)          # up to three copies of a coordinate
    feat = torch.cat([feat, feat[:15] + 1, feat[5:10] - 1], 0)
    torch.manual_seed(8)
    conv = spconv.SubMConv3d(c, 6, 3, padding=1, bias=False)
    f = feat.clone().requires_grad_(True)
    out = conv(spconv.SparseConvTensor(f, idx, shape, batch))
    g = torch.randn(out.features.shape, generator=torch.Generator().manual_seed(2))
    (out.features * g).sum().backward()
    pair = _cpu_ops.rulebook_subm(idx, _cpu_ops.CpuGrid(idx, shape, batch), [3, 3, 3])
    assert (pair[13] != torch.arange(idx.shape[0])).sum() == 20   # the unread copies
    fr = feat.clone().double().requires_grad_(True)
    wr = conv.weight.detach().double().requires_grad_(True)
    y = torch_ref_conv(fr, wr, pair)
    assert rel(out.features, y) < 1e-5
    (y * g.double()).sum().backward()
    assert rel(f.grad, fr.grad) < 1e-5 and rel(conv.weight.grad, wr.grad) < 1e-5
    assert (fr.grad[pair[13] != torch.arange(idx.shape[0])] == 0).all()


def test_gma_encoder_single_sample_path_gradients(cpu_ops, monkeypatch):
    """The sync-free one-sample-per-GPU path (``_grouped_sparse_conv_b1``: device-scan row lists, the
    zero-padded concatenation written by slice assignment) gives the same output and the same parameter
    gradients as the generic path in training mode."""
    from msmdfusion_b200 import fusion_encoder as fe
    from msmdfusion_b200 import ops
    from oracle import cpu

    def fps_nn_fast(query, key, fps_num, radius, nsample, thresh, base=0):
        out = cpu.fps_nn_fast(query.numpy(), key.numpy(), fps_num, radius, nsample, thresh)
        return torch.from_numpy(np.where(out >= 0, out + base, out))
    monkeypatch.setattr(fe, 'fps_nn_fast', fps_nn_fast)
    monkeypatch.setattr(ops, 'compact_unflagged',
                        lambda flags, count: torch.nonzero(flags == 0).flatten()[:int(count)])
    torch.manual_seed(12)
    enc = fe.SparseMultiModalEncoderPaint(in_channels_3D=(4, 8, 8, 8), in_channels_2D=(64,) * 4,
                                          out_channels=(8, 8, 8, 8), padding=(1, 1, [0, 1, 1], 0)).train()
    inputs = _encoder_inputs(31, 1)
    results = []
    for single in (False, True):
        v3l, v2l, s3l, s2l = inputs
        v3 = [spconv.SparseConvTensor(f.clone(), i.clone(), s, 1) for i, f, s in v3l]
        v2 = [spconv.SparseConvTensor(f.clone(), i.clone(), s, 1) for i, f, s in v2l]
        if single:  # what MSMDFusionDetector.voxel_modality_split leaves on the tensors
            for t in v3 + v2:
                t._mix = t.indices[:, 1].contiguous().int()
                t._bzyx = t.indices[:, [0, 2, 3, 4]].contiguous()
        torch.manual_seed(5)
        outs = enc(v3, v2, s3l, s2l, [6, 6, 6, 6], [6, 3, 2, 1], [20, 10, 5, 3], [13.3, 6.6, 3.3, 1.6])
        d = outs[-1].dense()
        g = torch.randn(d.shape, generator=torch.Generator().manual_seed(9))
        enc.zero_grad(set_to_none=True)
        (d * g).sum().backward()
        results.append((d.detach(), {k: v.grad.clone() for k, v in enc.named_parameters() if v.grad is not None}))
    (d0, g0), (d1, g1) = results
    assert rel(d1, d0) < 1e-5 and g0.keys() == g1.keys()
    for k in g0:
        assert rel(g1[k], g0[k]) < 1e-4, k


def test_voxel_space_train_step_on_a_stand_in_detector():
    """VoxelSpaceTrainStep: LiDAR components frozen the reference's way (tools/train.py:185-211: requires_grad off,
    BatchNorm.track_running_stats off, the model stays in train mode => batch statistics, running estimates never
    touched), parameters the step never reaches
    stay out of the flat gradient buffer and of the optimiser (find_unused_parameters semantics),
    gradients are views of one buffer, and three steps equal clip_grad_norm_ + torch AdamW by hand."""
    from msmdfusion_b200 import train

    class Det(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.pts_voxel_encoder = torch.nn.Identity()
            self.pts_middle_encoder = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.BatchNorm1d(8))
            self.multimodal_middle_encoder = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.BatchNorm1d(8),
                                                                 torch.nn.ReLU(), torch.nn.Linear(8, 4))
            self.score_net = torch.nn.Linear(3, 1)          # reached only through a detached input

        def extract_voxel_space(self, points, img_feats, img_metas):
            with torch.no_grad():
                x = self.pts_middle_encoder(points[0])
                gate = self.score_net(img_feats)
            return self.multimodal_middle_encoder(x * gate), ['stage_outs']

    def make():
        torch.manual_seed(0)
        return Det()
    det, ref = make(), make()
    pts = torch.randn(32, 6, generator=torch.Generator().manual_seed(1))
    img = torch.randn(32, 3, generator=torch.Generator().manual_seed(2))
    step = train.VoxelSpaceTrainStep(det, lambda bev: bev.pow(2).mean() * 50, lr=1e-2, weight_decay=0.01, grad_clip=1.0)
    assert det.pts_middle_encoder.training and det.multimodal_middle_encoder.training
    assert not det.pts_middle_encoder[1].track_running_stats
    assert not any(p.requires_grad for p in det.pts_middle_encoder.parameters())
    rm0 = det.pts_middle_encoder[1].running_mean.clone()

    ref.train()
    train.freeze_lidar_components(ref)
    used = list(ref.multimodal_middle_encoder.parameters())
    opt = torch.optim.AdamW(used, lr=1e-2, weight_decay=0.01)
    for it in range(3):
        loss = step([pts], img, None)
        opt.zero_grad(set_to_none=True)
        rl = ref.extract_voxel_space([pts], img, None)[0].pow(2).mean() * 50
        rl.backward()
        torch.nn.utils.clip_grad_norm_(used, 1.0)
        opt.step()
        assert abs(float(loss) - float(rl.detach())) < 1e-5 * max(1.0, abs(float(rl.detach()))), it
    assert step.last_stage_outs == ['stage_outs']
    assert len(step.grads.params) == len(used) and det.score_net.weight.grad is None
    lo, hi = step.grads.flat.data_ptr(), step.grads.flat.data_ptr() + step.grads.flat.numel() * 4
    assert all(lo <= p.grad.data_ptr() < hi for p in step.grads.params)
    for (name, a), b in zip(det.multimodal_middle_encoder.named_parameters(), used):
        if name == '0.bias':
            # a bias in front of a train-mode BatchNorm has an analytically ZERO gradient: what autograd returns is
            # rounding noise (~1e-8), which Adam's normalisation turns into +-lr steps -- not comparable
            continue
        assert torch.allclose(a, b, atol=1e-5), name
    assert torch.equal(det.pts_middle_encoder[0].weight, ref.pts_middle_encoder[0].weight)
    # fix_bn semantics: the frozen BatchNorm used the statistics of the batch (its output is normalised over the
    # 32 rows) and never moved its running estimates
    assert torch.equal(det.pts_middle_encoder[1].running_mean, rm0)
    with torch.no_grad():
        y = det.pts_middle_encoder(pts)
    assert float(y.mean(0).abs().max()) < 1e-5 and float((y.var(0, unbiased=False) - 1).abs().max()) < 1e-2
