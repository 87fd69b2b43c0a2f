"""CPU oracle of the module-level path (TEST INFRASTRUCTURE; see oracle/README.md).

Restates, on top of oracle/cpu.py:
* ``SparseEncoder.forward``  (mmdet3d/models/middle_encoders/sparse_encoder.py:96-209)
* ``SparseBasicBlock.forward`` / ``make_sparse_convmodule`` (mmdet3d/ops/sparse_block.py:103-191)
driven by a plain ``state_dict`` (numpy arrays) and the config dict -- it does not import
the product package.
"""
import numpy as np

from . import cpu


class SpTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features = features
        self.indices = indices
        self.spatial_shape = list(spatial_shape)
        self.batch_size = batch_size


def _np(sd, key):
    v = sd[key]
    return v.detach().cpu().numpy() if hasattr(v, 'detach') else np.asarray(v)


def _bn(sd, prefix, x, eps):
    return cpu.batchnorm_eval(x, _np(sd, prefix + '.weight'), _np(sd, prefix + '.bias'),
                              _np(sd, prefix + '.running_mean'), _np(sd, prefix + '.running_var'), eps)


def subm_conv(sd, key, x, ksize=3, dilation=1):
    pair = cpu.subm_rulebook(x.indices, x.spatial_shape, ksize, dilation)
    out = cpu.spconv_fwd(x.features, _np(sd, key + '.weight'), pair)
    return SpTensor(out, x.indices, x.spatial_shape, x.batch_size)


def strided_conv(sd, key, x, ksize, stride, padding, dilation=1):
    out_idx, pair, out_shape = cpu.conv_rulebook(x.indices, x.spatial_shape, ksize, stride, padding,
                                                 dilation)
    out = cpu.spconv_fwd(x.features, _np(sd, key + '.weight'), pair)
    return SpTensor(out, out_idx, out_shape, x.batch_size)


def convmodule(sd, prefix, x, conv_type, ksize, stride, padding, eps):
    """SparseSequential(conv, BN1d, ReLU) -> children 0,1,2 (sparse_block.py:161-190)."""
    if conv_type == 'SubMConv3d':
        y = subm_conv(sd, prefix + '.0', x, ksize)
    else:
        y = strided_conv(sd, prefix + '.0', x, ksize, stride, padding)
    y.features = np.maximum(_bn(sd, prefix + '.1', y.features, eps), 0)
    return y


def basic_block(sd, prefix, x, eps):
    """sparse_block.py:103-126."""
    identity = x.features
    out = subm_conv(sd, prefix + '.conv1', x)
    out.features = np.maximum(_bn(sd, prefix + '.bn1', out.features, eps), 0)
    out = subm_conv(sd, prefix + '.conv2', out)
    out.features = _bn(sd, prefix + '.bn2', out.features, eps)
    out.features = np.maximum(out.features + identity, 0)
    return out


def sparse_encoder(sd, cfg, voxel_features, coors, batch_size, prefix=''):
    """sparse_encoder.py:96-133.  cfg: the ``pts_middle_encoder`` dict of the config file.
    Returns (spatial_features (B, C*D, H, W), [SpTensor x5])."""
    eps = cfg.get('norm_cfg', dict(eps=1e-3)).get('eps', 1e-3)
    block_type = cfg.get('block_type', 'conv_module')
    enc_ch = cfg.get('encoder_channels', ((16, ), (32, 32, 32), (64, 64, 64), (64, 64, 64)))
    enc_pad = cfg.get('encoder_paddings', ((1, ), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)))
    x = SpTensor(np.ascontiguousarray(voxel_features, np.float32), np.ascontiguousarray(coors, np.int32),
                 cfg['sparse_shape'], batch_size)
    x = convmodule(sd, prefix + 'conv_input', x, 'SubMConv3d', 3, 1, 1, eps)
    feats = [x]
    for i, blocks in enumerate(enc_ch):
        for j, _ in enumerate(tuple(blocks)):
            padding = tuple(enc_pad[i])[j]
            name = f'{prefix}encoder_layers.encoder_layer{i + 1}.{j}'
            if i != 0 and j == 0 and block_type == 'conv_module':
                x = convmodule(sd, name, x, 'SparseConv3d', 3, 2, padding, eps)
            elif block_type == 'basicblock':
                if j == len(blocks) - 1 and i != len(enc_ch) - 1:
                    x = convmodule(sd, name, x, 'SparseConv3d', 3, 2, padding, eps)
                else:
                    x = basic_block(sd, name, x, eps)
            else:
                x = convmodule(sd, name, x, 'SubMConv3d', 3, 1, padding, eps)
        feats.append(x)
    out = convmodule(sd, prefix + 'conv_out', feats[-1], 'SparseConv3d', (3, 1, 1), (2, 1, 1), 0, eps)
    dense = cpu.dense(out.indices, out.features, out.spatial_shape, batch_size)
    N, C, D, H, W = dense.shape
    return dense.reshape(N, C * D, H, W), feats, out


def voxelize_batch(points_list, voxel_size, coors_range, max_points, max_voxels):
    """MVXTwoStageDetector.voxelize (mmdet3d/models/detectors/MSMDFusion.py:462-491)."""
    voxels, coors, nums = [], [], []
    for i, p in enumerate(points_list):
        v, c, n = cpu.hard_voxelize(p, voxel_size, coors_range, max_points, max_voxels)
        voxels.append(v)
        nums.append(n)
        coors.append(np.concatenate([np.full((c.shape[0], 1), i, np.int32), c], 1))
    return np.concatenate(voxels, 0), np.concatenate(nums, 0), np.concatenate(coors, 0)


# --------------------------------------------------------------------------------------
# Fusion side: get_foreground2D -> fetch_2D_voxels -> voxel_modality_split (cpu.py) ->
# SparseMultiModalEncoderPaint
# --------------------------------------------------------------------------------------
def _linear_relu(sd, prefix, x):
    w, b = _np(sd, prefix + '.weight').astype(np.float32), _np(sd, prefix + '.bias').astype(np.float32)
    return np.maximum(x.astype(np.float32) @ w.T + b, 0).astype(np.float32)


def depth_canvas(img_metas, H, W):
    """mmdet3d/models/detectors/MSMDFusion.py:335-356: (B*6, 1, H, W) sparse depth map; duplicate
    pixels -> the last real point in input order wins (sequential index_put_)."""
    B = len(img_metas)
    ncam = len(img_metas[0]['foreground2D_info']['fg_real_pixels'])
    canvas = np.zeros((B, ncam, H, W), np.float32)
    for i, meta in enumerate(img_metas):
        for j in range(ncam):
            r = np.asarray(meta['foreground2D_info']['fg_real_pixels'][j], np.float32)
            coors = r[:, :2].astype(np.int64)  # .long(): truncation
            for k in range(r.shape[0]):
                canvas[i, j, coors[k, 1], coors[k, 0]] = r[k, 2]
    return canvas.reshape(B * ncam, 1, H, W)


def get_foreground2d(img_feats, img_metas, score_w, score_b):
    """MSMDFusion.py:169-238.  img_feats (B*ncam, C, h, w) -> list (len B) of (M_b, 15+C)."""
    B = len(img_metas)
    BN, C, h, w = img_feats.shape
    ncam = BN // B
    input_w = img_metas[0]['input_shape'][-1]
    out = []
    for b, meta in enumerate(img_metas):
        info = meta['foreground2D_info']
        cams = []
        for v in range(ncam):
            pix = np.asarray(info['fg_pixels'][v], np.float32).reshape(-1, 3)
            pts = np.asarray(info['fg_points'][v], np.float32)
            pts = pts.reshape(pix.shape[0], -1) if pix.shape[0] else np.zeros((0, 15), np.float32)
            cams.append(cpu.lift_gather(img_feats[b * ncam + v], pix, pts, np.asarray(meta['lidar2img'][v]),
                                        score_w, score_b, input_w))
        out.append(np.concatenate(cams, 0))
    return out


def fetch_2d_voxels(img_feats, img_metas, score_w, score_b, spatial_shape, downscale_factor,
                    base_voxel_size, coors_range, max_points, max_voxels):
    """MSMDFusion.py:371-393 -> SpTensor((V,64) mean features with xyz / (13.5,13.5,2), (V,4))."""
    B = len(img_metas)
    fg = get_foreground2d(img_feats, img_metas, score_w, score_b)
    for i in range(B):
        if fg[i].shape[0] == 0:
            fg[i] = np.zeros((100, fg[i].shape[1]), np.float32)
    vs = [float(x) * downscale_factor for x in base_voxel_size]
    voxels, nums, coors = voxelize_batch(fg, vs, coors_range, max_points, max_voxels)
    feat = cpu.hard_simple_vfe(voxels, nums, voxels.shape[-1])
    feat[:, :3] = feat[:, :3] / np.array([13.5, 13.5, 2.0], np.float32)[None, :]
    return SpTensor(feat.astype(np.float32), coors, spatial_shape, B)


def _pad_missing_batch_id(indices, features, B):
    """sparse_multimodal_encoder_painting.py:208-225."""
    present = set(np.unique(indices[:, 0]).tolist())
    for b in range(B):
        if b not in present:
            pad_i = np.zeros((1, indices.shape[1]), indices.dtype)
            pad_i[0, 0] = b
            indices = np.concatenate([indices, pad_i], 0)
            features = np.concatenate([features, np.zeros((1, features.shape[1]), features.dtype)], 0)
    return indices, features


def grouped_sparse_conv(sd, prefix, v3, v2, syn3, syn2, stage_id, c3, fps_num, radius, nsample, thresh,
                        dummy_embedding, eps):
    """sparse_multimodal_encoder_painting.py:325-430.  v3 / v2 carry (N,5) indices (b,mix,z,y,x)."""
    B = v3.batch_size
    ind3, ind2, f3, f2 = v3.indices, v2.indices, v3.features, v2.features
    only3 = ind3[:, 1] == 0
    only2 = ind2[:, 1] == 0
    o2_idx, o2_feat = _pad_missing_batch_id(ind2[only2], f2[only2], B)
    o2_bzyx = o2_idx[:, [0, 2, 3, 4]]
    v3_bzyx = ind3[:, [0, 2, 3, 4]]
    nn = np.full((o2_bzyx.shape[0],), -1, np.int64)
    base = 0
    for b in np.unique(v3_bzyx[:, 0]).tolist():        # :355-369 (B = #batch ids present in voxel_3D)
        m2 = o2_bzyx[:, 0] == b
        m3 = v3_bzyx[:, 0] == b
        r = cpu.fps_nn_fast(o2_bzyx[m2], v3_bzyx[m3], fps_num, radius, nsample, thresh)
        r[r != -1] += base
        nn[m2] = r
        base = int(m3.sum())
    cross = _linear_relu(sd, f'{prefix}cross_gate_control.{stage_id}.0',
                         np.concatenate([f3, dummy_embedding.reshape(1, -1).astype(np.float32)], 0))
    o2_feat = cross[nn] * o2_feat                        # -1 -> last row = dummy embedding's gate
    x3 = SpTensor(f3[only3], np.ascontiguousarray(ind3[only3][:, [0, 2, 3, 4]]), v3.spatial_shape, B)
    m3f, m2f = f3[syn3], f2[syn2]
    gate = _linear_relu(sd, f'{prefix}gate_control.{stage_id}.0', m3f)
    mixed_feat = np.concatenate([m3f, gate * m2f], 1)
    mixed_idx, mixed_feat = _pad_missing_batch_id(ind2[syn2], mixed_feat, B)
    name = f'stage_{stage_id + 1}'
    x3 = convmodule(sd, f'{prefix}grouped_sp_conv_blocks_3D.{name}', x3, 'SubMConv3d', 3, 1, 1, eps)
    o2_feat = np.pad(o2_feat, ((0, 0), (c3, 0)))
    o3_feat = np.pad(x3.features, ((0, 0), (0, 64)))
    feat = np.concatenate([o3_feat, o2_feat, mixed_feat], 0).astype(np.float32)
    coors = np.concatenate([x3.indices, o2_bzyx, mixed_idx[:, [0, 2, 3, 4]]], 0).astype(np.int32)
    uni = SpTensor(feat, np.ascontiguousarray(coors), v2.spatial_shape, B)
    return basic_block(sd, f'{prefix}aggregation_blocks.{name}', uni, eps)


def multimodal_encoder(sd, cfg, v3_list, v2_list, syn3_list, syn2_list, fps_num_list, radius_list,
                       nsample_list, thresh_list, dummy_embeddings, prefix=''):
    """SparseMultiModalEncoderPaint.forward (:433-459).  dummy_embeddings[s] = the torch.rand(1,C3)
    draw of stage s (the caller draws them from the same seeded CPU generator)."""
    eps = cfg.get('norm_cfg', dict(eps=1e-3)).get('eps', 1e-3)
    c3s = cfg.get('in_channels_3D', (16, 32, 64, 128))
    pads = cfg.get('padding', (1, 1, 1, [0, 1, 1]))
    ksz = cfg.get('down_kernel_size', (3, 3, 3, [3, 1, 1]))
    strides = cfg.get('down_stride', (2, 2, 2, [2, 1, 1]))
    outs = []
    for s in range(len(v2_list)):
        x = grouped_sparse_conv(sd, prefix, v3_list[s], v2_list[s], syn3_list[s], syn2_list[s], s, c3s[s],
                                fps_num_list[s], radius_list[s], nsample_list[s], thresh_list[s],
                                dummy_embeddings[s], eps)
        if s > 0:
            oi, of = cpu.sparse_add(x.indices, x.features, outs[s - 1].indices, outs[s - 1].features,
                                    x.spatial_shape)
            x = SpTensor(of, oi, x.spatial_shape, x.batch_size)
        outs.append(convmodule(sd, f'{prefix}downscale_blocks.stage_{s + 1}', x, 'SparseConv3d', ksz[s],
                               strides[s], pads[s], eps))
    return outs


# --------------------------------------------------------------------------------------
# Detector level: MSMDFusionDetector.extract_pts_feat up to the tensor bev_fusion consumes
# --------------------------------------------------------------------------------------
def depth_aware_channel_compression(sd, feat_list, img_metas, prefix=''):
    """MSMDFusion.py:335-369.  The dense image-plane part (bilinear resample, Conv2d + BatchNorm2d(eval)
    + ReLU of `conv1x1_blocks`, :108-125) is not a kernel of this project: it is evaluated with plain
    torch fp32 ops on the CPU.  feat_list: three (B*6, 256, h, w) arrays -> three (B*6, 49, h, w)."""
    import torch
    import torch.nn.functional as F
    H, W = img_metas[0]['pad_shape'][:2]
    canvas = torch.from_numpy(depth_canvas(img_metas, H, W))
    out = []
    for i, feat in enumerate(feat_list):
        feat = torch.from_numpy(np.ascontiguousarray(feat, np.float32))
        depth = F.interpolate(canvas, feat.shape[-2:], mode='bilinear')
        w = torch.from_numpy(_np(sd, f'{prefix}conv1x1_blocks.{i}.0.weight').astype(np.float32))
        bn = [torch.from_numpy(_np(sd, f'{prefix}conv1x1_blocks.{i}.1.{k}').astype(np.float32))
              for k in ('running_mean', 'running_var', 'weight', 'bias')]
        y = F.conv2d(torch.cat([feat, depth], 1), w, padding=w.shape[-1] // 2)
        out.append(torch.relu(F.batch_norm(y, bn[0], bn[1], bn[2], bn[3], False, 0.0, 1e-3)).numpy())
    return out


def extract_voxel_space(sd, cfg, scenes, fpn_feats, img_metas, dummy_embeddings, compressed=None):
    """MSMDFusion.py:421-445 with extract_multiscale_voxel_feat (:400-419) inlined: voxelize ->
    HardSimpleVFE -> SparseEncoder -> (depth-aware compression -> 4x fetch_2D_voxels -> modality split) ->
    SparseMultiModalEncoderPaint -> dense -> cat.  `cfg`: the hot-path config (configs/msmd_lc_hotpath.py).
    `compressed`: the three compressed image features if the caller already has them.
    -> (bev (B, 256+384, 180, 180), [SpTensor] stage_outs, compressed)."""
    B = len(scenes)
    vl = cfg['pts_voxel_layer']
    max_voxels = vl['max_voxels'][1] if isinstance(vl['max_voxels'], (tuple, list)) else vl['max_voxels']
    ev, en, ec = voxelize_batch(scenes, vl['voxel_size'], vl['point_cloud_range'], vl['max_num_points'], max_voxels)
    mean = cpu.hard_simple_vfe(ev, en, cfg['pts_voxel_encoder']['num_features'])
    spatial, feats, _ = sparse_encoder(sd, dict(cfg['pts_middle_encoder']), mean, ec, B, prefix='pts_middle_encoder.')
    if compressed is None:
        compressed = depth_aware_channel_compression(sd, fpn_feats, img_metas)
    img_list = [compressed[0]] + list(compressed)                        # :404-405
    score_w = _np(sd, 'score_net.0.weight').reshape(-1)
    score_b = float(_np(sd, 'score_net.0.bias').reshape(-1)[0])
    v3l, v2l, s3l, s2l = [], [], [], []
    for i in range(4):
        v2 = fetch_2d_voxels(img_list[i], img_metas, score_w, score_b, cfg['spatial_shapes'][i],
                             cfg['downscale_factors'][i], vl['voxel_size'], vl['point_cloud_range'],
                             vl['max_num_points'], max_voxels)
        v3 = feats[i]
        c3, c2, s3, s2 = cpu.voxel_modality_split(v3.indices, v2.indices, B)
        v3l.append(SpTensor(v3.features, c3, v3.spatial_shape, B))
        v2l.append(SpTensor(v2.features, c2, v2.spatial_shape, B))
        s3l.append(s3)
        s2l.append(s2)
    outs = multimodal_encoder(sd, dict(cfg['multimodal_middle_encoder']), v3l, v2l, s3l, s2l, cfg['fps_num_list'],
                              cfg['radius_list'], cfg['max_cluster_samples_list'], cfg['dist_thresh_list'],
                              dummy_embeddings, prefix='multimodal_middle_encoder.')
    last = outs[-1]
    mm = cpu.dense(last.indices, last.features, last.spatial_shape, B)
    bev = np.concatenate([spatial, mm.reshape(B, -1, mm.shape[-2], mm.shape[-1])], 1)
    return bev, outs, compressed
