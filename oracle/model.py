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
