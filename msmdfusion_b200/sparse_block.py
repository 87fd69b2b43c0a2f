"""SparseBasicBlock and make_sparse_convmodule -- mirrors ``mmdet3d/ops/sparse_block.py``.

``SparseBasicBlock`` inherits mmdet's ``BasicBlock`` in the reference (``:68-101``); mmdet is
not installed, so the sub-module names it would create (conv1, bn1, conv2, bn2, relu,
downsample; ``norm1``/``norm2`` accessors) are reproduced so checkpoints load.
"""
from torch import nn

from . import spconv
from .registry import build_conv_layer, build_norm_layer
from .spconv import _bn_foldable, _bn_scale_shift


class SparseBasicBlock(spconv.SparseModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, conv_cfg=None, norm_cfg=None):
        spconv.SparseModule.__init__(self)
        norm_cfg = norm_cfg if norm_cfg is not None else dict(type='BN')
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, 3, stride=stride, padding=1,
                                      dilation=1, bias=False)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, 3, padding=1, bias=False)
        self.add_module(self.norm2_name, norm2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    @property
    def norm1(self):
        return getattr(self, self.norm1_name)

    @property
    def norm2(self):
        return getattr(self, self.norm2_name)

    def forward(self, x):
        """mmdet3d/ops/sparse_block.py:103-126; eval-mode BN, residual and ReLU are fused into
        the two convolutions' epilogues (2 launches instead of 2 convs + 5 elementwise)."""
        identity = x.features
        assert x.features.dim() == 2, f'x.features.dim()={x.features.dim()}'
        fuse = _bn_foldable(self.norm1) and _bn_foldable(self.norm2) and self.downsample is None \
            and isinstance(self.conv1, spconv.SparseConvolution) and not self.conv1.conv1x1
        if fuse:
            s1, b1 = _bn_scale_shift(self.norm1)
            out = self.conv1.forward_fused(x, s1, b1, relu=True)
            s2, b2 = _bn_scale_shift(self.norm2)
            return self.conv2.forward_fused(out, s2, b2, relu=True, residual=identity)
        out = self.conv1(x)
        out = out.replace_feature(self.norm1(out.features))
        out = out.replace_feature(self.relu(out.features))
        out = self.conv2(out)
        out = out.replace_feature(self.norm2(out.features))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out.replace_feature(out.features + identity)
        out = out.replace_feature(self.relu(out.features))
        return out


def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key, stride=1, padding=0,
                           conv_type='SubMConv3d', norm_cfg=None, order=('conv', 'norm', 'act')):
    """mmdet3d/ops/sparse_block.py:129-191."""
    assert isinstance(order, tuple) and len(order) <= 3
    assert set(order) | {'conv', 'norm', 'act'} == {'conv', 'norm', 'act'}
    conv_cfg = dict(type=conv_type, indice_key=indice_key)
    layers = []
    for layer in order:
        if layer == 'conv':
            if conv_type not in ['SparseInverseConv3d', 'SparseInverseConv2d', 'SparseInverseConv1d']:
                layers.append(build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size,
                                               stride=stride, padding=padding, bias=False))
            else:
                layers.append(build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size,
                                               bias=False))
        elif layer == 'norm':
            layers.append(build_norm_layer(norm_cfg, out_channels)[1])
        elif layer == 'act':
            layers.append(nn.ReLU(inplace=True))
    return spconv.SparseSequential(*layers)
