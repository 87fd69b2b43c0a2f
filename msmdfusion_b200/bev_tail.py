"""BEV tail behind the voxel-space path: the ``SECOND`` backbone and ``SECONDFPN`` neck
that ``configs/MSMDFusion_nusc_voxel_LC.py:191-206`` / ``transfusion_nusc_voxel_L.py`` put between the BEV
tensor (after ``bev_fusion`` = SPPModule) and ``TransFusionHead``.

Interface mirrored: ``mmdet3d/models/backbones/second.py:9-86`` (``SECOND(in_channels, out_channels, layer_nums,
layer_strides, norm_cfg, conv_cfg)``, ``forward(x) -> tuple``) and ``mmdet3d/models/necks/second_fpn.py:12-92``
(``SECONDFPN(in_channels, out_channels, upsample_strides, norm_cfg, upsample_cfg, conv_cfg,
use_conv_for_no_stride)``, ``forward(xs) -> [tensor]``), registered under the same ``BACKBONES`` / ``NECKS`` names,
with the same state-dict keys and shapes (``blocks.<stage>.<3j>.weight`` / ``blocks.<stage>.<3j+1>.*``,
``deblocks.<i>.0.weight`` / ``deblocks.<i>.1.*``) so reference checkpoints load.

These are DENSE 2-D convolutions: the kernels are cuDNN's (library code), there is no hand-written kernel of
this project here and no claim on it.  What is B200-first is the inference data path (``fused`` below): BatchNorm
(eval, under ``no_grad``) is folded into the convolution weights once per parameter version, activations stay channels-last from
the first layer to the concatenation, the ReLU runs in place on the conv output, and ``dtype='bf16'``
(``MSMD_BEV_DTYPE=bf16``) runs the stack under bf16 with fp32 folded weights cast once.  In training mode, or with
``fused=False``, the modules run layer by layer exactly like the reference (bit-identical on the same torch build:
the CPU test test_second_and_secondfpn_match_reference_classes_live).
"""
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .registry import BACKBONES, NECKS, build_conv_layer, build_norm_layer, build_upsample_layer

BEV_DTYPE = os.environ.get('MSMD_BEV_DTYPE', 'fp32')   # 'fp32' | 'bf16' (fused inference path only)


def _fold(conv, bn):
    """(weight, bias) of conv followed by BatchNorm in eval mode, as one convolution."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    w = conv.weight
    if isinstance(conv, nn.ConvTranspose2d):     # (in, out, kh, kw): output channels on dim 1
        w = w * scale.view(1, -1, 1, 1)
    else:
        w = w * scale.view(-1, 1, 1, 1)
    b = shift if conv.bias is None else shift + conv.bias * scale
    return w, b


class _FoldedStack:
    """Cached BN-folded copy of a ``Sequential(conv, bn, relu, conv, bn, relu, ...)``."""

    def __init__(self, seq):
        self.seq = seq
        self.key = None
        self.layers = []

    def _current_key(self, dtype):
        key = [dtype]
        for t in list(self.seq.parameters()) + list(self.seq.buffers()):
            key.append((t.data_ptr(), t._version))
        return tuple(key)

    def layers_for(self, dtype):
        key = self._current_key(dtype)
        if key != self.key:
            mods = list(self.seq)
            assert len(mods) % 3 == 0
            self.layers = []
            with torch.no_grad():
                for j in range(0, len(mods), 3):
                    conv, bn, act = mods[j:j + 3]
                    assert isinstance(bn, nn.BatchNorm2d) and isinstance(act, nn.ReLU)
                    w, b = _fold(conv, bn)
                    w = w.to(dtype).contiguous(memory_format=torch.channels_last)
                    self.layers.append((conv, w, b.to(dtype)))
            self.key = key
        return self.layers

    def __call__(self, x, dtype):
        for conv, w, b in self.layers_for(dtype):
            if isinstance(conv, nn.ConvTranspose2d):
                x = F.conv_transpose2d(x, w, b, conv.stride, conv.padding, conv.output_padding, conv.groups,
                                       conv.dilation)
            else:
                x = F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
            x = F.relu_(x)
        return x


def _foldable(seq):
    mods = list(seq)
    return len(mods) % 3 == 0 and all(
        isinstance(mods[j], (nn.Conv2d, nn.ConvTranspose2d)) and isinstance(mods[j + 1], nn.BatchNorm2d) and
        mods[j + 1].track_running_stats and isinstance(mods[j + 2], nn.ReLU) for j in range(0, len(mods), 3))


def _run_dtype(dtype):
    name = BEV_DTYPE if dtype is None else dtype
    return {'fp32': torch.float32, 'bf16': torch.bfloat16}[name]


@BACKBONES.register_module()
class SECOND(nn.Module):
    """second.py:9-86.  Stage i: conv3x3(stride = layer_strides[i]) + ``layer_nums[i]`` x conv3x3, each followed by
    BatchNorm + ReLU."""

    def __init__(self, in_channels=128, out_channels=(128, 128, 256), layer_nums=(3, 5, 5), layer_strides=(2, 2, 2),
                 norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False),
                 fused=True, dtype=None):
        super().__init__()
        out_channels, layer_nums, layer_strides = list(out_channels), list(layer_nums), list(layer_strides)
        assert len(layer_strides) == len(layer_nums) == len(out_channels)
        stages = []
        for i, width in enumerate(out_channels):
            cin = in_channels if i == 0 else out_channels[i - 1]
            plan = [(cin, layer_strides[i])] + [(width, 1)] * layer_nums[i]
            mods = []
            for c, stride in plan:
                mods += [build_conv_layer(conv_cfg, c, width, 3, stride=stride, padding=1),
                         build_norm_layer(norm_cfg, width)[1], nn.ReLU(inplace=True)]
            stages.append(nn.Sequential(*mods))
        self.blocks = nn.ModuleList(stages)
        self.fused, self.dtype = fused, dtype
        self._folded = None

    def init_weights(self, pretrained=None):
        if isinstance(pretrained, str):   # second.py:66-71 loads a checkpoint non-strictly
            self.load_state_dict(torch.load(pretrained, map_location='cpu').get('state_dict', {}), strict=False)

    def forward(self, x):
        if self.fused and not self.training and not torch.is_grad_enabled() and all(_foldable(b) for b in self.blocks):
            if self._folded is None:
                self._folded = [_FoldedStack(b) for b in self.blocks]
            dt = _run_dtype(self.dtype)
            x = x.to(dt).contiguous(memory_format=torch.channels_last)
            outs = []
            for stack in self._folded:
                x = stack(x, dt)
                outs.append(x)
            return tuple(outs)
        outs = []
        for block in self.blocks:
            x = block(x)
            outs.append(x)
        return tuple(outs)


@NECKS.register_module()
class SECONDFPN(nn.Module):
    """second_fpn.py:12-92.  Level i: transposed conv (kernel = stride = upsample_strides[i]) -- or a plain conv
    with kernel = stride = round(1 / stride) when the stride is 1 and ``use_conv_for_no_stride`` (or below 1) --
    + BatchNorm + ReLU; the levels are concatenated along the channels."""

    def __init__(self, in_channels=(128, 128, 256), out_channels=(256, 256, 256), upsample_strides=(1, 2, 4),
                 norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), upsample_cfg=dict(type='deconv', bias=False),
                 conv_cfg=dict(type='Conv2d', bias=False), use_conv_for_no_stride=False, fused=True, dtype=None):
        super().__init__()
        assert len(out_channels) == len(upsample_strides) == len(in_channels)
        self.in_channels, self.out_channels = list(in_channels), list(out_channels)
        self.fp16_enabled = False
        levels = []
        for cin, cout, stride in zip(in_channels, out_channels, upsample_strides):
            if stride > 1 or (stride == 1 and not use_conv_for_no_stride):
                up = build_upsample_layer(upsample_cfg, in_channels=cin, out_channels=cout, kernel_size=stride,
                                          stride=stride)
            else:
                k = int(np.round(1 / stride))
                up = build_conv_layer(conv_cfg, in_channels=cin, out_channels=cout, kernel_size=k, stride=k)
            levels.append(nn.Sequential(up, build_norm_layer(norm_cfg, cout)[1], nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(levels)
        self.fused, self.dtype = fused, dtype
        self._folded = None

    def init_weights(self):
        for mod in self.modules():   # second_fpn.py:66-72: kaiming_init convs, constant_init norms
            if isinstance(mod, nn.Conv2d):
                nn.init.kaiming_normal_(mod.weight, a=0, mode='fan_out', nonlinearity='relu')
                if mod.bias is not None:
                    nn.init.constant_(mod.bias, 0)
            elif isinstance(mod, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(mod.weight, 1)
                nn.init.constant_(mod.bias, 0)

    def forward(self, x):
        assert len(x) == len(self.in_channels)
        if self.fused and not self.training and not torch.is_grad_enabled() and \
                all(_foldable(b) for b in self.deblocks):
            if self._folded is None:
                self._folded = [_FoldedStack(b) for b in self.deblocks]
            dt = _run_dtype(self.dtype)
            ups = [stack(xi.to(dt).contiguous(memory_format=torch.channels_last), dt)
                   for xi, stack in zip(x, self._folded)]
        else:
            ups = [block(xi) for xi, block in zip(x, self.deblocks)]
        return [torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]]
