"""TEST INFRASTRUCTURE.  Stand-ins for the spconv-2.x / mmdet3d.ops names the reference's encoder classes
import, used when those classes are run in place (oracle/ref_lidar_encoder.py, oracle/ref_encoder.py).

The stand-ins RECORD the constructor arguments the reference passes and, when called, run the oracle's
conv restatement (oracle/model.py: convmodule / basic_block -- pinned separately against the reference's
vendored spconv-1.x) on the weights found under the module's own qualified name in a state dict.
"""
import types

import numpy as np

from . import cpu, model


def classes():
    import torch
    from torch import nn

    class Tensor:
        """The attributes / methods of spconv.SparseConvTensor the reference's encoders touch."""

        def __init__(self, features, indices, spatial_shape, batch_size):
            self.features, self.indices = features, indices
            self.spatial_shape, self.batch_size = list(spatial_shape), batch_size

        def replace_feature(self, features):
            return Tensor(features, self.indices, self.spatial_shape, self.batch_size)

        def dense(self):
            return torch.from_numpy(cpu.dense(np.ascontiguousarray(self.indices.numpy(), np.int32),
                                              np.ascontiguousarray(self.features.numpy(), np.float32),
                                              self.spatial_shape, self.batch_size))

    class Stub(nn.Module):
        sd, qualname = None, None

        def run(self, x):
            raise NotImplementedError

        def forward(self, x):
            y = self.run(to_oracle(x))
            return Tensor(torch.from_numpy(np.ascontiguousarray(y.features, np.float32)),
                          torch.from_numpy(np.ascontiguousarray(y.indices, np.int32)), y.spatial_shape, y.batch_size)

    class ConvModule(Stub):
        """make_sparse_convmodule (mmdet3d/ops/sparse_block.py:129-191), default order only."""

        def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0,
                     conv_type='SubMConv3d', norm_cfg=None, order=('conv', 'norm', 'act')):
            super().__init__()
            assert tuple(order) == ('conv', 'norm', 'act')
            self.spec = dict(kind='convmodule', conv_type=conv_type, cin=in_channels, cout=out_channels,
                             ksize=kernel_size, stride=stride, padding=padding, indice_key=indice_key,
                             eps=norm_cfg['eps'])

        def run(self, x):
            s = self.spec
            return model.convmodule(self.sd, self.qualname, x, s['conv_type'], s['ksize'], s['stride'], s['padding'],
                                    s['eps'])

    class BasicBlock(Stub):
        """SparseBasicBlock (mmdet3d/ops/sparse_block.py:68-126)."""

        def __init__(self, inplanes, planes, stride=1, downsample=None, conv_cfg=None, norm_cfg=None):
            super().__init__()
            assert stride == 1 and downsample is None and conv_cfg['type'] == 'SubMConv3d'
            self.spec = dict(kind='basicblock', cin=inplanes, cout=planes, eps=norm_cfg['eps'])

        def run(self, x):
            return model.basic_block(self.sd, self.qualname, x, self.spec['eps'])

    return types.SimpleNamespace(Tensor=Tensor, Stub=Stub, ConvModule=ConvModule, BasicBlock=BasicBlock)


def to_oracle(x):
    return model.SpTensor(np.ascontiguousarray(x.features.numpy(), np.float32),
                          np.ascontiguousarray(x.indices.numpy(), np.int32), x.spatial_shape, x.batch_size)


def namespace(st):
    """Globals for a reference encoder class body."""
    import torch
    import torch.nn.functional as F
    from torch import nn
    identity = lambda *a, **k: (lambda f: f)  # noqa: E731
    return {'nn': nn, 'torch': torch, 'F': F, 'auto_fp16': identity,
            'spconv': types.SimpleNamespace(SparseSequential=nn.Sequential, SparseConvTensor=st.Tensor),
            'make_sparse_convmodule': st.ConvModule, 'SparseBasicBlock': st.BasicBlock}


def bind(net, st, sd, prefix):
    """Give every stand-in its qualified name and the state dict; load the real torch sub-modules
    (the gate MLPs) from the same state dict."""
    import torch
    for name, mod in net.named_modules():
        if isinstance(mod, st.Stub):
            mod.sd, mod.qualname = sd, prefix + name
    if sd is not None:
        own = {k: torch.as_tensor(model._np(sd, prefix + k)).float() for k in net.state_dict()}
        net.load_state_dict(own)
    return net


def layer_table(net):
    return [(n, dict(m.spec)) for n, m in net.named_modules() if hasattr(m, 'spec')]
