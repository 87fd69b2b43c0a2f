"""Plan builder + Python face of the native sparse-network executor (csrc/executor.cu).

A *plan* is the flat layer list ``msmd_sparse_net_forward`` runs in one C-ABI call.  It is derived
from the module tree the reference builds (``SparseSequential(conv, BN1d, ReLU)`` of
``make_sparse_convmodule`` -- mmdet3d/ops/sparse_block.py:161-190 -- and ``SparseBasicBlock``
``:103-126``), so the modules, their parameters and their state-dict stay the single source of
truth; the plan only caches kernel-layout copies and is rebuilt when a parameter changes.
"""
import ctypes

import torch
from torch import nn

from . import ops, spconv
from ._cabi import ConvLayer, SparseDesc, check, lib, ptr, stream
from .sparse_block import SparseBasicBlock
from .spconv import SparseConvolution, SparseSequential, _bn_foldable, _bn_scale_shift


class Unsupported(Exception):
    """The module tree has something the fused executor does not take (training-mode BN, ...)."""


class SparseNetPlan:

    def __init__(self):
        self.layers = []     # dicts
        self.keep = []       # tensors the raw pointers in the C array point to
        self.version_key = None
        self.carray = None
        self.arena_bytes = 0

    # -- construction -------------------------------------------------------------------
    def _add_conv(self, conv, cur, bn=None, relu=False, residual=-1):
        if not isinstance(conv, SparseConvolution) or conv.conv1x1 or conv.transposed or conv.inverse:
            raise Unsupported(type(conv).__name__)
        scale = shift = None
        if bn is not None:
            if not _bn_foldable(bn, inference=True):
                raise Unsupported('BatchNorm1d is not in eval mode with running statistics')
            scale, shift = _bn_scale_shift(bn)
        if conv.bias is not None:
            b = conv.bias.detach().float()
            if scale is None:
                scale, shift = torch.ones_like(b), b.clone()
            else:
                shift = shift + b * scale
        w = conv.packed_weight()
        tcw = isinstance(w, ops.TcWeight)
        wt = w.packed if tcw else w
        self.keep += [wt, scale, shift]
        self.layers.append(dict(subm=int(conv.subm), ksize=conv.kernel_size, stride=conv.stride,
                                padding=conv.padding, dilation=conv.dilation, cin=conv.in_channels,
                                cout=conv.out_channels, weight=wt, weight_tc=(w.mode if tcw else 0), scale=scale,
                                shift=shift, relu=int(relu), input=cur, residual=residual))
        return len(self.layers)  # activation index of this layer's output

    def add(self, module, cur):
        """Append ``module`` (reading activation ``cur``); returns its output activation index."""
        if isinstance(module, SparseBasicBlock):
            if module.downsample is not None:
                raise Unsupported('SparseBasicBlock.downsample')
            mid = self._add_conv(module.conv1, cur, module.norm1, relu=True)
            return self._add_conv(module.conv2, mid, module.norm2, relu=True, residual=cur)
        if isinstance(module, SparseSequential):
            mods = list(module._modules.values())
            i = 0
            while i < len(mods):
                m = mods[i]
                if isinstance(m, SparseConvolution):
                    bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) else None
                    j = i + (2 if bn is not None else 1)
                    relu = j < len(mods) and isinstance(mods[j], nn.ReLU)
                    cur = self._add_conv(m, cur, bn, relu)
                    i = j + (1 if relu else 0)
                elif isinstance(m, (SparseSequential, SparseBasicBlock)):
                    cur = self.add(m, cur)
                    i += 1
                else:
                    raise Unsupported(type(m).__name__)
            return cur
        if isinstance(module, SparseConvolution):
            return self._add_conv(module, cur)
        raise Unsupported(type(module).__name__)

    def finalize(self):
        arr = (ConvLayer * len(self.layers))()
        for c, L in zip(arr, self.layers):
            c.subm = L['subm']
            for d in range(3):
                c.ksize[d], c.stride[d] = L['ksize'][d], L['stride'][d]
                c.padding[d], c.dilation[d] = L['padding'][d], L['dilation'][d]
            c.cin, c.cout = L['cin'], L['cout']
            c.weight = L['weight'].data_ptr()
            c.weight_tc = L['weight_tc']
            c.scale = L['scale'].data_ptr() if L['scale'] is not None else None
            c.shift = L['shift'].data_ptr() if L['shift'] is not None else None
            c.relu, c.input, c.residual = L['relu'], L['input'], L['residual']
        self.carray = arr
        return self

    # -- execution ----------------------------------------------------------------------
    def run(self, features, indices, spatial_shape, batch_size):
        """-> list of (features (n,C) f32, indices (n,4) i32, spatial_shape) per activation; the
        tensors are views into one arena allocation that they keep alive."""
        features = features.contiguous().float()
        indices = indices.contiguous()
        assert indices.dtype == torch.int32 and indices.shape[1] == 4
        dev = features.device
        n, c = features.shape
        nl = len(self.layers)
        acts = (SparseDesc * (nl + 1))()
        shape = (ctypes.c_int * 3)(*[int(s) for s in spatial_shape])
        nbytes = max(self.arena_bytes, (64 << 20) + 8192 * n)
        for _ in range(6):
            arena = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            with ops._Timed('sparse_net_forward', n=n, layers=nl):
                rc = lib().msmd_sparse_net_forward(self.carray, nl, ptr(features), ptr(indices), n, c,
                                                   int(batch_size), shape, ptr(arena), nbytes, acts,
                                                   stream(dev))
            if rc == -3:  # MSMD_ERR_WORKSPACE: grow the arena and run again (the plan is stateless)
                nbytes *= 2
                continue
            check(rc, 'msmd_sparse_net_forward')
            break
        else:
            raise RuntimeError('msmd_sparse_net_forward: arena keeps overflowing')
        self.arena_bytes = nbytes
        base = arena.data_ptr()
        out = [(features, indices, list(spatial_shape))]
        for a in acts[1:]:
            # a NULL descriptor pointer comes back from ctypes as None: an empty activation (n == 0)
            # has no arena bytes, and a SubM layer on an empty network input reuses the caller's
            # (NULL) index pointer
            if a.features is None or a.n == 0:
                f = features.new_zeros((int(a.n), int(a.channels)))
            else:
                fo = a.features - base
                f = arena[fo:fo + 4 * a.n * a.channels].view(torch.float32).view(a.n, a.channels)
            io = -1 if a.indices is None else a.indices - base
            if a.n == 0:
                idx = indices.new_zeros((0, 4))
            elif 0 <= io < nbytes:
                idx = arena[io:io + 16 * a.n].view(torch.int32).view(a.n, 4)
            else:  # SubM layers on the network input keep the caller's index tensor
                idx = indices
            out.append((f, idx, [int(s) for s in a.spatial_shape]))
        return out


class PlanWatch:
    """What a cached plan depends on, flattened ONCE: the parameter / buffer tensors of the module tree and its
    BatchNorm modules.  ``key()`` changes whenever one of them is replaced or modified in place, a BatchNorm changes
    mode, the conv path / precision switches or ``spconv.invalidate_caches()`` is called.  (Walking
    ``module.parameters()`` on every forward cost 1.4 ms of host time per LC scene -- profiles/r02f_lc_hostprofile.txt;
    the flat lists cost ~30 us.)  A parameter REPLACED by assignment is seen through ``_parameters`` identity."""

    def __init__(self, modules):
        self.owners, self.tensors, self.norms = [], [], []
        for m in modules:
            for sub in m.modules():
                for d in (sub._parameters, sub._buffers):
                    for name, t in d.items():
                        if t is not None:
                            self.owners.append((d, name))
                            self.tensors.append(t)
                if isinstance(sub, nn.modules.batchnorm._BatchNorm):
                    self.norms.append(sub)
        self.roots = list(modules)

    def key(self):
        ts = self.tensors
        for i, (d, name) in enumerate(self.owners):
            t = d.get(name)
            if t is not ts[i]:          # parameter object replaced (load with assign=True, manual assignment)
                return None
        return (spconv.CONV_PATH, spconv.CONV_PRECISION, spconv.cache_epoch(),
                tuple([t._version for t in ts]), tuple([t.data_ptr() for t in ts]),
                tuple([(b.training, b.track_running_stats) for b in self.norms]),
                tuple([m.training for m in self.roots]))


def plan_key(modules):
    """Changes whenever a parameter / buffer the plan baked in is replaced or modified in place."""
    return PlanWatch(modules).key()


def geometry_stream(device):
    """The executor's library-owned geometry stream of ``device`` as a torch stream (``msmd_executor_geometry_stream``)."""
    with torch.cuda.device(device):
        h = ctypes.c_void_p()
        check(lib().msmd_executor_geometry_stream(ctypes.byref(h)), 'msmd_executor_geometry_stream')
    return torch.cuda.ExternalStream(h.value, device=device)
