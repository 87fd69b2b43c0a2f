"""Test-only stand-ins for ``msmdfusion_b200.ops`` built on the CPU oracle.

They let the HOST logic of the train step (autograd graph, SubM offset reversal, rulebook
transposition, module plumbing) be checked in the CPU suite against torch-native autograd.  Each
stand-in restates the CONTRACT of the C-ABI entry it replaces (include/msmd_b200.h) -- e.g.
``spconv_bwd_data`` really is "the forward contraction over pair_bwd with the packed transposed
weight" -- so a host-side mistake (wrong flip, wrong table, wrong layout) shows up here.  The CUDA
kernels themselves are checked by the ``-m gpu`` tests.  Never imported by the product.
"""
import numpy as np
import torch

from msmdfusion_b200 import ops
from oracle import cpu


class CpuGrid:
    def __init__(self, indices, spatial_shape, batch_size):
        self.indices = indices
        self.spatial_shape, self.batch_size = [int(s) for s in spatial_shape], int(batch_size)
        self.bits = self.prefix = self.perm = None


def _np(t):
    return t.detach().cpu().numpy()


def _krsc_of_packed(packed):
    """[K, Cin, Cout] (``msmd_spconv_pack_weight``) -> [Cout, K, 1, 1, Cin]."""
    k, cin, cout = packed.shape
    return np.ascontiguousarray(_np(packed).transpose(2, 0, 1)).reshape(cout, k, 1, 1, cin)


def grid_build(indices, batch_size, spatial_shape, need_perm=True):
    return CpuGrid(indices, spatial_shape, batch_size)


def rulebook_subm(indices, grid, ksize, dilation=1):
    return torch.from_numpy(cpu.subm_rulebook(_np(indices), grid.spatial_shape, list(ksize), list(ops._triple(dilation))))


def rulebook_conv(indices, grid, ksize, stride, padding, dilation=1):
    oi, pair, oshape = cpu.conv_rulebook(_np(indices), grid.spatial_shape, list(ksize), list(stride),
                                         list(padding), list(ops._triple(dilation)))
    oi = torch.from_numpy(oi)
    return oi, torch.from_numpy(pair), CpuGrid(oi, oshape, grid.batch_size)


def pack_weight(weight):
    w = weight.detach().float()
    cout, cin = w.shape[0], w.shape[-1]
    return w.reshape(cout, -1, cin).permute(1, 2, 0).contiguous()


def spconv_fwd(features, packed, pair_fwd, scale=None, shift=None, residual=None, relu=False):
    out = torch.from_numpy(cpu.spconv_fwd(_np(features), _krsc_of_packed(packed), _np(pair_fwd)))
    if scale is not None:
        out = out * scale + shift
    if residual is not None:
        out = out + residual
    return torch.relu(out) if relu else out


def rulebook_transpose(pair_fwd, n_in):
    return torch.from_numpy(cpu.pair_transpose(_np(pair_fwd), int(n_in)))


def transpose_weight(weight, flip_k=False):
    """msmd_spconv_transpose_weight: Wt[ci, k', co] = W[co, k, ci], k' = K-1-k when flip_k."""
    w = weight.detach().float()
    cout, cin = w.shape[0], w.shape[-1]
    w3 = w.reshape(cout, -1, cin)
    if flip_k:
        w3 = w3.flip(1)
    return w3.permute(2, 1, 0).contiguous().reshape(cin, *w.shape[1:-1], cout)


def spconv_bwd_data(grad_out, packed_wt, pair_bwd):
    """msmd_spconv_bwd_data: the forward kernels on grad_out with the packed transposed weight."""
    return spconv_fwd(grad_out, packed_wt, pair_bwd)


def spconv_bwd_weight(features, grad_out, pair_fwd, weight_shape):
    cout, cin = int(weight_shape[0]), int(weight_shape[-1])
    k = pair_fwd.shape[0]
    w0 = np.zeros((cout, k, 1, 1, cin), np.float32)
    _, gw = cpu.spconv_bwd(_np(features), w0, _np(pair_fwd), _np(grad_out), need_input_grad=False)
    return torch.from_numpy(gw).reshape(tuple(weight_shape))


def to_dense(indices, features, spatial_shape, batch_size):
    return torch.from_numpy(cpu.dense(_np(indices), _np(features).astype(np.float32), list(spatial_shape),
                                      int(batch_size)))


def from_dense(indices, dense, spatial_shape, batch_size):
    i = indices.long()
    return dense[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]].contiguous()


def sparse_add(idx_a, feat_a, idx_b, feat_b, spatial_shape, batch_size):
    oi, of = cpu.sparse_add(_np(idx_a), _np(feat_a), _np(idx_b), _np(feat_b), list(spatial_shape))
    oi = torch.from_numpy(oi)
    return oi, torch.from_numpy(of), CpuGrid(oi, spatial_shape, batch_size)


def grid_rows(indices, grid):
    D, H, W = grid.spatial_shape

    def lin(t):
        t = t.long()
        return ((t[:, 0] * D + t[:, 1]) * H + t[:, 2]) * W + t[:, 3]
    keys = lin(grid.indices)            # ascending
    q = lin(indices)
    pos = torch.searchsorted(keys, q).clamp_max(max(keys.numel() - 1, 0))
    return torch.where(keys[pos] == q, pos, torch.full_like(pos, -1))


STANDINS = dict(grid_build=grid_build, rulebook_subm=rulebook_subm, rulebook_conv=rulebook_conv,
                pack_weight=pack_weight, spconv_fwd=spconv_fwd, rulebook_transpose=rulebook_transpose,
                transpose_weight=transpose_weight, spconv_bwd_data=spconv_bwd_data,
                spconv_bwd_weight=spconv_bwd_weight, to_dense=to_dense, from_dense=from_dense,
                sparse_add=sparse_add, grid_rows=grid_rows, tc_supported=lambda *a: False)


def install(monkeypatch):
    for name, fn in STANDINS.items():
        monkeypatch.setattr(ops, name, fn)
