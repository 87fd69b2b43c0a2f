"""``spconv.pytorch.functional`` surface used by the hot path (imported as ``Fsp`` in
``mmdet3d/models/middle_encoders/sparse_multimodal_encoder_painting.py:12``)."""
import torch

from . import autograd as _ag
from . import ops
from . import spconv as _sp


def sparse_add(a, b):
    """``Fsp.sparse_add(a, b)`` (call site sparse_multimodal_encoder_painting.py:455).

    Union of the two active-voxel sets, features of coincident voxels summed, output rows in
    ascending (batch, z, y, x) order -- what spconv v2.1.21's ``torch.sparse`` coalesce
    produces.  The union's occupancy bit grid is kept on the result, so the strided
    convolution that follows does not rebuild it.
    """
    assert isinstance(a, _sp.SparseConvTensor) and isinstance(b, _sp.SparseConvTensor)
    assert list(a.spatial_shape) == list(b.spatial_shape), 'sparse_add: spatial shapes differ'
    assert a.batch_size == b.batch_size, 'sparse_add: batch sizes differ'
    assert a.features.shape[1] == b.features.shape[1], 'sparse_add: channel sizes differ'
    ia = a.indices if a.indices.dtype == torch.int32 else a.indices.int()
    ib = b.indices if b.indices.dtype == torch.int32 else b.indices.int()
    if torch.is_grad_enabled() and (a.features.requires_grad or b.features.requires_grad):
        holder = {}   # train step: Fsp.sparse_add is differentiable w.r.t. both feature tensors
        out_feat = _ag.SparseAddFunction.apply(a.features, b.features, ia.contiguous(), ib.contiguous(),
                                               a.spatial_shape, a.batch_size, holder)
        out_idx, grid = holder['out_idx'], holder['grid']
    else:
        out_idx, out_feat, grid = ops.sparse_add(ia, a.features, ib, b.features, a.spatial_shape,
                                                 a.batch_size)
    out = _sp.SparseConvTensor(out_feat, out_idx, a.spatial_shape, a.batch_size,
                               benchmark=a.benchmark)
    _sp._attach_iset(out, _sp.IndexSet(out_idx, a.spatial_shape, a.batch_size, grid=grid, unique=True))
    return out
