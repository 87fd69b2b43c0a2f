"""The spconv-2.x FUNCTIONAL boundary of the path ("Under spconv-2.x"): the two functions the
reference's patched convolution module calls --

    ops.get_indice_pairs_implicit_gemm(indices, batch_size, spatial_shape, algo, ksize, stride, padding,
                                       dilation, out_padding, subm, transpose, is_train, alloc, timer)
        -> (outids, num_inds_per_loc, pair_fwd, pair_bwd, pair_mask_fwd_splits, pair_mask_bwd_splits,
            mask_argsort_fwd_splits, mask_argsort_bwd_splits, masks)              bug_fix/conv.py:382-415
    Fsp.implicit_gemm(features, filters, pair_fwd, pair_bwd, pair_mask_fwd_splits, pair_mask_bwd_splits,
                      mask_argsort_fwd_splits, mask_argsort_bwd_splits, num_activate_out, masks, is_train,
                      is_subm, timer, fp32_accum)  -> out_features                bug_fix/conv.py:442-447

-- with the same positional / keyword signatures and return structure, implemented on this project's C ABI.
``msmdfusion_b200.spconv.SubMConv3d / SparseConv3d`` do NOT go through here (they fuse BatchNorm / ReLU / residual
and cache rulebooks per index set); this module is the lower drop-in point: a host project that keeps spconv's own
Python modules (the reference ships ``bug_fix/conv.py`` to be copied over spconv's ``conv.py``) swaps these two
functions and nothing else.  The CPU suite runs the reference's ``SparseConvolution.forward`` from its source in place
on top of them.

Conventions: ``pair_fwd[k, o]`` = input row or -1, kernel offset k row-major (kz, ky, kx);
SubM ``outids`` is ``indices`` itself; strided ``outids`` ascending by linear index; ``pair_bwd[k, i]`` = output row or
-1 (strided / training only -- empty otherwise, as spconv returns it); one mask split (kernel volume <= 32):
``pair_mask_*_splits[0][r]`` = bit k set iff ``pair[k, r] >= 0``; ``mask_argsort_*_splits[0]`` = the row order of the
mask-sorted tiles (15-bit digest order of ``msmd_rulebook_mask_sort`` for 3x3x3 kernels, ascending mask otherwise);
``masks[0]`` = all-ones split mask; ``num_inds_per_loc[k]`` = number of pairs of offset k.
"""
import enum
import math
import types
import weakref

import numpy as np
import torch

from . import autograd as _ag
from . import ops as _ops


class ConvAlgo(enum.Enum):
    """spconv.core.ConvAlgo (bug_fix/conv.py:28,90-99)."""
    Native = 0
    MaskImplicitGemm = 1
    MaskSplitImplicitGemm = 2


class ImplicitGemmIndiceData:
    """spconv.pytorch.core.ImplicitGemmIndiceData as bug_fix/conv.py:415-432 constructs and reads it."""

    def __init__(self, out_indices, indices, pair_fwd, pair_bwd, pair_mask_fwd_splits, pair_mask_bwd_splits,
                 mask_argsort_fwd_splits, mask_argsort_bwd_splits, masks, is_subm, spatial_shape, out_spatial_shape,
                 algo, ksize, stride, padding, dilation):
        self.out_indices, self.indices = out_indices, indices
        self.pair_fwd, self.pair_bwd = pair_fwd, pair_bwd
        self.pair_mask_fwd_splits, self.pair_mask_bwd_splits = pair_mask_fwd_splits, pair_mask_bwd_splits
        self.mask_argsort_fwd_splits, self.mask_argsort_bwd_splits = mask_argsort_fwd_splits, mask_argsort_bwd_splits
        self.masks = masks
        self.is_subm = is_subm
        self.spatial_shape, self.out_spatial_shape = spatial_shape, out_spatial_shape
        self.algo = algo
        self.ksize, self.stride, self.padding, self.dilation = ksize, stride, padding, dilation


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    """spconv.pytorch.ops.get_conv_output_size (vendored spconv-1.x equivalent mmdet3d/ops/spconv/ops.py:20-31)."""
    out = []
    for i, k, s, p, d in zip(input_size, kernel_size, stride, padding, dilation):
        size = (i + 2 * p - d * (k - 1) - 1) // s + 1
        out.append(size if k != -1 else 1)
    return out


def _pair_mask(pair):
    """(K, n) table -> (n,) int32 mask, bit k set iff pair[k, r] >= 0 (K <= 32)."""
    k = pair.shape[0]
    bits = (torch.ones(k, dtype=torch.int64, device=pair.device) << torch.arange(k, device=pair.device))
    m = ((pair >= 0).to(torch.int64) * bits[:, None]).sum(0)
    return torch.where(m >= 2 ** 31, m - 2 ** 32, m).to(torch.int32)   # bit 31 = the sign bit of the int32 word


def _mask_argsort(pair, mask):
    if pair.shape[0] == 27 and pair.shape[1] > 0:
        return _ops.rulebook_mask_sort(pair)[0]
    return torch.argsort(mask.to(torch.int64) & 0xFFFFFFFF, stable=True).to(torch.int32)


def get_indice_pairs_implicit_gemm(indices, batch_size, spatial_shape, algo, ksize, stride, padding, dilation,
                                   out_padding, subm=False, transpose=False, is_train=True, alloc=None, timer=None):
    if transpose:
        raise NotImplementedError('transposed sparse convolution is not on the MSMDFusion path')
    ksize, stride, padding, dilation = (list(_ops._triple(v)) for v in (ksize, stride, padding, dilation))
    kvol = int(math.prod(ksize))
    assert kvol <= 32, "implicit gemm don't support kv >= 32 for now"    # bug_fix/conv.py:99
    idx = indices if indices.dtype == torch.int32 else indices.int()
    idx = idx.contiguous()
    grid = _ops.grid_build(idx, batch_size, spatial_shape)
    dev = idx.device
    empty = torch.empty((0,), dtype=torch.int32, device=dev)
    if subm:
        pair_fwd = _ops.rulebook_subm(idx, grid, ksize, dilation)
        outids, pair_bwd = indices, empty
    else:
        outids, pair_fwd, _ = _ops.rulebook_conv(idx, grid, ksize, stride, padding, dilation)
        pair_bwd = _ops.rulebook_transpose(pair_fwd, idx.shape[0]) if is_train else empty
    mask_fwd = _pair_mask(pair_fwd)
    fwd_masks, fwd_sorts = [mask_fwd], [_mask_argsort(pair_fwd, mask_fwd)]
    if pair_bwd.numel():
        mask_bwd = _pair_mask(pair_bwd)
        bwd_masks, bwd_sorts = [mask_bwd], [_mask_argsort(pair_bwd, mask_bwd)]
    else:
        bwd_masks, bwd_sorts = [], []
    num_inds_per_loc = (pair_fwd >= 0).sum(1).to(torch.int32)
    masks = [np.array([0xFFFFFFFF], dtype=np.uint32)]
    return (outids, num_inds_per_loc, pair_fwd, pair_bwd, fwd_masks, bwd_masks, fwd_sorts, bwd_sorts, masks)


# id(weight) -> (weakref to the tensor, key, packed): the kernel-layout copy follows the parameter.  The weak
# reference is compared with `is`, so a new tensor that happens to reuse a dead one's id / address / version
# cannot hit the old entry.
_PACKED = {}


def _packed_of(weight):
    from . import spconv as _sp
    kvol = int(math.prod(weight.shape[1:-1]))
    use_tc = _sp.CONV_PATH == 'tc' and _ops.tc_supported(weight.shape[0], kvol, weight.shape[-1])
    mode = _ops.TC_MODES[_sp.CONV_PRECISION] if use_tc else 0
    key = (_sp.cache_epoch(), weight._version, weight.data_ptr(), mode)
    ent = _PACKED.get(id(weight))
    if ent is None or ent[0]() is not weight or ent[1] != key:
        packed = _ops.pack_weight_tc(weight, mode) if use_tc else _ops.pack_weight(weight)
        _PACKED[id(weight)] = ent = (weakref.ref(weight), key, packed)
        if len(_PACKED) > 512:   # parameters that went away
            for k in [k for k, v in _PACKED.items() if v[0]() is None]:
                _PACKED.pop(k, None)
    return ent[2]


def implicit_gemm(features, filters, pair_fwd, pair_bwd, pair_mask_fwd_splits, pair_mask_bwd_splits,
                  mask_argsort_fwd_splits, mask_argsort_bwd_splits, num_activate_out, masks, is_train=False,
                  is_subm=False, timer=None, fp32_accum=None):
    """out[o] = sum_k W[:, k, :] . features[pair_fwd[k, o]]; ``filters`` is the KRSC parameter
    [Cout, kz, ky, kx, Cin] (bug_fix/conv.py:114-117).  Differentiable w.r.t. features and filters (the backward
    of the reference's autograd function: csrc/spconv_bwd.cu)."""
    assert pair_fwd.shape[1] == num_activate_out
    packed = _packed_of(filters)
    if torch.is_grad_enabled() and (features.requires_grad or filters.requires_grad):
        from . import spconv as _sp
        rb = dict(pair_fwd=pair_fwd, subm=bool(is_subm), path=_sp.CONV_PATH, unique=False)
        if not is_subm and pair_bwd is not None and pair_bwd.numel():
            rb['pair_bwd'] = pair_bwd
        return _ag.SparseConvFunction.apply(features, filters, packed, rb)
    return _ops.spconv_fwd(features, packed, pair_fwd)


def sparse_add(a, b):
    from . import functional as _fn
    return _fn.sparse_add(a, b)


# the two namespaces bug_fix/conv.py imports as `ops` and `Fsp` (:31-32)
ops = types.SimpleNamespace(get_indice_pairs_implicit_gemm=get_indice_pairs_implicit_gemm,
                            get_conv_output_size=get_conv_output_size)
Fsp = types.SimpleNamespace(implicit_gemm=implicit_gemm, sparse_add=sparse_add)
