"""Autograd functions of the voxel-space path (config 5 of BASELINE.json: the train step).

The reference differentiates its sparse convolutions through spconv-2.x's ``Fsp.implicit_gemm``
(backward over ``pair_bwd`` / ``mask_argsort_bwd_splits``, ``bug_fix/conv.py:382-415,442-447``),
``Fsp.sparse_add`` and ``SparseConvTensor.dense()``; everything else on the path is plain torch
(BatchNorm1d on ``.features``, the gate ``nn.Linear``s, indexing, concatenation).  Gradient scope as in
the reference: ``voxelize`` runs under ``no_grad`` and the lift has no backward, so the
virtual-point features are constants; the LiDAR ``SparseEncoder`` is frozen; gradients reach the 16
convolutions of ``SparseMultiModalEncoderPaint``, its BatchNorms and its gate MLPs.

All compute goes through the C ABI (``ops.py`` -> ``csrc/spconv_bwd.cu``); there is no CPU fallback.
"""
import math

import torch

from . import ops


class SparseConvFunction(torch.autograd.Function):
    """out = conv(features; weight KRSC, pair_fwd).  ``rb`` carries the rulebook: ``pair_fwd``
    (K,N_out), ``subm`` and a slot for the lazily built ``pair_bwd`` (strided convs only)."""

    @staticmethod
    def forward(ctx, features, weight, packed, rb):
        out = ops.spconv_fwd(features, packed, rb['pair_fwd'])
        ctx.save_for_backward(features, weight)
        ctx.rb = rb
        ctx.tc_mode = packed.mode if isinstance(packed, ops.TcWeight) else 0   # dgrad runs in the forward's precision
        return out

    @staticmethod
    def backward(ctx, grad_out):
        features, weight = ctx.saved_tensors
        rb = ctx.rb
        grad_out = grad_out.contiguous()
        grad_in = grad_w = None
        cout, cin = weight.shape[0], weight.shape[-1]
        kvol = int(math.prod(weight.shape[1:-1]))
        if ctx.needs_input_grad[0]:
            use_tc = ctx.tc_mode != 0 and rb.get('path', 'tc') == 'tc' and ops.tc_supported(cin, kvol, cout)
            mode = ctx.tc_mode

            def pack(w):
                return ops.pack_weight_tc(w, mode) if use_tc else ops.pack_weight(w)
            if rb['subm']:
                # o reads i through offset k  <=>  i reads o through offset K-1-k: pair_fwd itself is
                # the transposed rulebook once the weight's offsets are reversed
                pair = rb['pair_fwd']
                packed_t = pack(ops.transpose_weight(weight, flip_k=True))
                if rb.get('unique'):
                    grad_in = ops.spconv_bwd_data(grad_out, packed_t, pair)
                else:
                    # The index set may hold DUPLICATE coordinates (the unified voxel list of the GMA
                    # conv does whenever the float32 keys of voxel_modality_split collide, z >= 17).
                    # The forward lets the largest row of a coordinate be the one neighbours read
                    # (its "owner" = the centre-offset entry of the rulebook); all rows of a
                    # coordinate have the same rulebook column.  So: fold the output gradients of a
                    # coordinate onto its owner, run the mirrored contraction, and give the rows
                    # nobody reads a zero gradient.  With unique coordinates this is the identity.
                    rows = torch.arange(pair.shape[1], device=pair.device)
                    owner = pair[kvol // 2].long()
                    owner = torch.where(owner < 0, rows, owner)
                    folded = torch.zeros_like(grad_out).index_add_(0, owner, grad_out)
                    grad_in = ops.spconv_bwd_data(folded, packed_t, pair)
                    grad_in = grad_in * (owner == rows).to(grad_in.dtype)[:, None]
            else:
                # strided conv: output coordinates are unique, so (k, i) has at most one reader o
                pair_bwd = rb.get('pair_bwd')
                if pair_bwd is None:
                    pair_bwd = rb['pair_bwd'] = ops.rulebook_transpose(rb['pair_fwd'], features.shape[0])
                grad_in = ops.spconv_bwd_data(grad_out, pack(ops.transpose_weight(weight)), pair_bwd)
        if ctx.needs_input_grad[1]:
            grad_w = ops.spconv_bwd_weight(features, grad_out, rb['pair_fwd'], weight.shape)
            if grad_w.dtype != weight.dtype:
                grad_w = grad_w.to(weight.dtype)
        return grad_in, grad_w, None, None


class ToDenseFunction(torch.autograd.Function):
    """``SparseConvTensor.dense()``: backward gathers the active rows out of the dense gradient."""

    @staticmethod
    def forward(ctx, features, indices, spatial_shape, batch_size):
        ctx.indices, ctx.shape, ctx.batch = indices, list(spatial_shape), int(batch_size)
        return ops.to_dense(indices, features, spatial_shape, batch_size)

    @staticmethod
    def backward(ctx, grad_dense):
        g = ops.from_dense(ctx.indices, grad_dense, ctx.shape, ctx.batch)
        return g, None, None, None


class SparseAddFunction(torch.autograd.Function):
    """``Fsp.sparse_add``: out[r] = a[rows_a -> r] + b[rows_b -> r]; backward is two row gathers
    through the union grid (``ops.grid_rows``)."""

    @staticmethod
    def forward(ctx, feat_a, feat_b, idx_a, idx_b, spatial_shape, batch_size, holder):
        out_idx, out_feat, grid = ops.sparse_add(idx_a, feat_a, idx_b, feat_b, spatial_shape, batch_size)
        holder['out_idx'], holder['grid'] = out_idx, grid
        ctx.idx_a, ctx.idx_b, ctx.grid = idx_a, idx_b, grid
        return out_feat

    @staticmethod
    def backward(ctx, grad_out):
        grad_out = grad_out.contiguous()
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = grad_out.index_select(0, ops.grid_rows(ctx.idx_a, ctx.grid))
        if ctx.needs_input_grad[1]:
            gb = grad_out.index_select(0, ops.grid_rows(ctx.idx_b, ctx.grid))
        return ga, gb, None, None, None, None, None
