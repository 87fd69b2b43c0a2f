"""TEST INFRASTRUCTURE.  Runs the reference's own `SparseMultiModalEncoderPaint` class where it lies.

The class body (mmdet3d/models/middle_encoders/sparse_multimodal_encoder_painting.py:99-459 --
constructor, `make_*_block(s)`, `pad_missing_batch_id`, `fps_NN_fast`, `grouped_sparse_conv`, `forward`)
is torch code around spconv-2.x modules.  spconv-2.x is not available here, so the class is compiled from
the reference's source text in place (oracle/ref_inplace.py) with the names it imports bound to the
stand-ins of oracle/ref_stubs.py (sparse-conv sub-modules = the oracle's conv restatement, pinned
separately against the reference's vendored spconv-1.x), `Fsp.sparse_add` bound to oracle.cpu.sparse_add
and the two CUDA-only point ops bound to the pinned C restatements (oracle/ref_assign.py).  The gate MLPs
are the reference's real `nn.Linear`+`ReLU` modules.

What this pins is the module layout / state-dict names and everything *between* the convolutions: the
mask selections, the missing-batch padding, the per-sample NN assignment with its `base` offset, the
cross-gating with the dummy embedding row, the gate on the mixed voxels, the channel pads, the
concatenation order of the unified voxel set and the sparse_add / downscale chain.

Used by tests/test_oracle.py.  Needs /root/reference.
"""
import types

import numpy as np

from . import cpu, ref_stubs
from .ref_assign import _ball_query, _fps
from .ref_inplace import available, load_def  # noqa: F401

REF_ENCODER = 'mmdet3d/models/middle_encoders/sparse_multimodal_encoder_painting.py'


def build(cfg, sd=None, prefix=''):
    """-> the reference's SparseMultiModalEncoderPaint instance built from the config dict."""
    import torch
    st = ref_stubs.classes()

    def sparse_add(a, b):
        oi, of = cpu.sparse_add(np.ascontiguousarray(a.indices.numpy(), np.int32), a.features.numpy(),
                                np.ascontiguousarray(b.indices.numpy(), np.int32), b.features.numpy(), a.spatial_shape)
        return st.Tensor(torch.from_numpy(of), torch.from_numpy(oi), a.spatial_shape, a.batch_size)

    ns = ref_stubs.namespace(st)
    ns.update(furthest_point_sample=_fps, ball_query=_ball_query, Fsp=types.SimpleNamespace(sparse_add=sparse_add))
    cls = load_def(REF_ENCODER, 'SparseMultiModalEncoderPaint', ns, keyword='class')
    net = ref_stubs.bind(cls(**{k: v for k, v in cfg.items() if k != 'type'}), st, sd, prefix).eval()
    net._tensor = st.Tensor
    return net


def layer_table(cfg):
    return ref_stubs.layer_table(build(cfg))


def forward(sd, cfg, v3_list, v2_list, syn3_list, syn2_list, fps_num_list, radius_list, nsample_list,
            thresh_list, seed, prefix=''):
    """Same arguments as oracle.model.multimodal_encoder, except that the per-stage dummy embedding is
    drawn by the reference's own `torch.rand(1, C3)` (:372) after `torch.manual_seed(seed)`.
    -> list of oracle.model.SpTensor."""
    import torch
    net = build(cfg, sd, prefix)

    def t(x):
        return net._tensor(torch.from_numpy(np.ascontiguousarray(x.features, np.float32)),
                           torch.from_numpy(np.ascontiguousarray(x.indices)), x.spatial_shape, x.batch_size)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)     # duplicate-index index_put_ in fps_NN_fast: sequential, last write wins
    try:
        torch.manual_seed(seed)
        with torch.no_grad():
            outs = net([t(x) for x in v3_list], [t(x) for x in v2_list],
                       [torch.from_numpy(np.asarray(s, np.int64)) for s in syn3_list],
                       [torch.from_numpy(np.asarray(s, np.int64)) for s in syn2_list],
                       fps_num_list, radius_list, nsample_list, thresh_list)
    finally:
        torch.set_num_threads(threads)
    return [ref_stubs.to_oracle(o) for o in outs]
