"""TEST INFRASTRUCTURE.  Runs the reference's own `SparseEncoder` class where it lies.

The class body (mmdet3d/models/middle_encoders/sparse_encoder.py:13-209 -- constructor,
`make_encoder_layers`, `forward`) is compiled from the reference's source text in place
(oracle/ref_inplace.py) with the names it imports bound to the stand-ins of oracle/ref_stubs.py
(`make_sparse_convmodule` / `SparseBasicBlock` record their arguments and run the oracle's conv
restatement; `spconv.SparseSequential` is `nn.Sequential`; `auto_fp16` and the registry decorator are
no-ops).

What this pins is the layer table (which position is SubM / strided / residual block, its channels,
stride and padding, for both `block_type`s), the state-dict names, the order of `encode_features`
and the dense (N, C*D, H, W) view of `forward`.  Used by tests/test_oracle.py.  Needs /root/reference.
"""
import numpy as np

from . import model, ref_stubs
from .ref_inplace import available, load_def  # noqa: F401

REF_FILE = 'mmdet3d/models/middle_encoders/sparse_encoder.py'


def build(cfg, sd=None, prefix=''):
    """-> the reference's SparseEncoder instance built from the config dict (keys of its constructor)."""
    st = ref_stubs.classes()
    cls = load_def(REF_FILE, 'SparseEncoder', ref_stubs.namespace(st), keyword='class')
    return ref_stubs.bind(cls(**{k: v for k, v in cfg.items() if k != 'type'}), st, sd, prefix)


def layer_table(cfg):
    """[(qualified name, spec dict)] in execution order, as the reference's constructor lays them out."""
    return ref_stubs.layer_table(build(cfg))


def forward(sd, cfg, voxel_features, coors, batch_size, prefix=''):
    """-> (spatial_features (B, C*D, H, W) array, [oracle.model.SpTensor] encode_features)."""
    import torch
    enc = build(cfg, sd, prefix)
    with torch.no_grad():
        spatial, feats = enc(torch.from_numpy(np.ascontiguousarray(voxel_features, np.float32)),
                             torch.from_numpy(np.ascontiguousarray(coors)), batch_size)
    return spatial.numpy(), [ref_stubs.to_oracle(f) for f in feats]
