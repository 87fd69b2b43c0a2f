"""TEST INFRASTRUCTURE.  Runs the reference's own 2D->3D lift methods where they lie.

`MSMDFusionDetector.get_foreground2D` (mmdet3d/models/detectors/MSMDFusion.py:169-238) and
`.depth_aware_channel_compression` (:335-369) are plain torch code over `img_metas`.  Their method
bodies are compiled from the reference's source text in place (oracle/ref_inplace.py) and called with
a stand-in `self` that carries only the sub-modules they touch (`score_net`, `conv1x1_blocks`), built
with the reference's constructor arguments (:108-129) and loaded with the caller's weights.

Used by tests/test_oracle.py to pin oracle.cpu.lift_gather / oracle.model.get_foreground2d /
oracle.model.depth_canvas.  Needs /root/reference.
"""
import types

import numpy as np

from .ref_inplace import available, load_def  # noqa: F401

REF_DETECTOR = 'mmdet3d/models/detectors/MSMDFusion.py'


def _metas_for_reference(img_metas):
    """The dataset hands `fg_points` over as LiDARPoints (attribute `.tensor`); arrays here."""
    import torch
    out = []
    for m in img_metas:
        info = dict(m['foreground2D_info'])
        info['fg_points'] = [types.SimpleNamespace(tensor=torch.from_numpy(np.ascontiguousarray(p, np.float32)))
                             for p in info['fg_points']]
        m = dict(m)
        m['foreground2D_info'] = info
        out.append(m)
    return out


def _single_thread(fn, *args):
    """index_put_ with duplicate pixels is sequential (last write wins) on one intra-op thread."""
    import torch
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        with torch.no_grad():
            return fn(*args)
    finally:
        torch.set_num_threads(threads)


def get_foreground2d(img_feats, img_metas, score_w, score_b):
    """img_feats (B*ncam, C, h, w) f32 array; score_net = Linear(C+17, 1)+ReLU with the given weights.
    -> list (len B) of (M_b, 15+C) arrays, from the reference's method."""
    import torch
    from torch import nn
    fn = load_def(REF_DETECTOR, 'get_foreground2D', {'torch': torch})
    C = img_feats.shape[1]
    score_net = nn.Sequential(nn.Linear(C + 17, 1), nn.ReLU())          # :126-129 (50+16 at C = 49)
    score_net[0].weight.data = torch.from_numpy(np.asarray(score_w, np.float32).reshape(1, -1).copy())
    score_net[0].bias.data = torch.from_numpy(np.asarray(score_b, np.float32).reshape(1).copy())
    me = types.SimpleNamespace(score_net=score_net)
    out = _single_thread(fn, me, torch.from_numpy(np.ascontiguousarray(img_feats)), _metas_for_reference(img_metas))
    return [o.numpy() for o in out]


def depth_maps(feat_shapes, img_metas):
    """The bilinear-resampled sparse depth channel the reference concatenates to each FPN level
    (:337-366), obtained by running the method with identity `conv1x1_blocks` on zero features.
    feat_shapes: [(h,w)]*3 -> list of (B*6, 1, h, w) arrays."""
    import torch
    import torch.nn.functional as F
    from torch import nn
    fn = load_def(REF_DETECTOR, 'depth_aware_channel_compression', {'torch': torch, 'F': F})
    me = types.SimpleNamespace(conv1x1_blocks=[nn.Identity()] * 3)
    B = len(img_metas)
    feats = [torch.zeros(B * 6, 1, h, w) for h, w in feat_shapes]
    out = _single_thread(fn, me, feats, _metas_for_reference(img_metas))
    return [o[:, 1:].numpy() for o in out]
