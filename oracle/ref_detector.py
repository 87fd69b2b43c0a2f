"""TEST INFRASTRUCTURE.  Runs the reference's own `MSMDFusionDetector.extract_pts_feat` where it lies.

`MSMDFusionDetector` (mmdet3d/models/detectors/MSMDFusion.py:92-) derives from mmdet's
`MVXTwoStageDetector`, which cannot be imported here, so its constructor cannot run.  Everything the
voxel-space hot path executes *can*: the methods `extract_pts_feat`, `voxelize`, `extract_multiscale_voxel_feat`,
`depth_aware_channel_compression`, `fetch_2D_voxels`, `get_foreground2D`, `voxelize_fg_pcd`,
`voxel_modality_split` (+ numba `type_assign`), the classes `SPPModule`, `HardSimpleVFE`
(voxel_encoders/voxel_encoder.py:15-46), `Voxelization` / `_Voxelization` (ops/voxel/voxelize.py:10-122),
`SparseEncoder` and `SparseMultiModalEncoderPaint` are all compiled from the reference's source text in
place (oracle/ref_inplace.py) and assembled on a stand-in object:

* `hard_voxelize` inside `_Voxelization` is the reference's OWN C++ CPU op, compiled from
  /root/reference into oracle/_ref (oracle/build.py:build_ref);
* the sparse-conv blocks inside the two encoders are the oracle's conv restatement (oracle/ref_stubs.py;
  pinned separately against the reference's vendored spconv-1.x), `Fsp.sparse_add` is oracle.cpu.sparse_add
  and the two CUDA-only point ops are the pinned C restatements;
* `conv1x1_blocks` and `score_net`, which the reference builds in the constructor that cannot run, are
  built here with the same layer arguments (:108-129) and loaded strictly from the caller's state dict;
* `pts_backbone` / `pts_neck` (dense 2D, out of scope) are left out: `bev_fusion` is replaced by a probe
  that records the (B, 640, 180, 180) tensor it is handed -- the end of the voxel-space path -- and returns.

Used by tests/test_oracle.py and tests/golden/make_golden_detector.py.  Needs /root/reference.
"""
import importlib.util
import types

import numpy as np

from . import build as _build
from . import model, ref_encoder, ref_lidar_encoder, ref_split, ref_stubs
from .ref_inplace import available, load_def  # noqa: F401

REF_DETECTOR = 'mmdet3d/models/detectors/MSMDFusion.py'
METHODS = ('extract_pts_feat', 'voxelize', 'extract_multiscale_voxel_feat', 'depth_aware_channel_compression',
           'fetch_2D_voxels', 'get_foreground2D', 'voxelize_fg_pcd', 'voxel_modality_split')


def _reference_voxel_op():
    so = _build.ref_so_path() or _build.build_ref()
    if so is None:
        raise RuntimeError('oracle/_ref is not built and /root/reference is absent')
    spec = importlib.util.spec_from_file_location(_build.REF_NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def spp_module():
    """The reference's SPPModule class (MSMDFusion.py:47-90), real torch layers."""
    import torch
    from torch import nn
    return load_def(REF_DETECTOR, 'SPPModule', {'nn': nn, 'torch': torch}, keyword='class')


def build(sd, cfg):
    """sd: state dict with this package's / the reference's parameter names; cfg: hot-path config."""
    import torch
    import torch.nn.functional as F
    from torch import nn
    from torch.autograd import Function
    from torch.nn.modules.utils import _pair
    st = ref_stubs.classes()
    identity = lambda *a, **k: (lambda f: f)  # noqa: E731
    op = _reference_voxel_op()
    ns = {'torch': torch, 'F': F, 'nn': nn, 'np': np, 'type_assign': ref_split.type_assign(),
          'spconv': types.SimpleNamespace(SparseConvTensor=st.Tensor)}
    ns['torch'] = ref_split._StableSortTorch()      # see oracle/ref_split.py: one legal, pinned sort order
    me = type('ReferenceDetector', (), {name: load_def(REF_DETECTOR, name, ns) for name in METHODS})()

    vns = {'torch': torch, 'nn': nn, 'Function': Function, '_pair': _pair, 'hard_voxelize': op.hard_voxelize,
           'dynamic_voxelize': None}
    vns['_Voxelization'] = load_def('mmdet3d/ops/voxel/voxelize.py', '_Voxelization', vns, keyword='class')
    vns['voxelization'] = vns['_Voxelization'].apply
    Voxelization = load_def('mmdet3d/ops/voxel/voxelize.py', 'Voxelization', vns, keyword='class')
    HardSimpleVFE = load_def('mmdet3d/models/voxel_encoders/voxel_encoder.py', 'HardSimpleVFE',
                             {'nn': nn, 'torch': torch, 'force_fp32': identity}, keyword='class')

    me.with_pts_bbox, me.with_pts_neck = True, False
    for k in ('spatial_shapes', 'downscale_factors', 'fps_num_list', 'radius_list', 'max_cluster_samples_list',
              'dist_thresh_list'):
        setattr(me, k, cfg[k])
    me.pts_voxel_layer = Voxelization(**cfg['pts_voxel_layer']).eval()
    me.pts_voxel_encoder = HardSimpleVFE(**{k: v for k, v in cfg['pts_voxel_encoder'].items() if k != 'type'})
    me.pts_middle_encoder = ref_lidar_encoder.build(dict(cfg['pts_middle_encoder']), sd, 'pts_middle_encoder.').eval()
    me.multimodal_middle_encoder = ref_encoder.build(dict(cfg['multimodal_middle_encoder']), sd,
                                                     'multimodal_middle_encoder.')

    def compress(k):   # :108-125
        return nn.Sequential(nn.Conv2d(256 + 1, 49, kernel_size=k, stride=1, padding=k // 2, bias=False),
                         nn.BatchNorm2d(49, eps=0.001, momentum=0.01, affine=True, track_running_stats=True), nn.ReLU())
    me.conv1x1_blocks = nn.ModuleList([compress(5), compress(5), compress(3)]).eval()
    me.score_net = nn.Sequential(nn.Linear(50 + 16, 1), nn.ReLU())      # :126-129
    for name, mod in (('conv1x1_blocks', me.conv1x1_blocks), ('score_net', me.score_net)):
        mod.load_state_dict({k: torch.as_tensor(model._np(sd, f'{name}.{k}')) for k in mod.state_dict()})

    me.probe = {}

    class Stop(Exception):
        pass

    def bev_probe(x):
        me.probe['bev'] = x
        raise Stop()
    me.bev_fusion, me._stop = bev_probe, Stop
    mm_forward = me.multimodal_middle_encoder.forward

    def mm_probe(*a):
        me.probe['stage_outs'] = mm_forward(*a)
        return me.probe['stage_outs']
    me.multimodal_middle_encoder.forward = mm_probe
    return me


def extract_voxel_space(sd, cfg, scenes, fpn_feats, img_metas, seed):
    """Run the reference's extract_pts_feat up to bev_fusion.  scenes: list of (N,5) arrays; fpn_feats: three
    (B*6, 256, h, w) arrays; the per-stage dummy embeddings are the reference's own torch.rand draws after
    torch.manual_seed(seed).  -> (bev array (B, 640, 180, 180), [oracle.model.SpTensor] stage_outs)."""
    import torch

    from .ref_lift import _metas_for_reference
    me = build(sd, cfg)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)   # duplicate-index index_put_ (depth canvas, fps_NN_fast): sequential, last wins
    try:
        torch.manual_seed(seed)
        with torch.no_grad():
            try:
                me.extract_pts_feat([torch.from_numpy(np.ascontiguousarray(s, np.float32)) for s in scenes],
                                    [torch.from_numpy(np.ascontiguousarray(f, np.float32)) for f in fpn_feats],
                                    _metas_for_reference(img_metas))
            except me._stop:
                pass
    finally:
        torch.set_num_threads(threads)
    return me.probe['bev'].numpy(), [ref_stubs.to_oracle(o) for o in me.probe['stage_outs']]
