"""Shared builders for the detector-level tests and the golden generators (CPU-side, deterministic)."""
import os

import numpy as np
import torch

import msmdfusion_b200 as m
from msmdfusion_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DETECTOR_KEYS = ('pts_voxel_layer', 'pts_voxel_encoder', 'pts_middle_encoder', 'multimodal_middle_encoder',
                 'spatial_shapes', 'downscale_factors', 'fps_num_list', 'radius_list', 'max_cluster_samples_list',
                 'dist_thresh_list')


def randomize_bn(module, seed=0):
    g = torch.Generator().manual_seed(seed)
    for mod in module.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            d = mod.weight.device
            mod.weight.data = (torch.rand(mod.weight.shape, generator=g) + 0.5).to(d)
            mod.bias.data = (torch.randn(mod.bias.shape, generator=g) * 0.1).to(d)
            mod.running_mean.data = (torch.randn(mod.running_mean.shape, generator=g) * 0.1).to(d)
            mod.running_var.data = (torch.rand(mod.running_var.shape, generator=g) + 0.5).to(d)


def hotpath_cfg():
    return m.Config.fromfile(os.path.join(ROOT, 'configs', 'msmd_lc_hotpath.py')).hotpath


def build_msmd_detector(seed=0, device=None):
    """Random-init MSMDFusionDetector (parameters drawn on the CPU generator, so identical on every
    machine with this torch build), eval-mode BN with non-trivial statistics, score gate open for roughly
    half of the points."""
    cfg = hotpath_cfg()
    torch.manual_seed(seed)
    det = m.MSMDFusionDetector(**{k: cfg[k] for k in DETECTOR_KEYS})
    if device is not None:
        det = det.to(device)
    randomize_bn(det, seed + 1)
    with torch.no_grad():
        det.score_net[0].weight.mul_(0.2)
        det.score_net[0].bias.fill_(0.05)
    return det.eval(), cfg


def lc_scene(batch, points=None, virtual=(3000, 300)):
    """The seeded LiDAR + camera scene of the detector-level tests: sample 1 has an empty camera."""
    scenes = [synthetic.lidar_scene(30 + b, 1) for b in range(batch)]
    if points is not None:
        scenes = [s[:points] for s in scenes]
    metas = [synthetic.camera_scene(30 + b, scenes[b], virtual_per_camera=virtual[min(b, 1)],
                                    empty_cameras=(() if b == 0 else (2,))) for b in range(batch)]
    fpn = synthetic.fpn_features(1, batch=batch)
    return scenes, metas, fpn


def state_dict_crc(sd):
    import zlib
    crc = 0
    for k in sorted(sd):
        crc = zlib.crc32(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes(), crc)
    return crc
