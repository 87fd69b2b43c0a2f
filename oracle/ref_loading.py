"""TEST INFRASTRUCTURE.  Runs the reference's own foreground-2D loading pipeline where it lies.

SURVEY §8(f) rank 3 -- the virtual-point wire format and its loader.  The five pipeline classes of
mmdet3d/datasets/pipelines/my_loading_multi_proj.py (`LoadForeground2D` :14-160,
`LoadForeground2DFromMultiSweeps` :162-337, `GlobalRotTransFilterForeground2D` :341-416,
`ImgScaleCropFlipForeground2D` :419-455, `ShuffleForeground2D` :457-489) are plain numpy / torch code
behind imports that do not exist here (mmcv, mmdet).  They are compiled from the reference's source text
in place (oracle/ref_inplace.py); `get_points_type('LIDAR')` is bound to the reference's own
`LiDARPoints` / `BasePoints` classes (mmdet3d/core/points/{lidar_points,base_points}.py), compiled the
same way.  Nothing is copied into this repository.

Used by tests/test_loading.py and tests/tools/bench_loader.py.  Needs /root/reference.
"""
import os
from abc import abstractmethod
from functools import partial

import numpy as np

from .ref_inplace import available, load_def  # noqa: F401

REF_FILE = 'mmdet3d/datasets/pipelines/my_loading_multi_proj.py'
_cache = {}


def classes():
    """-> dict name -> the reference class."""
    if _cache:
        return _cache
    import torch
    base = load_def('mmdet3d/core/points/base_points.py', 'BasePoints',
                    {'np': np, 'torch': torch, 'abstractmethod': abstractmethod}, keyword='class')
    lidar = load_def('mmdet3d/core/points/lidar_points.py', 'LiDARPoints', {'BasePoints': base}, keyword='class')
    ns = {'np': np, 'os': os, 'partial': partial, 'torch': torch, 'get_points_type': lambda kind: lidar}
    for name in ('LoadForeground2D', 'LoadForeground2DFromMultiSweeps', 'GlobalRotTransFilterForeground2D',
                 'ImgScaleCropFlipForeground2D', 'ShuffleForeground2D'):
        _cache[name] = load_def(REF_FILE, name, ns, keyword='class')
    _cache['LiDARPoints'] = lidar
    return _cache


def test_pipeline(point_cloud_range, sweeps_num=10, dataset='NuScenesDataset'):
    """The foreground part of the reference's test pipeline (configs/MSMDFusion_nusc_voxel_LC.py:97-103),
    in its order."""
    c = classes()
    multi = c['LoadForeground2DFromMultiSweeps'](dataset=dataset, sweeps_num=sweeps_num)
    multi.test_mode = True   # read at :300 but never set by the class itself
    return [c['LoadForeground2D'](dataset=dataset), multi,
            c['GlobalRotTransFilterForeground2D'](point_cloud_range=point_cloud_range),
            c['ImgScaleCropFlipForeground2D']()]


def run(stages, results):
    for stage in stages:
        results = stage(results)
    return results
