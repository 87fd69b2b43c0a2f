"""Voxelization layer and HardSimpleVFE -- mirrors ``mmdet3d/ops/voxel/voxelize.py`` and
``mmdet3d/models/voxel_encoders/voxel_encoder.py:14-47`` on top of the hash-based CUDA
``hard_voxelize`` (csrc/voxelize.cu).  No CPU path.
"""
import torch
from torch import nn
from torch.nn.modules.utils import _pair

from . import _cabi, ops
from .registry import VOXEL_ENCODERS


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range,
                  max_points, max_voxels, NDim=3):
    """Drop-in for ``voxel_layer.hard_voxelize`` (mmdet3d/ops/voxel/src/voxelization.h:61-78).

    The caller pre-allocates ``voxels (max_voxels,max_points,C)``, ``coors (max_voxels,3)``,
    ``num_points_per_voxel (max_voxels,)``; they are filled in place and ``voxel_num`` is
    returned (one device->host read, as the reference does).
    """
    assert NDim == 3
    _cabi.require_cuda(points, 'hard_voxelize: points must be a CUDA tensor')
    v, c, n, _ = ops.hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels,
                                   want_voxels=True, out=(voxels, coors, num_points_per_voxel))
    k = v.shape[0]
    if v.data_ptr() != voxels.data_ptr():   # buffers the kernels cannot write directly (dtype, strides, capacity)
        voxels[:k].copy_(v)
        coors[:k].copy_(c)
        num_points_per_voxel[:k].copy_(n)
    return k


def voxelization(points, voxel_size, coors_range, max_points=35, max_voxels=20000):
    """``_Voxelization.forward`` (mmdet3d/ops/voxel/voxelize.py:13-59) without the zero-filled
    (max_voxels,max_points,C) staging buffers: outputs are produced at their final size."""
    if max_points == -1 or max_voxels == -1:
        raise NotImplementedError('dynamic voxelization is outside the MSMDFusion hot path')
    with torch.no_grad():
        voxels, coors, num, _ = ops.hard_voxelize(points, voxel_size, coors_range, max_points,
                                                  max_voxels, want_voxels=True)
    return voxels, coors, num


class Voxelization(nn.Module):
    """mmdet3d/ops/voxel/voxelize.py:65-114 (same constructor, attributes and repr)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else _pair(max_voxels)
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        vs = torch.tensor(voxel_size, dtype=torch.float32)
        grid_size = torch.round((pcr[3:] - pcr[:3]) / vs).long()
        self.grid_size = grid_size
        self.pcd_shape = [*grid_size[:2], 1][::-1]

    def current_max_voxels(self):
        return self.max_voxels[0] if self.training else self.max_voxels[1]

    def forward(self, input):
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points,
                            self.current_max_voxels())

    def forward_mean(self, input, num_features, batch_idx=None):
        """Fused Voxelization + HardSimpleVFE: returns (mean (V,F), coors, num_points) without
        ever materialising the (V,max_points,C) voxel buffer."""
        with torch.no_grad():
            _, coors, num, mean = ops.hard_voxelize(
                input, self.voxel_size, self.point_cloud_range, self.max_num_points,
                self.current_max_voxels(), want_voxels=False, mean_features=num_features,
                batch_idx=batch_idx)
        return mean, coors, num

    def __repr__(self):
        return (f'{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range='
                f'{self.point_cloud_range}, max_num_points={self.max_num_points}, max_voxels='
                f'{self.max_voxels})')


@VOXEL_ENCODERS.register_module()
class HardSimpleVFE(nn.Module):
    """mmdet3d/models/voxel_encoders/voxel_encoder.py:14-47: mean of the points of a voxel."""

    def __init__(self, num_features=4):
        super().__init__()
        self.num_features = num_features
        self.fp16_enabled = False

    def forward(self, features, num_points, coors):
        points_mean = features[:, :, :self.num_features].sum(dim=1, keepdim=False) / \
            num_points.type_as(features).view(-1, 1)
        return points_mean.contiguous()
