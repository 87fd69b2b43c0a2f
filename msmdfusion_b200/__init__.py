"""msmdfusion_b200 -- B200-native (sm_100a) implementation of MSMDFusion's voxel-space fusion
hot path behind the reference's mmdet3d registry / spconv operator surface.

Layout: ``csrc/`` hand-written CUDA + the C ABI (include/msmd_b200.h); ``_cabi.py`` ctypes
binding; ``ops.py`` tensor-level wrappers; the remaining modules mirror the reference's
Python interface for this path.  There is no CPU fallback.
"""
from . import registry  # noqa: F401
from .registry import (CONV_LAYERS, DETECTORS, FUSION_LAYERS, MIDDLE_ENCODERS,  # noqa: F401
                       VOXEL_ENCODERS, Config, build_from_cfg)
from . import spconv  # noqa: F401
from .sparse_block import SparseBasicBlock, make_sparse_convmodule  # noqa: F401
from .sparse_encoder import SparseEncoder  # noqa: F401
from .voxel import HardSimpleVFE, Voxelization, hard_voxelize, voxelization  # noqa: F401
from . import functional  # noqa: F401
from .fusion_encoder import SparseMultiModalEncoderPaint  # noqa: F401
from .detector import MSMDFusionDetector, SPPModule, TransFusionDetector  # noqa: F401
from .bev_tail import SECOND, SECONDFPN  # noqa: F401  (BEV tail behind the path; SURVEY §8(f) rank 1, cuDNN convs)
from . import loading  # noqa: F401  (virtual-point wire format -> packed scene)
from .loading import PIPELINES  # noqa: F401

__version__ = '0.1.0'
