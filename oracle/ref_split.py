"""TEST INFRASTRUCTURE.  Runs the reference's own `voxel_modality_split` arithmetic where it lies.

The reference module (mmdet3d/models/detectors/MSMDFusion.py) cannot be imported in this container
(mmcv / mmdet / spconv-2.x are absent), but its numba merge `type_assign` (:26-45) is a free function
of numpy arrays: it is compiled here from the reference's own source text, read at run time from
/root/reference -- nothing of it is copied into this repository.  The method around it
(`MSMDFusionDetector.voxel_modality_split`, :251-325 -- float key, sort, scatter of the flags, batch
offsets) is plain torch and is run the same way, on CPU tensors.

Used by tests/test_oracle.py (live pin of oracle.cpu.voxel_modality_split) and by
tests/golden/make_golden_split.py (fixtures for the CUDA path).  Needs /root/reference.
"""
import numpy as np

from .ref_inplace import available, load_def  # noqa: F401

REF_DETECTOR = 'mmdet3d/models/detectors/MSMDFusion.py'
_fn = None


def type_assign():
    """The reference's jitted merge, built from MSMDFusion.py's own `def type_assign` block."""
    global _fn
    if _fn is None:
        _fn = load_def(REF_DETECTOR, 'type_assign', {},
                       prefix='from numba import jit\nimport numpy as np\n@jit(nopython=True)\n')
    return _fn


def split_single(coors3, coors2):
    """One sample.  coors*: (n,3) int32 (z,y,x).  -> mix3 (n3,), mix2 (n2,) int32 flags in input row
    order; syn3, syn2: input row ids of the mixed voxels in sorted-key order -- from the whole reference
    method (`voxel_modality_split` below) at B = 1."""
    z = lambda c: np.concatenate([np.zeros((c.shape[0], 1), np.int32), np.asarray(c, np.int32)], 1)  # noqa: E731
    i3, i2, syn3, syn2 = voxel_modality_split(z(coors3), z(coors2), 1)
    return i3[:, 1].copy(), i2[:, 1].copy(), syn3, syn2


class _StableSortTorch:
    """`torch` as the method sees it, with `sort` made stable.  The reference calls `torch.sort(keys)`
    without `stable=`; which of several rows with EQUAL float keys (x / x+1 collisions at z >= 17) comes
    first is then unspecified -- torch's CPU and CUDA kernels each make their own choice -- and it decides
    which of them gets the mix flag.  A stable sort is one of the legal outcomes; it is the one
    oracle.cpu.voxel_modality_split and the CUDA path fix."""

    def __getattr__(self, name):
        import torch
        return getattr(torch, name)

    @staticmethod
    def sort(x, dim=-1):
        import torch
        return torch.sort(x, dim=dim, stable=True)


def voxel_modality_split(indices3, indices2, B):
    """The whole reference method (MSMDFusion.py:251-325) run in place on (N,4) int32 (b,z,y,x) index
    arrays.  -> (indices3 (N3,5) int32 (b,mix,z,y,x), indices2 (N2,5), syn_mix_3D, syn_mix_2D int64)."""
    import types

    import torch
    import torch.nn.functional as F
    fn = load_def(REF_DETECTOR, 'voxel_modality_split',
                  {'torch': _StableSortTorch(), 'F': F, 'type_assign': type_assign()})
    v3 = types.SimpleNamespace(indices=torch.from_numpy(np.ascontiguousarray(indices3)))
    v2 = types.SimpleNamespace(indices=torch.from_numpy(np.ascontiguousarray(indices2)))
    v3, v2, s3, s2 = fn(None, v3, v2, B)
    return v3.indices.numpy(), v2.indices.numpy(), s3.numpy(), s2.numpy()
