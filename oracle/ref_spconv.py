"""The reference's own VENDORED spconv-1.x (mmdet3d/ops/spconv/src/*.cc,*.cu, include/spconv/*.h),
compiled unmodified from /root/reference into oracle/_ref/ and run on the CPU.

TEST INFRASTRUCTURE (see oracle/README.md).  The hot path itself uses the un-vendored spconv v2.1.21,
but spconv-1.x is the same operator family from the same author: identical cross-correlation
arithmetic (gather - GEMM - scatter-add per kernel offset, ``spconv_ops.h:299-356``), identical
kernel-offset <-> output mapping (``geometry.h:24-85``) and output-size rule (``ops.py:20-31``).  It
differs only in row order of strided-conv outputs on the CPU (first touch instead of ascending
linear index) and in the weight layout ([kz,ky,kx,Cin,Cout] instead of KRSC).  It is therefore used
to PIN the oracle's conv / rulebook restatement against code that lives in the reference tree:
results are compared after sorting output rows by linear index.
"""
import glob
import os

import numpy as np
import torch

from . import build as _build

REF_ROOT = _build.REF_ROOT
REF_DIR = _build.REF_DIR
NAME = 'ref_spconv1x'
_MOD = None


def so_path():
    hits = glob.glob(os.path.join(REF_DIR, NAME + '*.so'))
    return hits[0] if hits else None


def build(verbose=False):
    """Compile the vendored extension where its sources lie (needs /root/reference); ~40 s."""
    if so_path():
        return so_path()
    src = os.path.join(REF_ROOT, 'mmdet3d', 'ops', 'spconv', 'src')
    inc = os.path.join(REF_ROOT, 'mmdet3d', 'ops', 'spconv', 'include')
    if not os.path.isdir(src):
        return None
    from torch.utils.cpp_extension import load
    bd = os.path.join(REF_DIR, '_obj_spconv1x')
    os.makedirs(bd, exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
    load(name=NAME, sources=[os.path.join(src, f) for f in (
        'all.cc', 'indice.cc', 'reordering.cc', 'maxpool.cc', 'indice_cuda.cu', 'reordering_cuda.cu',
        'maxpool_cuda.cu')], extra_include_paths=[inc], build_directory=bd, extra_cflags=['-O2', '-w'],
        extra_cuda_cflags=['-O2', '-w'], with_cuda=True, verbose=verbose)
    so = glob.glob(os.path.join(bd, NAME + '*.so'))
    if not so:
        return None
    dst = os.path.join(REF_DIR, os.path.basename(so[0]))
    os.replace(so[0], dst)
    return dst


def module():
    global _MOD
    if _MOD is None:
        path = so_path() or build()
        if path is None:
            return None
        import importlib.util
        spec = importlib.util.spec_from_file_location(NAME, path)
        _MOD = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_MOD)
    return _MOD


def _triple(v):
    return [int(x) for x in v] if isinstance(v, (list, tuple)) else [int(v)] * 3


def conv(indices, features, weight_krsc, spatial_shape, batch_size, ksize, stride=1, padding=0,
         dilation=1, subm=False):
    """Run the reference's spconv-1.x CPU ops.  weight_krsc: [Cout,kz,ky,kx,Cin] (spconv-2.x layout).
    Returns (out_indices (M,4) i32, out_features (M,Cout) f32) with rows sorted by linear index
    (SubM: the input order is kept, as spconv defines)."""
    m = module()
    assert m is not None, 'reference spconv-1.x is not built (needs /root/reference)'
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    shape = [int(s) for s in spatial_shape]
    if subm:
        out_shape = shape
    else:
        out_shape = [(shape[i] + 2 * pd[i] - dl[i] * (ks[i] - 1) - 1) // st[i] + 1 for i in range(3)]
    idx = torch.from_numpy(np.ascontiguousarray(indices, np.int32))
    feat = torch.from_numpy(np.ascontiguousarray(features, np.float32))
    w = torch.from_numpy(np.ascontiguousarray(weight_krsc, np.float32))
    filters = w.permute(1, 2, 3, 4, 0).contiguous()  # [kz,ky,kx,Cin,Cout]
    outids, pairs, pair_num = m.get_indice_pairs_3d(idx, int(batch_size), out_shape, shape, ks, st, pd, dl,
                                                    [0, 0, 0], int(subm), 0)
    out = m.indice_conv_fp32(feat, filters, pairs, pair_num, outids.shape[0], 0, int(subm))
    oi, of = outids.numpy().astype(np.int32), out.numpy().astype(np.float32)
    if not subm:
        lin = ((oi[:, 0].astype(np.int64) * out_shape[0] + oi[:, 1]) * out_shape[1] + oi[:, 2]) * out_shape[2] + oi[:, 3]
        order = np.argsort(lin, kind='stable')
        oi, of = oi[order], of[order]
    return oi, of, out_shape


def conv_backward(indices, features, weight_krsc, grad_out_sorted, spatial_shape, batch_size, ksize, stride=1,
                  padding=0, dilation=1, subm=False):
    """The reference's ``indice_conv_backward_fp32`` (spconv_ops.h:364-457) on the CPU.
    ``grad_out_sorted`` is given in the row order ``conv`` returns (ascending linear index for strided
    convs, input order for SubM).  Returns (grad_features (N_in,Cin), grad_weight KRSC)."""
    m = module()
    assert m is not None, 'reference spconv-1.x is not built (needs /root/reference)'
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    shape = [int(s) for s in spatial_shape]
    if subm:
        out_shape = shape
    else:
        out_shape = [(shape[i] + 2 * pd[i] - dl[i] * (ks[i] - 1) - 1) // st[i] + 1 for i in range(3)]
    idx = torch.from_numpy(np.ascontiguousarray(indices, np.int32))
    feat = torch.from_numpy(np.ascontiguousarray(features, np.float32))
    w = torch.from_numpy(np.ascontiguousarray(weight_krsc, np.float32))
    filters = w.permute(1, 2, 3, 4, 0).contiguous()  # [kz,ky,kx,Cin,Cout]
    outids, pairs, pair_num = m.get_indice_pairs_3d(idx, int(batch_size), out_shape, shape, ks, st, pd, dl,
                                                    [0, 0, 0], int(subm), 0)
    g = np.ascontiguousarray(grad_out_sorted, np.float32)
    if not subm:  # un-sort: the reference's own output rows are in first-touch order
        oi = outids.numpy().astype(np.int64)
        lin = ((oi[:, 0] * out_shape[0] + oi[:, 1]) * out_shape[1] + oi[:, 2]) * out_shape[2] + oi[:, 3]
        order = np.argsort(lin, kind='stable')
        native = np.empty_like(g)
        native[order] = g
        g = native
    gin, gfil = m.indice_conv_backward_fp32(feat, filters, torch.from_numpy(g), pairs, pair_num, 0, int(subm))
    return gin.numpy().astype(np.float32), gfil.permute(4, 0, 1, 2, 3).contiguous().numpy().astype(np.float32)
