"""ctypes binding of libmsmd_b200.so (the C ABI declared in include/msmd_b200.h).

This is the only place the product touches native code.  There is no CPU fallback: if
the shared library is missing or an entry point fails, a RuntimeError is raised.
PyTorch is used for device memory and streams only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_C', 'libmsmd_b200.so')
if os.environ.get('MSMD_LIB'):   # debug builds of the same ABI (tools/tc_trace.py: the -DMSMD_TC_TRACE build)
    LIB_PATH = os.environ['MSMD_LIB']

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_float_p = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p
_i = ctypes.c_int
_sz = ctypes.c_size_t

class ConvLayer(ctypes.Structure):
    """msmd_conv_layer (include/msmd_b200.h)."""
    _fields_ = [('subm', ctypes.c_int), ('ksize', ctypes.c_int * 3), ('stride', ctypes.c_int * 3),
                ('padding', ctypes.c_int * 3), ('dilation', ctypes.c_int * 3), ('cin', ctypes.c_int),
                ('cout', ctypes.c_int), ('weight', ctypes.c_void_p), ('weight_tc', ctypes.c_int),
                ('scale', ctypes.c_void_p), ('shift', ctypes.c_void_p), ('relu', ctypes.c_int),
                ('input', ctypes.c_int), ('residual', ctypes.c_int)]


class SparseDesc(ctypes.Structure):
    """msmd_sparse_desc (include/msmd_b200.h)."""
    _fields_ = [('features', ctypes.c_void_p), ('indices', ctypes.c_void_p), ('n', ctypes.c_int),
                ('channels', ctypes.c_int), ('spatial_shape', ctypes.c_int * 3)]


class GmaStage(ctypes.Structure):
    """msmd_gma_stage (include/msmd_b200.h)."""
    _fields_ = [('only3d', ctypes.POINTER(ConvLayer)), ('n_only3d', ctypes.c_int),
                ('agg', ctypes.POINTER(ConvLayer)), ('n_agg', ctypes.c_int),
                ('down', ctypes.POINTER(ConvLayer)), ('n_down', ctypes.c_int),
                ('w_cross', ctypes.c_void_p), ('b_cross', ctypes.c_void_p),
                ('w_gate', ctypes.c_void_p), ('b_gate', ctypes.c_void_p),
                ('c3', ctypes.c_int), ('c2', ctypes.c_int)]


# name -> (restype, argtypes).  Device pointers travel as void*; host arrays as int*/float*.
SIGNATURES = {
    'msmd_last_error': (ctypes.c_char_p, []),
    'msmd_abi_version': (_i, []),
    'msmd_launch_count': (ctypes.c_ulonglong, []),
    'msmd_hard_voxelize_workspace': (_sz, [_i]),
    'msmd_hard_voxelize': (_i, [_vp, _i, _i, _c_float_p, _c_float_p, _i, _i, _vp, _vp, _i, _i, _vp,
                                _vp, _i, _vp, _vp, _sz, _vp]),
    'msmd_grid_num_words': (_sz, [_i, _c_int_p]),
    'msmd_scan_workspace': (_sz, []),
    'msmd_grid_build': (_i, [_vp, _i, _i, _c_int_p, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'msmd_rulebook_subm': (_i, [_vp, _i, _i, _c_int_p, _c_int_p, _c_int_p, _vp, _vp, _vp, _vp, _vp]),
    'msmd_conv_out_shape': (_i, [_c_int_p] * 6),
    'msmd_rulebook_conv_outputs': (_i, [_vp, _i, _i, _c_int_p, _c_int_p, _c_int_p, _c_int_p,
                                        _c_int_p, _vp, _vp, _vp, _vp, _sz, _vp]),
    'msmd_rulebook_conv_pairs': (_i, [_vp, _vp, _i, _i, _c_int_p, _c_int_p, _c_int_p, _c_int_p,
                                      _c_int_p, _vp, _vp, _vp, _vp, _vp, _vp]),
    'msmd_spconv_pack_weight': (_i, [_vp, _i, _i, _i, _vp, _vp]),
    'msmd_spconv_fwd': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    'msmd_spconv_tc_supported': (_i, [_i, _i, _i]),
    'msmd_spconv_tc_set_variant': (_i, [_i]),
    'msmd_spconv_tc_set_tuning': (_i, [_i, _i]),
    'msmd_spconv_tc_packed_floats': (_sz, [_i, _i, _i]),
    'msmd_spconv_tc_pack_weight': (_i, [_vp, _i, _i, _i, _vp, _vp]),
    'msmd_spconv_fwd_tc': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    'msmd_spconv_tc_workspace': (_sz, [_i, _i]),
    'msmd_spconv_fwd_tc_ws': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    'msmd_spconv_tc16_packed_bytes': (_sz, [_i, _i, _i, _i]),
    'msmd_spconv_tc16_pack_weight': (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    'msmd_spconv_fwd_tc16': (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    'msmd_gather_rows': (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    'msmd_gma_assemble': (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _vp]),
    'msmd_gma_stage_forward': (_i, [ctypes.POINTER(GmaStage), _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp,
                                    _vp, _i, _vp, _vp, _vp, _i, _i, _c_int_p, _vp, _sz, ctypes.POINTER(SparseDesc), _vp,
                                    _vp]),
    'msmd_executor_geometry_stream': (_i, [ctypes.POINTER(ctypes.c_void_p)]),
    'msmd_split_width': (_i, [_i]),
    'msmd_split_bf16': (_i, [_vp, _i, _i, _vp, _vp]),
    'msmd_spconv_sb_packed_bytes': (_sz, [_i, _i, _i]),
    'msmd_spconv_sb_set_variant': (_i, [_i]),
    'msmd_spconv_sb_uses_tile_masks': (_i, []),
    'msmd_spconv_sb_set_pdl': (_i, [_i]),
    'msmd_rulebook_tile_masks': (_i, [_vp, _i, _i, _vp, _vp]),
    'msmd_spconv_fwd_sb_ex': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    'msmd_spconv_sb_pack_weight': (_i, [_vp, _i, _i, _i, _vp, _vp]),
    'msmd_spconv_fwd_sb': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    'msmd_spconv_tc16_set_variant': (_i, [_i]),
    'msmd_spconv_tc16_workspace': (_sz, [_i, _i]),
    'msmd_spconv_fwd_tc16_ws': (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _sz,
                                    _vp]),
    'msmd_rulebook_mask_sort_workspace': (_sz, [_i]),
    'msmd_rulebook_mask_sort': (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    'msmd_spconv_fwd_tc_sorted': (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp,
                                       _sz, _vp]),
    'msmd_spconv_set_mask_sort': (_i, [_i]),
    'msmd_rulebook_transpose': (_i, [_vp, _i, _i, _i, _vp, _vp]),
    'msmd_spconv_transpose_weight': (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    'msmd_spconv_bwd_data': (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    'msmd_spconv_bwd_weight_workspace': (_sz, [_i, _i, _i, _i]),
    'msmd_spconv_bwd_weight': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    'msmd_spconv_bwd_weight_tc_supported': (_i, [_i, _i, _i]),
    'msmd_spconv_bwd_weight_tc_workspace': (_sz, [_i, _i, _i, _i]),
    'msmd_spconv_bwd_weight_tc': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    'msmd_spconv_set_wgrad_tc': (_i, [_i]),
    'msmd_from_dense': (_i, [_vp, _vp, _i, _i, _i, _c_int_p, _vp, _vp]),
    'msmd_grid_rows': (_i, [_vp, _i, _i, _c_int_p, _vp, _vp, _vp, _vp]),
    'msmd_sparse_net_forward': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _c_int_p, _vp, _sz, _vp, _vp]),
    'msmd_sparse_net_forward_ex': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _c_int_p, _vp, _sz, _vp, ctypes.POINTER(_sz), _i,
                                        _vp]),
    'msmd_to_dense': (_i, [_vp, _vp, _i, _i, _i, _c_int_p, _vp, _vp]),
    'msmd_fps_workspace': (_sz, [_i]),
    'msmd_fps_set_threads': (_i, [_i]),
    'msmd_fps': (_i, [_vp, _i, _i, _vp, _vp, _sz, _vp]),
    'msmd_ball_query': (_i, [_vp, _i, _vp, _i, ctypes.c_float, ctypes.c_float, _i, _vp, _vp]),
    'msmd_nn_search': (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    'msmd_group_assign': (_i, [_vp, _i, _i, _vp, _vp, ctypes.c_float, _i, _i, _vp, _vp, _vp]),
    'msmd_modality_split_workspace': (_sz, [_i, _i]),
    'msmd_modality_split': (_i, [_vp, _i, _vp, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _sz, _vp]),
    'msmd_modality_split_sort': (_i, [_vp, _i, _vp, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, _vp,
                                      _vp, _vp, _vp, _vp, _sz, _vp]),
    'msmd_compact_unflagged': (_i, [_vp, _i, _vp, _vp, _vp, _sz, _vp]),
    'msmd_sparse_add_outputs': (_i, [_vp, _i, _vp, _i, _i, _c_int_p, _vp, _vp, _vp, _vp, _sz, _vp]),
    'msmd_sparse_add_finish': (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _c_int_p, _vp,
                                    _vp, _vp]),
    'msmd_lift_gather': (_i, [_vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                              ctypes.c_longlong, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp,
                              ctypes.c_float, _vp, ctypes.c_float, _vp, _vp]),
}

_LIB = None


def lib():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m msmdfusion_b200.build` '
                '(or __graft_entry__.build()).  msmdfusion_b200 has no CPU fallback.')
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.msmd_abi_version() != 1:
            raise RuntimeError('libmsmd_b200.so ABI version mismatch')
        if os.environ.get('MSMD_MASK_SORT', '0') not in ('', '0'):  # opt-in: mask-sorted tiles (executor path)
            L.msmd_spconv_set_mask_sort(1)
        if os.environ.get('MSMD_WGRAD_TC', '0') not in ('', '0'):  # opt-in: tensor-core weight gradient
            L.msmd_spconv_set_wgrad_tc(1)
        if os.environ.get('MSMD_TC_TUNE'):  # e.g. "occ=1,stages=3,split=1,cps=2": launch-heuristic A/B switches
            for item in os.environ['MSMD_TC_TUNE'].split(','):
                name, _, val = item.partition('=')
                if L.msmd_spconv_tc_set_tuning({'occ': 0, 'stages': 1, 'split': 2, 'cps': 3, 'epi': 4}[name.strip()], int(val)) != 0:
                    raise RuntimeError('bad MSMD_TC_TUNE')
        if os.environ.get('MSMD_TC16_VARIANT'):  # 16-bit modes: 2 = A via shared memory (default), 3 = A via tensor memory
            if L.msmd_spconv_tc16_set_variant(int(os.environ['MSMD_TC16_VARIANT'])) != 0:
                raise RuntimeError('bad MSMD_TC16_VARIANT')
        if os.environ.get('MSMD_SB_VARIANT'):  # schedule of the split-operand conv kernel: 1 tile per CTA | 2 persistent
            if L.msmd_spconv_sb_set_variant(int(os.environ['MSMD_SB_VARIANT'])) != 0:
                raise RuntimeError('bad MSMD_SB_VARIANT')
        if os.environ.get('MSMD_SB_PDL'):  # programmatic dependent launch of the persistent conv kernel (default off)
            L.msmd_spconv_sb_set_pdl(int(os.environ['MSMD_SB_PDL']))
        if os.environ.get('MSMD_FPS_THREADS'):  # A/B switch of the cluster FPS kernel's CTA width
            if L.msmd_fps_set_threads(int(os.environ['MSMD_FPS_THREADS'])) != 0:
                raise RuntimeError('bad MSMD_FPS_THREADS')
        if os.environ.get('MSMD_TC_VARIANT'):  # A/B switch of the tensor-core conv kernel (2 | 3)
            if L.msmd_spconv_tc_set_variant(int(os.environ['MSMD_TC_VARIANT'])) != 0:
                raise RuntimeError('bad MSMD_TC_VARIANT')
        _LIB = L
    return _LIB


def check(status, what):
    if status != 0:
        msg = lib().msmd_last_error().decode(errors='replace')
        raise RuntimeError(f'{what} failed ({status}): {msg}')


def require_cuda(t, what):
    """There is no CPU fallback: ops that take raw strides / pre-allocated outputs check their tensors here."""
    if not t.is_cuda:
        raise RuntimeError(f'{what} (no CPU fallback)')


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be contiguous CUDA memory."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('msmdfusion_b200 ops need CUDA tensors (there is no CPU fallback)')
    if not t.is_contiguous():
        raise RuntimeError('msmdfusion_b200 ops need contiguous tensors')
    return ctypes.c_void_p(t.data_ptr())


def _dev_index(device):
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    idx = torch.device(device).index
    return torch.cuda.current_device() if idx is None else idx


def raw_stream(device=None):
    """cudaStream_t of torch's current stream as an int (the private raw getter is ~10x cheaper than
    building a torch.cuda.Stream object; it is called once per C-ABI call)."""
    return torch._C._cuda_getCurrentRawStream(_dev_index(device))


def stream(device=None):
    return ctypes.c_void_p(raw_stream(device))


_INTS, _FLOATS = {}, {}


def ints(vals):
    """Host int array for a geometry tuple (cached: the same few tuples recur every call)."""
    key = tuple(int(v) for v in vals)
    a = _INTS.get(key)
    if a is None:
        a = _INTS[key] = (ctypes.c_int * len(key))(*key)
    return a


def floats(vals):
    key = tuple(float(v) for v in vals)
    a = _FLOATS.get(key)
    if a is None:
        a = _FLOATS[key] = (ctypes.c_float * len(key))(*key)
    return a


class _Scratch:
    """Per-device grow-only scratch buffers (stream-ordered reuse on the current stream)."""

    def __init__(self):
        self.buf = {}

    def get(self, device, nbytes, slot='ws'):
        # one buffer per (device, stream): reuse is stream-ordered, and concurrent streams (the
        # overlapped NN-assignment chains of the GMA encoder) never share scratch memory
        key = (_dev_index(device), slot, raw_stream(device))
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


scratch = _Scratch()
