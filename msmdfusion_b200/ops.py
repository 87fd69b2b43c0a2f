"""Tensor-level wrappers over the C ABI (include/msmd_b200.h).

Every function takes/returns CUDA torch tensors, enqueues on the current stream and maps
1:1 onto an ``msmd_*`` entry point.  Nothing here computes on the CPU.
"""
import torch

from . import _cabi
from ._cabi import check, floats, ints, lib, ptr, scratch, stream


# bench.py sets PROFILE to a list to time every C-ABI call with CUDA events on the launching
# stream (per-kernel roofline accounting); None = no instrumentation.
PROFILE = None
PROFILE_SYNC = False  # debugging: synchronise before each timed call


class _Timed:
    """Context manager: brackets one C-ABI call with CUDA events when PROFILE is a list."""

    def __init__(self, op, replay=None, **meta):
        self.rec = None
        if PROFILE is not None:
            self.rec = dict(op=op, **meta)
            if replay is not None:   # re-issues the identical C-ABI call (same operands): bench.py times kernels with it
                self.rec['replay'] = replay

    def __enter__(self):
        if self.rec is not None:
            if PROFILE_SYNC:
                torch.cuda.synchronize()
            self.rec['start'] = torch.cuda.Event(enable_timing=True)
            self.rec['end'] = torch.cuda.Event(enable_timing=True)
            self.rec['start'].record()
        return self

    def __exit__(self, *exc):
        if self.rec is not None:
            self.rec['end'].record()
            PROFILE.append(self.rec)
        return False


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3, v
        return [int(x) for x in v]
    return [int(v)] * 3


# --------------------------------------------------------------------------------------
# hard_voxelize (+ fused HardSimpleVFE mean)
# reference: mmdet3d/ops/voxel/voxelize.py:13-59, src/voxelization_cpu.cpp:43-142
# --------------------------------------------------------------------------------------
def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels, want_voxels=True,
                  mean_features=0, batch_idx=None, out=None):
    """Returns (voxels|None, coors, num_points_per_voxel, mean|None), sliced to voxel_num.

    coors is (V,3) (z,y,x) or, when ``batch_idx`` is given, (V,4) (batch_idx,z,y,x).
    One host read-back (voxel_num) at the end -- the reference does the same
    (voxelization_cuda.cu:322-323) after four device synchronisations.
    ``out``: the caller's pre-allocated (voxels, coors, num_points_per_voxel) of the reference's calling convention
    (voxelize.py:44-52), filled in place when they are contiguous CUDA tensors of the right type and capacity.
    """
    if points.dtype != torch.float32:
        points = points.float()
    points = points.contiguous()
    n, c = points.shape
    dev = points.device
    max_voxels_eff = max_voxels if max_voxels >= 0 else n
    cap = max(1, min(n, max_voxels_eff))
    ncol = 3 if batch_idx is None else 4
    direct = False
    if out is not None and want_voxels and batch_idx is None and not mean_features:
        ov, oc, on = out
        direct = (all(t.is_cuda and t.is_contiguous() and t.device == dev for t in out) and
                  ov.dtype == torch.float32 and oc.dtype == torch.int32 and on.dtype == torch.int32 and
                  tuple(ov.shape[1:]) == (max_points, c) and oc.dim() == 2 and oc.shape[1] == 3 and
                  min(ov.shape[0], oc.shape[0], on.shape[0]) >= cap)
    if direct:
        voxels, coors, num = out
    else:
        voxels = torch.empty((cap, max_points, c), dtype=torch.float32, device=dev) if want_voxels else None
        coors = torch.empty((cap, ncol), dtype=torch.int32, device=dev)
        num = torch.empty((cap,), dtype=torch.int32, device=dev)
    mean = (torch.empty((cap, mean_features), dtype=torch.float32, device=dev)
            if mean_features else None)
    voxel_num = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws_bytes = lib().msmd_hard_voxelize_workspace(n)
    ws = scratch.get(dev, ws_bytes)
    with _Timed('hard_voxelize', n=n, c=c, mean_features=int(mean_features), voxels=bool(want_voxels)):
        check(lib().msmd_hard_voxelize(ptr(points), n, c, floats(voxel_size), floats(coors_range),
                                       int(max_points), int(max_voxels), ptr(voxels), ptr(coors), ncol,
                                       0 if batch_idx is None else int(batch_idx), ptr(num), ptr(mean),
                                       int(mean_features), ptr(voxel_num), ptr(ws), ws.numel(),
                                       stream(dev)), 'msmd_hard_voxelize')
    v = int(voxel_num.item())
    return (voxels[:v] if want_voxels else None, coors[:v], num[:v],
            mean[:v] if mean_features else None)


# --------------------------------------------------------------------------------------
# occupancy bit grid + rulebooks
# reference: ops.get_indice_pairs_implicit_gemm, call site bug_fix/conv.py:382-415
# --------------------------------------------------------------------------------------
class BitGrid:
    """bits/prefix(/perm) of one active-voxel set; see include/msmd_b200.h."""

    __slots__ = ('bits', 'prefix', 'perm', 'spatial_shape', 'batch_size', 'num_active')

    def __init__(self, bits, prefix, perm, spatial_shape, batch_size, num_active=None):
        self.bits, self.prefix, self.perm = bits, prefix, perm
        self.spatial_shape, self.batch_size = list(spatial_shape), int(batch_size)
        self.num_active = num_active


def _indices_ok(indices):
    assert indices.dtype == torch.int32 and indices.dim() == 2 and indices.shape[1] == 4, \
        'indices must be (N,4) int32 (batch,z,y,x)'
    return indices.contiguous()


def grid_build(indices, batch_size, spatial_shape, need_perm=True):
    indices = _indices_ok(indices)
    dev = indices.device
    shape = _triple(spatial_shape)
    words = lib().msmd_grid_num_words(int(batch_size), ints(shape))
    bits = torch.empty((words,), dtype=torch.int32, device=dev)
    prefix = torch.empty((words,), dtype=torch.int32, device=dev)
    n = indices.shape[0]
    perm = torch.empty((max(n, 1),), dtype=torch.int32, device=dev) if need_perm else None
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = scratch.get(dev, lib().msmd_scan_workspace())
    with _Timed('grid_build', n=n, words=int(words)):
        check(lib().msmd_grid_build(ptr(indices), n, int(batch_size), ints(shape), ptr(bits),
                                    ptr(prefix), ptr(perm), ptr(count), ptr(ws), ws.numel(),
                                    stream(dev)), 'msmd_grid_build')
    return BitGrid(bits, prefix, perm, shape, batch_size, count)


def rulebook_subm(indices, grid, ksize, dilation=1):
    indices = _indices_ok(indices)
    ks, dl = _triple(ksize), _triple(dilation)
    n = indices.shape[0]
    kvol = ks[0] * ks[1] * ks[2]
    pair = torch.empty((kvol, n), dtype=torch.int32, device=indices.device)
    with _Timed('rulebook_subm', n=n, kvol=kvol):
        check(lib().msmd_rulebook_subm(ptr(indices), n, grid.batch_size, ints(grid.spatial_shape),
                                       ints(ks), ints(dl), ptr(grid.bits), ptr(grid.prefix),
                                       ptr(grid.perm), ptr(pair), stream(indices.device)),
              'msmd_rulebook_subm')
    return pair


def conv_out_shape(spatial_shape, ksize, stride, padding, dilation):
    out = (_cabi.ctypes.c_int * 3)()  # output buffer: never the cached constant arrays of ints()
    check(lib().msmd_conv_out_shape(ints(_triple(spatial_shape)), ints(_triple(ksize)),
                                    ints(_triple(stride)), ints(_triple(padding)),
                                    ints(_triple(dilation)), out), 'msmd_conv_out_shape')
    return [int(x) for x in out]


def rulebook_conv(indices, grid, ksize, stride, padding, dilation=1):
    """Returns (out_indices (N_out,4) ascending linear order, pair_fwd (K,N_out), out_grid)."""
    indices = _indices_ok(indices)
    dev = indices.device
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    shape = grid.spatial_shape
    out_shape = conv_out_shape(shape, ks, st, pd, dl)
    B = grid.batch_size
    words = lib().msmd_grid_num_words(B, ints(out_shape))
    out_bits = torch.empty((words,), dtype=torch.int32, device=dev)
    out_prefix = torch.empty((words,), dtype=torch.int32, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = scratch.get(dev, lib().msmd_scan_workspace())
    n = indices.shape[0]
    with _Timed('rulebook_conv_outputs', n=n, words=int(words)):
        check(lib().msmd_rulebook_conv_outputs(ptr(indices), n, B, ints(shape), ints(ks), ints(st),
                                               ints(pd), ints(dl), ptr(out_bits), ptr(out_prefix),
                                               ptr(count), ptr(ws), ws.numel(), stream(dev)),
              'msmd_rulebook_conv_outputs')
    n_out = int(count.item())  # host needs N_out to size the outputs
    kvol = ks[0] * ks[1] * ks[2]
    out_indices = torch.empty((n_out, 4), dtype=torch.int32, device=dev)
    pair = torch.empty((kvol, n_out), dtype=torch.int32, device=dev)
    with _Timed('rulebook_conv_pairs', n=n_out, kvol=kvol):
        check(lib().msmd_rulebook_conv_pairs(ptr(out_bits), ptr(out_prefix), n_out, B, ints(shape),
                                             ints(ks), ints(st), ints(pd), ints(dl), ptr(grid.bits),
                                             ptr(grid.prefix), ptr(grid.perm), ptr(out_indices),
                                             ptr(pair), stream(dev)), 'msmd_rulebook_conv_pairs')
    out_grid = BitGrid(out_bits, out_prefix, None, out_shape, B, count)
    return out_indices, pair, out_grid


# --------------------------------------------------------------------------------------
# sparse conv forward, dense()
# reference: Fsp.implicit_gemm call site bug_fix/conv.py:442-447
# --------------------------------------------------------------------------------------
def pack_weight(weight):
    """KRSC [Cout,kz,ky,kx,Cin] parameter -> kernel layout [K, Cin, Cout]."""
    w = weight.detach()
    if w.dtype != torch.float32:
        w = w.float()
    w = w.contiguous()
    cout, cin = w.shape[0], w.shape[-1]
    kvol = w.numel() // (cout * cin)
    packed = torch.empty((kvol, cin, cout), dtype=torch.float32, device=w.device)
    check(lib().msmd_spconv_pack_weight(ptr(w), cout, kvol, cin, ptr(packed), stream(w.device)),
          'msmd_spconv_pack_weight')
    return packed


class TcWeight:
    """Weight packed for the tensor-core path.  ``mode`` = the value of ``msmd_conv_layer.weight_tc``:
    1 = tf32 hi/lo image (csrc/spconv_tc.cu, 3xTF32), 2 = bf16 hi/lo image (csrc/spconv_tc16.cu, bf16x3),
    3 = bf16 image (one MMA per product: the train-step arithmetic of BASELINE configs[4]),
    4 = bf16 hi/lo image with the input channels padded to 8 (csrc/spconv_sb.cu: bf16x3 through the split-bf16
    operand cache -- same arithmetic as mode 2, the activations' hi/lo split done by their producer)."""

    __slots__ = ('packed', 'cout', 'kvol', 'cin', 'mode')

    def __init__(self, packed, cout, kvol, cin, mode=1):
        self.packed, self.cout, self.kvol, self.cin, self.mode = packed, cout, kvol, cin, int(mode)


TC_MODES = {'tf32x3': 1, 'bf16x3': 2, 'bf16': 3, 'bf16x3c': 4}


def set_mask_sort(enable):
    """Opt-in: the native executor mask-sorts its 3x3x3 SubM rulebooks (``msmd_spconv_set_mask_sort``)."""
    check(lib().msmd_spconv_set_mask_sort(int(bool(enable))), 'msmd_spconv_set_mask_sort')


def set_wgrad_tc(enable):
    """Opt-in: ``spconv_bwd_weight`` runs the tensor-core kernel (csrc/spconv_wgrad_tc.cu, 3xTF32) for the
    shapes it supports instead of the exact-fp32 FFMA kernel (``msmd_spconv_set_wgrad_tc``)."""
    check(lib().msmd_spconv_set_wgrad_tc(int(bool(enable))), 'msmd_spconv_set_wgrad_tc')


def set_tc16_variant(variant):
    """16-bit operand modes: 2 (default) = A operand in shared memory, 3 = A operand in tensor memory."""
    check(lib().msmd_spconv_tc16_set_variant(int(variant)), 'msmd_spconv_tc16_set_variant')


def set_tc_variant(variant):
    """0 (default): chosen by Cout; 3: A operand staged in tensor memory; 2: A operand in shared memory."""
    check(lib().msmd_spconv_tc_set_variant(int(variant)), 'msmd_spconv_tc_set_variant')


def tc_supported(cout, kvol, cin):
    return bool(lib().msmd_spconv_tc_supported(int(cout), int(kvol), int(cin)))


def pack_weight_tc(weight, mode=1):
    """KRSC [Cout,kz,ky,kx,Cin] parameter -> swizzled K-chunk image of the tensor-core kernels: tf32 hi/lo
    (mode 1), bf16 hi/lo (mode 2) or bf16 (mode 3)."""
    w = weight.detach()
    if w.dtype != torch.float32:
        w = w.float()
    w = w.contiguous()
    cout, cin = w.shape[0], w.shape[-1]
    kvol = w.numel() // (cout * cin)
    mode = int(mode)
    assert mode in (1, 2, 3, 4)
    if mode == 4:
        n = lib().msmd_spconv_sb_packed_bytes(cout, kvol, cin)
        assert n > 0, 'shape not supported by the tensor-core path'
        packed = torch.empty((n // 2,), dtype=torch.int16, device=w.device)   # raw bf16 bit patterns
        check(lib().msmd_spconv_sb_pack_weight(ptr(w), cout, kvol, cin, ptr(packed), stream(w.device)),
              'msmd_spconv_sb_pack_weight')
    elif mode == 1:
        n = lib().msmd_spconv_tc_packed_floats(cout, kvol, cin)
        assert n > 0, 'shape not supported by the tensor-core path'
        packed = torch.empty((n,), dtype=torch.float32, device=w.device)
        check(lib().msmd_spconv_tc_pack_weight(ptr(w), cout, kvol, cin, ptr(packed), stream(w.device)),
              'msmd_spconv_tc_pack_weight')
    else:
        x3 = int(mode == 2)
        n = lib().msmd_spconv_tc16_packed_bytes(cout, kvol, cin, x3)
        assert n > 0, 'shape not supported by the tensor-core path'
        packed = torch.empty((n // 2,), dtype=torch.int16, device=w.device)   # raw bf16 bit patterns
        check(lib().msmd_spconv_tc16_pack_weight(ptr(w), cout, kvol, cin, x3, ptr(packed), stream(w.device)),
              'msmd_spconv_tc16_pack_weight')
    return TcWeight(packed, cout, kvol, cin, mode)


def rulebook_mask_sort(pair_fwd):
    """Mask-sorted tiles (spconv-2.x ``mask_argsort_fwd_splits``): (row_perm (n) i32, pair_sorted (27,n))
    with pair_sorted = pair_fwd[:, row_perm], rows grouped by a 15-bit digest of their neighbour mask."""
    pair_fwd = pair_fwd.contiguous()
    kvol, n = pair_fwd.shape
    assert kvol == 27 and pair_fwd.dtype == torch.int32
    dev = pair_fwd.device
    row_perm = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)[:n]
    pair_sorted = torch.empty_like(pair_fwd)
    need = lib().msmd_rulebook_mask_sort_workspace(n)
    ws = scratch.get(dev, need, slot='mask_sort')
    with _Timed('rulebook_mask_sort', n=n, kvol=kvol):
        check(lib().msmd_rulebook_mask_sort(ptr(pair_fwd), kvol, n, ptr(row_perm), ptr(pair_sorted), ptr(ws),
                                            ws.numel(), stream(dev)), 'msmd_rulebook_mask_sort')
    return row_perm, pair_sorted


def split_bf16(features):
    """(n, C) fp32 rows -> split image (n, 2*round_up(C, 8)) of bf16 bit patterns [hi | lo] (csrc/spconv_sb.cu).
    The image is remembered on the tensor object (keyed by its version), so a tensor that feeds several
    convolutions -- or came out of one that already wrote its image -- is split once."""
    cached = getattr(features, '_msmd_split', None)
    if cached is not None and cached[0] == (features.data_ptr(), features._version, tuple(features.shape)):
        return cached[1]
    f = features.contiguous()
    if f.dtype != torch.float32:
        f = f.float()
    n, c = f.shape
    xs = torch.empty((n, lib().msmd_split_width(c)), dtype=torch.int16, device=f.device)
    with _Timed('split_bf16', n=n, c=c):
        check(lib().msmd_split_bf16(ptr(f), n, c, ptr(xs), stream(f.device)), 'msmd_split_bf16')
    _remember_split(features, xs)
    return xs


def _remember_split(t, xs):
    try:
        t._msmd_split = ((t.data_ptr(), t._version, tuple(t.shape)), xs)
    except (AttributeError, RuntimeError):   # tensors that refuse attributes: just do not cache
        pass


def rulebook_tile_masks(pair_fwd):
    """``msmd_rulebook_tile_masks``: one uint32 per 128-row tile of the pair table, bit k = the tile has a pair at
    kernel offset k.  Built once per table and kept on the tensor (rulebooks are never written in place)."""
    cached = pair_fwd.__dict__.get('_msmd_tile_mask')
    if cached is not None:
        return cached
    kvol, n_out = pair_fwd.shape
    tm = torch.empty(((n_out + 127) // 128,), dtype=torch.int32, device=pair_fwd.device)
    if n_out:
        with _Timed('rulebook_tile_masks', n=n_out, kvol=kvol):
            check(lib().msmd_rulebook_tile_masks(ptr(pair_fwd), kvol, n_out, ptr(tm), stream(pair_fwd.device)),
                  'msmd_rulebook_tile_masks')
    pair_fwd._msmd_tile_mask = tm
    return tm


def spconv_fwd_sb(features, tcw, pair_fwd, scale=None, shift=None, residual=None, relu=False, want_split=True,
                  row_perm=None):
    """bf16x3 sparse convolution through the split-bf16 operand cache (``msmd_spconv_fwd_sb_ex``): the gather reads
    the split image of ``features``; the epilogue writes the fp32 result and (``want_split``) its split image, which
    is attached to the returned tensor for the next convolution.  ``row_perm``: the table is a mask-sorted one."""
    assert tcw.mode == 4 and features.shape[1] == tcw.cin, 'channel size mismatch'
    assert pair_fwd.shape[0] == tcw.kvol and pair_fwd.dtype == torch.int32
    n_out = pair_fwd.shape[1]
    dev = features.device
    out = torch.empty((n_out, tcw.cout), dtype=torch.float32, device=dev)
    if n_out == 0:
        return out
    if residual is not None:
        residual = residual.contiguous()
        assert residual.shape == out.shape
    if not pair_fwd.is_contiguous():
        pair_fwd = pair_fwd.contiguous()
    xs = split_bf16(features)
    out_s = torch.empty((n_out, lib().msmd_split_width(tcw.cout)), dtype=torch.int16, device=dev) \
        if want_split else None
    tm = rulebook_tile_masks(pair_fwd) if lib().msmd_spconv_sb_uses_tile_masks() else None
    def launch():
        check(lib().msmd_spconv_fwd_sb_ex(ptr(xs), features.shape[0], ptr(tcw.packed), ptr(pair_fwd), ptr(row_perm),
                                          ptr(tm), n_out, tcw.cin, tcw.cout, tcw.kvol, ptr(scale), ptr(shift),
                                          ptr(residual), int(bool(relu)), ptr(out), ptr(out_s), stream(dev)),
              'msmd_spconv_fwd_sb_ex')
    with _Timed('spconv_fwd', replay=launch, n_in=features.shape[0], n_out=n_out, cin=tcw.cin, cout=tcw.cout,
                kvol=tcw.kvol, residual=residual is not None, pair=pair_fwd, path='tc', tc_mode=tcw.mode):
        launch()
    if out_s is not None:
        _remember_split(out, out_s)
    return out


def spconv_fwd_tc(features, tcw, pair_fwd, scale=None, shift=None, residual=None, relu=False, row_perm=None):
    """Sparse conv forward on tcgen05 tensor cores (3xTF32 / bf16x3 / bf16 by ``tcw.mode``, fp32
    accumulate in TMEM).  With
    ``row_perm`` the table is a mask-sorted one (``rulebook_mask_sort``); the output keeps the original
    row order."""
    if tcw.mode == 4:   # split-bf16 operand cache
        return spconv_fwd_sb(features, tcw, pair_fwd, scale, shift, residual, relu, row_perm=row_perm)
    features = features.contiguous()
    if features.dtype != torch.float32:
        features = features.float()
    assert features.shape[1] == tcw.cin, 'channel size mismatch'
    assert pair_fwd.shape[0] == tcw.kvol and pair_fwd.dtype == torch.int32
    n_out = pair_fwd.shape[1]
    out = torch.empty((n_out, tcw.cout), dtype=torch.float32, device=features.device)
    if residual is not None:
        residual = residual.contiguous()
        assert residual.shape == out.shape
    pair_fwd = pair_fwd.contiguous()
    with _Timed('spconv_fwd', n_in=features.shape[0], n_out=n_out, cin=tcw.cin, cout=tcw.cout,
                kvol=tcw.kvol, residual=residual is not None, pair=pair_fwd, path='tc', tc_mode=tcw.mode):
        if row_perm is not None:
            assert row_perm.dtype == torch.int32 and row_perm.shape[0] == n_out
            row_perm = row_perm.contiguous()
        if tcw.mode != 1:   # 16-bit operand kernels (bf16x3 / bf16)
            need = lib().msmd_spconv_tc16_workspace(n_out, tcw.cout)   # > 0: variant 3 with split-K pairs
            ws = scratch.get(features.device, need, slot='tc_ws') if need else None
            check(lib().msmd_spconv_fwd_tc16_ws(ptr(features), features.shape[0], ptr(tcw.packed), ptr(pair_fwd),
                                                ptr(row_perm), n_out, tcw.cin, tcw.cout, tcw.kvol,
                                                int(tcw.mode == 2), ptr(scale), ptr(shift), ptr(residual),
                                                int(bool(relu)), ptr(out), ptr(ws),
                                                ws.numel() if ws is not None else 0, stream(features.device)),
                  'msmd_spconv_fwd_tc16_ws')
            return out
        need = lib().msmd_spconv_tc_workspace(n_out, tcw.cout)  # > 0: split-K pairs (tail balance)
        ws = scratch.get(features.device, need, slot='tc_ws') if need else None
        if row_perm is not None:
            check(lib().msmd_spconv_fwd_tc_sorted(ptr(features), features.shape[0], ptr(tcw.packed),
                                                  ptr(pair_fwd), ptr(row_perm.contiguous()), n_out, tcw.cin,
                                                  tcw.cout, tcw.kvol, ptr(scale), ptr(shift), ptr(residual),
                                                  int(bool(relu)), ptr(out), ptr(ws),
                                                  ws.numel() if ws is not None else 0,
                                                  stream(features.device)), 'msmd_spconv_fwd_tc_sorted')
            return out
        check(lib().msmd_spconv_fwd_tc_ws(ptr(features), features.shape[0], ptr(tcw.packed), ptr(pair_fwd),
                                          n_out, tcw.cin, tcw.cout, tcw.kvol, ptr(scale), ptr(shift),
                                          ptr(residual), int(bool(relu)), ptr(out), ptr(ws),
                                          ws.numel() if ws is not None else 0,
                                          stream(features.device)), 'msmd_spconv_fwd_tc_ws')
    return out


def spconv_fwd(features, packed_weight, pair_fwd, scale=None, shift=None, residual=None, relu=False):
    if isinstance(packed_weight, TcWeight):
        return spconv_fwd_tc(features, packed_weight, pair_fwd, scale, shift, residual, relu)
    features = features.contiguous()
    if features.dtype != torch.float32:
        features = features.float()
    kvol, cin, cout = packed_weight.shape
    assert features.shape[1] == cin, 'channel size mismatch'
    assert pair_fwd.shape[0] == kvol and pair_fwd.dtype == torch.int32
    n_out = pair_fwd.shape[1]
    out = torch.empty((n_out, cout), dtype=torch.float32, device=features.device)
    if residual is not None:
        residual = residual.contiguous()
        assert residual.shape == out.shape
    pair_fwd = pair_fwd.contiguous()
    with _Timed('spconv_fwd', n_in=features.shape[0], n_out=n_out, cin=cin, cout=cout, kvol=kvol,
                residual=residual is not None, pair=pair_fwd, path='simt'):
        check(lib().msmd_spconv_fwd(ptr(features), features.shape[0], ptr(packed_weight),
                                    ptr(pair_fwd), n_out, cin, cout, kvol, ptr(scale), ptr(shift),
                                    ptr(residual), int(bool(relu)), ptr(out),
                                    stream(features.device)), 'msmd_spconv_fwd')
    return out


def to_dense(indices, features, spatial_shape, batch_size):
    indices = _indices_ok(indices)
    features = features.contiguous().float()
    n, c = features.shape
    d, h, w = _triple(spatial_shape)
    out = torch.empty((int(batch_size), c, d, h, w), dtype=torch.float32, device=features.device)
    with _Timed('to_dense', n=n, c=c, cells=int(batch_size) * d * h * w):
        check(lib().msmd_to_dense(ptr(indices), ptr(features), n, c, int(batch_size), ints([d, h, w]),
                                  ptr(out), stream(features.device)), 'msmd_to_dense')
    return out


# --------------------------------------------------------------------------------------
# sparse conv backward (config 5: the train step), dense() / sparse_add backward helpers
# reference: Fsp.implicit_gemm backward through pair_bwd (bug_fix/conv.py:382-415,442-447);
# arithmetic of mmdet3d/ops/spconv/include/spconv/spconv_ops.h:364-457
# --------------------------------------------------------------------------------------
def rulebook_transpose(pair_fwd, n_in):
    """pair_fwd (K,N_out) -> pair_bwd (K,N_in): the output row reading input row i through offset k."""
    pair_fwd = pair_fwd.contiguous()
    assert pair_fwd.dtype == torch.int32 and pair_fwd.dim() == 2
    kvol, n_out = pair_fwd.shape
    pair_bwd = torch.empty((kvol, int(n_in)), dtype=torch.int32, device=pair_fwd.device)
    with _Timed('rulebook_transpose', n_in=int(n_in), n_out=n_out, kvol=kvol):
        check(lib().msmd_rulebook_transpose(ptr(pair_fwd), kvol, n_out, int(n_in), ptr(pair_bwd),
                                            stream(pair_fwd.device)), 'msmd_rulebook_transpose')
    return pair_bwd


def transpose_weight(weight, flip_k=False):
    """KRSC [Cout,kz,ky,kx,Cin] -> [Cin,kz,ky,kx,Cout] (kernel offsets reversed when ``flip_k``):
    the weight whose FORWARD contraction over pair_bwd is the data gradient."""
    w = weight.detach()
    if w.dtype != torch.float32:
        w = w.float()
    w = w.contiguous()
    cout, cin = w.shape[0], w.shape[-1]
    kvol = w.numel() // (cout * cin)
    wt = torch.empty((cin, *w.shape[1:-1], cout), dtype=torch.float32, device=w.device)
    check(lib().msmd_spconv_transpose_weight(ptr(w), cout, kvol, cin, int(bool(flip_k)), ptr(wt),
                                             stream(w.device)), 'msmd_spconv_transpose_weight')
    return wt


def spconv_bwd_data(grad_out, packed_wt, pair_bwd):
    """grad_in (N_in,Cin): forward contraction of grad_out (N_out,Cout) over pair_bwd (K,N_in) with the
    packed TRANSPOSED weight (``pack_weight`` / ``pack_weight_tc`` of ``transpose_weight``)."""
    grad_out = grad_out.contiguous()
    if grad_out.dtype != torch.float32:
        grad_out = grad_out.float()
    pair_bwd = pair_bwd.contiguous()
    n_in = pair_bwd.shape[1]
    if isinstance(packed_wt, TcWeight):
        cin, cout, kvol, tc, wbuf = packed_wt.cout, packed_wt.cin, packed_wt.kvol, packed_wt.mode, packed_wt.packed
    else:
        kvol, cout, cin = packed_wt.shape
        tc, wbuf = 0, packed_wt
    assert grad_out.shape[1] == cout and pair_bwd.shape[0] == kvol and pair_bwd.dtype == torch.int32
    if tc == 4:   # the data gradient IS the forward contraction: split dY once, run the split-operand kernel
        return spconv_fwd_sb(grad_out, packed_wt, pair_bwd, want_split=False)
    grad_in = torch.empty((n_in, cin), dtype=torch.float32, device=grad_out.device)
    with _Timed('spconv_bwd_data', n_in=n_in, n_out=grad_out.shape[0], cin=cin, cout=cout, kvol=kvol,
                pair=pair_bwd, path='tc' if tc else 'simt'):
        need = lib().msmd_spconv_tc_workspace(n_in, cin) if tc == 1 else 0
        ws = scratch.get(grad_out.device, need, slot='tc_ws') if need else None
        check(lib().msmd_spconv_bwd_data(ptr(grad_out), grad_out.shape[0], ptr(wbuf), tc, ptr(pair_bwd),
                                         n_in, cin, cout, kvol, ptr(grad_in), ptr(ws),
                                         ws.numel() if ws is not None else 0, stream(grad_out.device)),
              'msmd_spconv_bwd_data')
    return grad_in


def spconv_bwd_weight(features, grad_out, pair_fwd, weight_shape):
    """grad_weight in the parameter's KRSC shape; deterministic, exact fp32."""
    features, grad_out, pair_fwd = features.contiguous(), grad_out.contiguous(), pair_fwd.contiguous()
    if features.dtype != torch.float32:
        features = features.float()
    if grad_out.dtype != torch.float32:
        grad_out = grad_out.float()
    cout, cin = int(weight_shape[0]), int(weight_shape[-1])
    kvol, n_out = pair_fwd.shape
    assert features.shape[1] == cin and grad_out.shape == (n_out, cout) and pair_fwd.dtype == torch.int32
    grad_w = torch.empty(tuple(weight_shape), dtype=torch.float32, device=features.device)
    assert grad_w.numel() == cout * kvol * cin
    need = lib().msmd_spconv_bwd_weight_workspace(n_out, cin, cout, kvol)
    ws = scratch.get(features.device, need, slot='wgrad_ws')
    with _Timed('spconv_bwd_weight', n_in=features.shape[0], n_out=n_out, cin=cin, cout=cout, kvol=kvol,
                pair=pair_fwd):
        check(lib().msmd_spconv_bwd_weight(ptr(features), features.shape[0], ptr(grad_out), ptr(pair_fwd),
                                           n_out, cin, cout, kvol, ptr(grad_w), ptr(ws), ws.numel(),
                                           stream(features.device)), 'msmd_spconv_bwd_weight')
    return grad_w


def from_dense(indices, dense, spatial_shape, batch_size):
    """Backward of ``to_dense``: (B,C,D,H,W) -> the active rows (n,C)."""
    indices = _indices_ok(indices)
    dense = dense.contiguous().float()
    n, c = indices.shape[0], dense.shape[1]
    d, h, w = _triple(spatial_shape)
    assert tuple(dense.shape) == (int(batch_size), c, d, h, w)
    out = torch.empty((n, c), dtype=torch.float32, device=dense.device)
    with _Timed('from_dense', n=n, c=c):
        check(lib().msmd_from_dense(ptr(indices), ptr(dense), n, c, int(batch_size), ints([d, h, w]),
                                    ptr(out), stream(dense.device)), 'msmd_from_dense')
    return out


def grid_rows(indices, grid):
    """Row of each voxel in ``grid``'s ascending order (int64, -1 = absent)."""
    indices = _indices_ok(indices)
    n = indices.shape[0]
    rows = torch.empty((n,), dtype=torch.int32, device=indices.device)
    with _Timed('grid_rows', n=n):
        check(lib().msmd_grid_rows(ptr(indices), n, grid.batch_size, ints(grid.spatial_shape),
                                   ptr(grid.bits), ptr(grid.prefix), ptr(rows), stream(indices.device)),
              'msmd_grid_rows')
    return rows.long()


# --------------------------------------------------------------------------------------
# FPS / ball query / nearest 3-D voxel (fps_NN_fast, painting.py:276-323)
# --------------------------------------------------------------------------------------
def furthest_point_sample_single(xyz, m):
    """xyz (n,3) f32 -> (m,) i32; start index 0, reference tie-break."""
    xyz = xyz.contiguous().float()
    n = xyz.shape[0]
    idx = torch.empty((m,), dtype=torch.int32, device=xyz.device)
    ws = scratch.get(xyz.device, lib().msmd_fps_workspace(n))
    with _Timed('fps', n=n, m=int(m)):
        check(lib().msmd_fps(ptr(xyz), n, int(m), ptr(idx), ptr(ws), ws.numel(), stream(xyz.device)),
              'msmd_fps')
    return idx


def ball_query_single(min_radius, max_radius, nsample, xyz, center_xyz):
    """xyz (n,3), center_xyz (m,3) f32 -> (m,nsample) i32."""
    xyz, center_xyz = xyz.contiguous().float(), center_xyz.contiguous().float()
    n, mc = xyz.shape[0], center_xyz.shape[0]
    idx = torch.empty((mc, nsample), dtype=torch.int32, device=xyz.device)
    with _Timed('ball_query', n=n, m=mc, nsample=int(nsample)):
        check(lib().msmd_ball_query(ptr(xyz), n, ptr(center_xyz), mc, float(min_radius),
                                    float(max_radius), int(nsample), ptr(idx), stream(xyz.device)),
              'msmd_ball_query')
    return idx


def furthest_point_sample(points_xyz, num_points):
    """Drop-in for ``mmdet3d.ops.furthest_point_sample`` (furthest_point_sample/furthest_point_sample.py:8-37):
    points_xyz (B, N, 3) f32 contiguous -> (B, num_points) int32, start index 0 per batch element."""
    assert points_xyz.dim() == 3 and points_xyz.shape[2] == 3 and points_xyz.is_contiguous()
    return torch.stack([furthest_point_sample_single(points_xyz[b], num_points) for b in range(points_xyz.shape[0])])


def ball_query(min_radius, max_radius, sample_num, xyz, center_xyz):
    """Drop-in for ``mmdet3d.ops.ball_query`` (ball_query/ball_query.py:8-46): xyz (B, N, 3), center_xyz
    (B, npoint, 3) f32 contiguous -> (B, npoint, sample_num) int32."""
    assert center_xyz.is_contiguous() and xyz.is_contiguous() and min_radius < max_radius
    return torch.stack([ball_query_single(min_radius, max_radius, sample_num, xyz[b], center_xyz[b])
                        for b in range(xyz.shape[0])])


def nn_search(query_zyx, key_zyx):
    """(nq,3)/(nk,3) int32 voxel coordinates -> (val f32 (nq), idx i32 (nq))."""
    q, k = query_zyx.contiguous(), key_zyx.contiguous()
    assert q.dtype == torch.int32 and k.dtype == torch.int32
    nq, nk = q.shape[0], k.shape[0]
    val = torch.empty((nq,), dtype=torch.float32, device=q.device)
    idx = torch.empty((nq,), dtype=torch.int32, device=q.device)
    with _Timed('nn_search', nq=nq, nk=nk):
        check(lib().msmd_nn_search(ptr(q), q.shape[1], nq, ptr(k), k.shape[1], nk, ptr(val), ptr(idx),
                                   stream(q.device)), 'msmd_nn_search')
    return val, idx


def group_assign(group, val, nn_idx, dist_thresh, nq, base=0):
    """query_NN_key_idx (nq,) int64: nearest key (+base) of each query's representative or -1."""
    dev = val.device
    out = torch.empty((nq,), dtype=torch.int64, device=dev)
    if group is None:
        m = nsample = 0
        winner = None
    else:
        group = group.contiguous()
        m, nsample = group.shape
        winner = torch.empty((max(nq, 1),), dtype=torch.int32, device=dev)
    with _Timed('group_assign', nq=nq, m=m):
        check(lib().msmd_group_assign(ptr(group), m, nsample, ptr(val), ptr(nn_idx), float(dist_thresh),
                                      nq, int(base), ptr(winner), ptr(out), stream(dev)),
              'msmd_group_assign')
    return out


# --------------------------------------------------------------------------------------
# voxel_modality_split (one sample), sparse_add, lift gather
# --------------------------------------------------------------------------------------
FORCE_SPLIT_SORT = False   # tests: run the sort path of voxel_modality_split as well


def modality_split_single(coord3, coord2, offset3=0, offset2=0):
    """coord3 (n3,4) / coord2 (n2,4) int32 rows of one sample ->
    (mix3 (n3) i32, mix2 (n2) i32, syn3 (P) i64, syn2 (P) i64)."""
    coord3, coord2 = _indices_ok(coord3), _indices_ok(coord2)
    dev = coord3.device
    n3, n2 = coord3.shape[0], coord2.shape[0]
    mix3 = torch.empty((n3,), dtype=torch.int32, device=dev)
    mix2 = torch.empty((n2,), dtype=torch.int32, device=dev)
    cap = max(1, min(n3, n2))
    syn3 = torch.empty((cap,), dtype=torch.int64, device=dev)
    syn2 = torch.empty((cap,), dtype=torch.int64, device=dev)
    count = torch.empty((2,), dtype=torch.int32, device=dev)   # [pairs, overflow]: zeroed by the call
    ws = scratch.get(dev, lib().msmd_modality_split_workspace(n3, n2))
    args = (ptr(coord3), n3, ptr(coord2), n2, int(offset3), int(offset2), ptr(mix3), ptr(mix2), ptr(syn3), ptr(syn2),
            ptr(count), ptr(ws), ws.numel(), stream(dev))
    with _Timed('modality_split', n3=n3, n2=n2):
        check(lib().msmd_modality_split(*args), 'msmd_modality_split')
    p, overflow = count.tolist()
    if overflow or FORCE_SPLIT_SORT:   # a key run longer than 8 rows / more than 8192 pairs: the general (sort) path
        with _Timed('modality_split_sort', n3=n3, n2=n2):
            check(lib().msmd_modality_split_sort(*args), 'msmd_modality_split_sort')
        p = int(count[0].item())
    return mix3, mix2, syn3[:p], syn2[:p]


def compact_unflagged(flags, count):
    """Rows with flags == 0, ascending, as int64 (count known on the host -> no synchronisation)."""
    flags = flags.contiguous()
    assert flags.dtype == torch.int32
    n = flags.shape[0]
    out = torch.empty((max(n, 1),), dtype=torch.int64, device=flags.device)
    cnt = torch.empty((1,), dtype=torch.int32, device=flags.device)
    ws = scratch.get(flags.device, lib().msmd_scan_workspace())
    with _Timed('compact_unflagged', n=n):
        check(lib().msmd_compact_unflagged(ptr(flags), n, ptr(out), ptr(cnt), ptr(ws), ws.numel(),
                                           stream(flags.device)), 'msmd_compact_unflagged')
    return out[:int(count)]


def sparse_add(idx_a, feat_a, idx_b, feat_b, spatial_shape, batch_size):
    """Returns (out_indices ascending, out_features, BitGrid of the union)."""
    idx_a, idx_b = _indices_ok(idx_a), _indices_ok(idx_b)
    feat_a, feat_b = feat_a.contiguous().float(), feat_b.contiguous().float()
    assert feat_a.shape[1] == feat_b.shape[1]
    dev = idx_a.device
    shape = _triple(spatial_shape)
    B = int(batch_size)
    words = lib().msmd_grid_num_words(B, ints(shape))
    bits = torch.empty((words,), dtype=torch.int32, device=dev)
    prefix = torch.empty((words,), dtype=torch.int32, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = scratch.get(dev, lib().msmd_scan_workspace())
    na, nb, c = idx_a.shape[0], idx_b.shape[0], feat_a.shape[1]
    with _Timed('sparse_add_outputs', na=na, nb=nb, words=int(words)):
        check(lib().msmd_sparse_add_outputs(ptr(idx_a), na, ptr(idx_b), nb, B, ints(shape), ptr(bits),
                                            ptr(prefix), ptr(count), ptr(ws), ws.numel(), stream(dev)),
              'msmd_sparse_add_outputs')
    n_out = int(count.item())
    out_idx = torch.empty((n_out, 4), dtype=torch.int32, device=dev)
    out_feat = torch.empty((n_out, c), dtype=torch.float32, device=dev)
    with _Timed('sparse_add_finish', n_out=n_out, c=c):
        check(lib().msmd_sparse_add_finish(ptr(bits), ptr(prefix), n_out, ptr(idx_a), ptr(feat_a), na,
                                           ptr(idx_b), ptr(feat_b), nb, c, B, ints(shape), ptr(out_idx),
                                           ptr(out_feat), stream(dev)), 'msmd_sparse_add_finish')
    return out_idx, out_feat, BitGrid(bits, prefix, None, shape, B, count)


def lift_gather(img_feat, pixels, cam_ids, points, lidar2img, downscale, score_weight, score_bias):
    """img_feat (ncam,C,h,w) any strides; pixels (M,3) f32; cam_ids (M) i32; points (M,P) f32;
    lidar2img (ncam,16) f32; score_weight (C+17) f32 -> (M, P+C) f32."""
    assert img_feat.dim() == 4 and img_feat.dtype == torch.float32
    ncam, C, h, w = img_feat.shape
    pixels, points = pixels.contiguous().float(), points.contiguous().float()
    cam_ids = cam_ids.contiguous().int()
    lidar2img = lidar2img.contiguous().float()
    score_weight = score_weight.detach().contiguous().float().view(-1)
    assert score_weight.numel() == C + 17 and lidar2img.shape == (ncam, 16)
    M, P = points.shape
    out = torch.empty((M, P + C), dtype=torch.float32, device=points.device)
    _cabi.require_cuda(img_feat, 'lift_gather needs CUDA tensors')
    s = img_feat.stride()
    with _Timed('lift_gather', m=M, c=C):
        check(lib().msmd_lift_gather(_cabi.ctypes.c_void_p(img_feat.data_ptr()), s[0], s[1], s[2], s[3],
                                     C, h, w, ptr(pixels), ptr(cam_ids), ptr(points), P, M,
                                     ptr(lidar2img), float(downscale), ptr(score_weight),
                                     float(score_bias), ptr(out), stream(points.device)),
              'msmd_lift_gather')
    return out


# --------------------------------------------------------------------------------------
# Gated Modality-Aware stage: row gather + fused gates / concatenation (csrc/gma.cu)
# reference: sparse_multimodal_encoder_painting.py:371-377, :391-401, :414-425
# --------------------------------------------------------------------------------------
def gather_rows(features, rows, coords=None):
    """features[rows] (and coords[rows]) in one launch; rows int64."""
    features = features.contiguous()
    n, c = rows.shape[0], features.shape[1]
    out = torch.empty((n, c), dtype=torch.float32, device=features.device)
    oc = torch.empty((n, 4), dtype=torch.int32, device=features.device) if coords is not None else None
    with _Timed('gather_rows', n=n, c=c):
        check(lib().msmd_gather_rows(ptr(features), c, ptr(coords), ptr(rows), n, ptr(out), ptr(oc),
                                     stream(features.device)), 'msmd_gather_rows')
    return (out, oc) if coords is not None else out


def gma_assemble(y_only3, idx_only3, feat3, feat2, bz2, only2_rows, only2_bzyx, nn_idx, syn3, syn2, dummy,
                 w_cross, b_cross, w_gate, b_gate):
    """Unified voxel list of one GMA stage (one sample per GPU): (features (n, c3 + 64), indices (n, 4))."""
    dev = feat3.device
    n_o3, c3 = y_only3.shape
    n_o2, n_mix = only2_bzyx.shape[0], syn3.shape[0]
    total = n_o3 + max(n_o2, 1) + max(n_mix, 1)
    out = torch.empty((total, c3 + feat2.shape[1]), dtype=torch.float32, device=dev)
    oidx = torch.empty((total, 4), dtype=torch.int32, device=dev)
    with _Timed('gma_assemble', n=total, c3=c3):
        check(lib().msmd_gma_assemble(ptr(y_only3), ptr(idx_only3), n_o3, ptr(feat3), feat3.shape[0], c3, ptr(feat2),
                                      ptr(bz2), feat2.shape[0], feat2.shape[1], ptr(only2_rows), ptr(only2_bzyx),
                                      ptr(nn_idx), n_o2, ptr(syn3) if n_mix else None, ptr(syn2) if n_mix else None,
                                      n_mix, ptr(dummy), ptr(w_cross), ptr(b_cross), ptr(w_gate), ptr(b_gate),
                                      ptr(out), ptr(oidx), stream(dev)), 'msmd_gma_assemble')
    return out, oidx
