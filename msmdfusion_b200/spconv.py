"""Host-side mirror of the ``spconv.pytorch`` surface the reference hot path uses.

Same names, constructor arguments and error behaviour as spconv v2.1.21 as frozen by the
reference's API copy ``bug_fix/conv.py`` (SparseConvolution ``:40-486``, SubMConv3d
``:545-572``, SparseConv3d ``:871-899``) -- so ``import msmdfusion_b200.spconv as spconv``
is a drop-in for ``import spconv.pytorch as spconv`` on this path
(``mmdet3d/ops/sparse_block.py:5``, ``middle_encoders/sparse_encoder.py:6``).  All compute
goes through the C ABI (``ops.py``); there is no CPU fallback.

B200-first differences that cannot change results:
* every active-voxel set owns one occupancy bit grid (``IndexSet``); SubM rulebooks are
  cached on it, so the 16 key-less SubM layers of the LiDAR backbone build 4 rulebooks
  instead of 16 (no indice_key on this path is ever re-used, so caching by
  index-set identity is a pure optimisation);
* ``SparseSequential`` and ``SparseBasicBlock`` fuse eval-mode BatchNorm1d, the residual
  add and ReLU into the convolution's store.
"""
import math
import os
from collections import OrderedDict

import torch
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

from . import autograd as _ag
from . import ops
from .registry import CONV_LAYERS


# contraction kernel: 'tc' (tcgen05, 3xTF32) or 'simt' (exact fp32 FFMA); MSMD_CONV_PATH overrides
CONV_PATH = os.environ.get('MSMD_CONV_PATH', 'tc')
# operand precision of the tensor-core path:
#   'bf16x3c' (DEFAULT since round 2) bf16 hi/lo split, 3 MMAs per product, fp32 accumulate (~2.5e-5 absolute per
#             layer at |x| ~ 5) through the split-bf16 operand cache of csrc/spconv_sb.cu: activations are split by
#             their producer, the gather is cp.async -- the fastest parity mode on the B200 (profiles/r02*)
#   'bf16x3'  the same arithmetic, split inside the gather warps (csrc/spconv_tc16.cu)
#   'tf32x3'  3xTF32 (csrc/spconv_tc.cu): ~1e-6 of fp32 per layer, the most accurate and the slowest
#   'bf16'    operands rounded to bf16, one MMA: the train-step arithmetic BASELINE configs[4] names -- outside the
#             inference parity bound
# MSMD_CONV_PRECISION overrides; train.VoxelSpaceTrainStep(precision=...) sets it for a train step.
CONV_PRECISION = os.environ.get('MSMD_CONV_PRECISION', 'bf16x3c')
# opt-in: mask-sorted tiles for the 3x3x3 SubM layers of the tensor-core path (spconv-2.x
# mask_argsort_fwd_splits); MSMD_MASK_SORT=1 also switches it on inside the native executor
MASK_SORT = os.environ.get('MSMD_MASK_SORT', '0') not in ('', '0')


# Kernel-layout weight copies, BatchNorm folds and executor plans are cached and re-derived when a parameter
# is replaced or modified in place (data_ptr / _version).  Writes through ``.data`` (mmcv EMAHook's parameter
# swap, ``p.data.copy_()``) bump neither: call ``invalidate_caches()`` after such writes.  ``load_state_dict``
# on any sparse convolution does it automatically.
_CACHE_EPOCH = [0]


def invalidate_caches():
    """Drop every derived copy of the parameters (packed weights, BN folds, native-executor plans)."""
    _CACHE_EPOCH[0] += 1


def cache_epoch():
    return _CACHE_EPOCH[0]


def expand_nd(ndim, val):
    if isinstance(val, (list, tuple)):
        assert len(val) == ndim
        return [int(v) for v in val]
    return [int(val)] * ndim


class IndiceData:
    """What ``find_indice_pair`` returns (ImplicitGemmIndiceData, bug_fix/conv.py:415-432)."""

    def __init__(self, out_indices, indices, pair_fwd, is_subm, spatial_shape, out_spatial_shape,
                 ksize, stride, padding, dilation, algo='MaskImplicitGemm'):
        self.out_indices = out_indices
        self.indices = indices
        self.pair_fwd = pair_fwd
        self.pair_bwd = None
        self.is_subm = is_subm
        self.spatial_shape = spatial_shape
        self.out_spatial_shape = out_spatial_shape
        self.ksize, self.stride, self.padding, self.dilation = ksize, stride, padding, dilation
        self.algo = algo


class IndexSet:
    """One set of active voxels: (N,4) int32 indices + its bit grid + cached SubM rulebooks."""

    def __init__(self, indices, spatial_shape, batch_size, grid=None, unique=False):
        self.indices = indices
        self.spatial_shape = list(spatial_shape)
        self.batch_size = int(batch_size)
        self._grid = grid
        self._subm = {}
        # True when the rows are known to be distinct voxels (enumerated from a bit grid: strided-conv
        # outputs, sparse_add); user-assigned index lists may repeat a coordinate
        self.unique = bool(unique)

    @property
    def grid(self):
        if self._grid is None:
            self._grid = ops.grid_build(self.indices, self.batch_size, self.spatial_shape,
                                        need_perm=True)
        return self._grid

    def subm_pairs(self, ksize, dilation):
        key = (tuple(ksize), tuple(dilation))
        pair = self._subm.get(key)
        if pair is None:
            pair = ops.rulebook_subm(self.indices, self.grid, ksize, dilation)
            self._subm[key] = pair
        return pair

    def subm_pairs_sorted(self, ksize, dilation):
        """(row_perm, pair_sorted) of the SubM rulebook (``ops.rulebook_mask_sort``), built once."""
        key = ('sorted', tuple(ksize), tuple(dilation))
        ent = self._subm.get(key)
        if ent is None:
            ent = self._subm[key] = ops.rulebook_mask_sort(self.subm_pairs(ksize, dilation))
        return ent

    def rulebook_record(self, pair, subm, ksize, dilation, path):
        """The dict ``autograd.SparseConvFunction`` differentiates through.  One record per rulebook
        tensor, so that the transposed table of a strided conv is built once however many layers or
        backward passes use it."""
        recs = self.__dict__.setdefault('_records', {})
        rec = recs.get(id(pair))
        if rec is None or rec['pair_fwd'] is not pair:
            rec = recs[id(pair)] = dict(pair_fwd=pair, subm=bool(subm), path=path, unique=self.unique)
        return rec


class _NullTimer:
    """Stand-in for spconv's CUDAKernelTimer: ``tensor._timer.namespace(name)`` is a no-op context
    (bug_fix/conv.py:380 enters it unconditionally)."""

    def namespace(self, name):
        import contextlib
        return contextlib.nullcontext()


_NULL_TIMER = _NullTimer()


class SparseConvTensor:
    """spconv.SparseConvTensor(features, indices, spatial_shape, batch_size)."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None,
                 indice_dict=None, benchmark=False):
        self._features = features
        self._indices = indices
        self.spatial_shape = list(spatial_shape)
        self.batch_size = int(batch_size) if not torch.is_tensor(batch_size) else int(batch_size.item())
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark
        self.benchmark_record = {}
        self.thrust_allocator = None
        self._timer = _NULL_TIMER
        self._iset = None

    # -- features / indices ------------------------------------------------------------
    @property
    def features(self):
        return self._features

    @features.setter
    def features(self, val):
        self._features = val

    @property
    def indices(self):
        return self._indices

    @indices.setter
    def indices(self, val):
        # MSMDFusion.py:322-323 assigns (N,5) indices; any assignment drops cached structure
        self._indices = val
        self._iset = None

    def index_set(self):
        """The cached IndexSet (bit grid + SubM rulebooks) of the current indices."""
        return _iset_of(self)

    @property
    def spatial_size(self):
        return int(torch.tensor(self.spatial_shape).prod().item())

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def replace_feature(self, feature):
        new = self.shadow_copy()
        new._features = feature
        return new

    def shadow_copy(self):
        t = SparseConvTensor(self._features, self._indices, self.spatial_shape, self.batch_size,
                             self.grid, self.voxel_num, self.indice_dict, self.benchmark)
        t.benchmark_record = self.benchmark_record
        t.thrust_allocator = self.thrust_allocator
        t._timer = self._timer
        t._iset = self._iset
        return t

    def dense(self, channels_first=True):
        idx = self._indices if self._indices.dtype == torch.int32 else self._indices.int()
        if torch.is_grad_enabled() and self._features.requires_grad:
            out = _ag.ToDenseFunction.apply(self._features, idx.contiguous(), self.spatial_shape,
                                            self.batch_size)
        else:
            out = ops.to_dense(idx, self._features, self.spatial_shape, self.batch_size)
        if not channels_first:
            return out.permute(0, 2, 3, 4, 1).contiguous()
        return out

    @property
    def sparity(self):
        return self._indices.shape[0] / (self.spatial_size * self.batch_size)


def _iset_of(t):
    """IndexSet of a tensor, rebuilt only when the user replaced ``indices``."""
    s = t._iset
    if s is not None and getattr(s, '_user', None) is t._indices and \
            s.spatial_shape == list(t.spatial_shape):
        return s
    idx = t._indices
    if idx.dim() != 2 or idx.shape[1] != 4:
        raise ValueError('sparse convolution needs (N,4) indices (batch,z,y,x), got '
                         f'{tuple(idx.shape)}')
    idx32 = idx if idx.dtype == torch.int32 else idx.int()
    s = IndexSet(idx32.contiguous(), t.spatial_shape, t.batch_size)
    s._user = t._indices
    t._iset = s
    return s


def _attach_iset(t, iset):
    iset._user = t._indices
    t._iset = iset


class SparseModule(nn.Module):
    """Marker base class: modules that take a SparseConvTensor (spconv.SparseModule)."""

    def __init__(self, name=None):
        super().__init__()
        self.name = name
        self._sparse_unique_name = ''


def is_spconv_module(module):
    return isinstance(module, SparseModule)


def _bn_scale_shift(bn):
    """Fold an eval-mode BatchNorm1d into y = x*scale + shift (fp32).  Cached on the module,
    keyed by the versions of its parameters/buffers, so inference folds each BN once."""
    key = (_CACHE_EPOCH[0],) + tuple((t.data_ptr(), t._version) for t in
                                      (bn.weight, bn.bias, bn.running_mean, bn.running_var) if t is not None)
    cached = getattr(bn, '_msmd_fold', None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    scale, shift = _bn_scale_shift_compute(bn)
    bn._msmd_fold = (key, scale, shift)
    return scale, shift


def _bn_scale_shift_compute(bn):
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    w = bn.weight.detach().float() if bn.weight is not None else torch.ones_like(inv)
    b = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(inv)
    scale = (w * inv).contiguous()
    shift = (b - bn.running_mean.float() * scale).contiguous()
    return scale, shift


def _bn_foldable(m, inference=False):
    """Eval-mode BatchNorm1d whose y = x*scale + shift can ride in a conv epilogue.  Not while autograd would need a
    gradient for its affine parameters (frozen statistics, trainable weight / bias): the fold detaches them, so
    such a BN runs as the torch op.  ``inference``: the caller only ever runs without autograd (the executor plans)."""
    if not (isinstance(m, nn.BatchNorm1d) and not m.training and m.track_running_stats and
            m.running_var is not None):
        return False
    if not inference and torch.is_grad_enabled() and \
            any(p is not None and p.requires_grad for p in (m.weight, m.bias)):
        return False
    return True


class SparseSequential(SparseModule):
    """spconv.SparseSequential: mixes sparse modules and plain modules acting on features.

    conv -> BatchNorm1d(eval) -> ReLU runs is executed as ONE fused kernel launch.
    """

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError('name exists.')
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError(f'index {idx} is out of range')
        if idx < 0:
            idx += len(self)
        it = iter(self._modules.values())
        for _ in range(idx):
            next(it)
        return next(it)

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError('name exists')
        self.add_module(name, module)

    def forward(self, input):
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, SparseConvolution) and not m.conv1x1 and i + 1 < len(mods) and \
                    _bn_foldable(mods[i + 1]) and isinstance(input, SparseConvTensor):
                relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                scale, shift = _bn_scale_shift(mods[i + 1])
                input = m.forward_fused(input, scale, shift, relu=relu)
                i += 3 if relu else 2
                continue
            if is_spconv_module(m):
                input = m(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input = input.replace_feature(m(input.features))
            else:
                input = m(input)
            i += 1
        return input


class SparseConvolution(SparseModule):
    """bug_fix/conv.py:40-486.  ``algo``/``fp32_accum`` are accepted and ignored: there is one
    algorithm (bit-grid rulebook + gathered contraction, fp32 accumulate)."""

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0,
                 dilation=1, groups=1, bias=True, subm=False, output_padding=0, transposed=False,
                 inverse=False, indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(name=name)
        assert groups == 1, "don't support groups for now"
        assert ndim == 3, 'only 3-D sparse convolution is on the MSMDFusion hot path'
        if transposed or inverse:
            raise NotImplementedError('transposed / inverse sparse conv is outside the hot path')
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = expand_nd(ndim, kernel_size)
        self.stride = expand_nd(ndim, stride)
        kv = int(math.prod(self.kernel_size))
        kv_stride = int(math.prod(self.stride))
        self.dilation = expand_nd(ndim, dilation)
        self.padding = expand_nd(ndim, padding)
        self.conv1x1 = kv == 1
        if not subm:
            self.conv1x1 &= kv_stride == 1
            if self.conv1x1:
                assert self.padding == [0] * ndim, 'padding must be zero for 1x1 conv (k=1,s=1)'
        self.transposed = transposed
        self.inverse = inverse
        self.output_padding = expand_nd(ndim, output_padding)
        self.groups = groups
        self.subm = subm
        self.indice_key = indice_key
        self.algo = algo
        self.fp32_accum = fp32_accum
        # KRSC, the spconv-2.x implicit-GEMM layout (bug_fix/conv.py:114-117)
        self.weight = Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()
        self._packed = None
        self._packed_key = None
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: invalidate_caches())

    def extra_repr(self):
        s = f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}'
        if self.padding != [0] * self.ndim:
            s += f', padding={self.padding}'
        if self.dilation != [1] * self.ndim:
            s += f', dilation={self.dilation}'
        if self.bias is None:
            s += ', bias=False'
        return s

    def reset_parameters(self):
        # kaiming_uniform_(a=sqrt(5)) on fan_in = Cin * prod(kernel)  (bug_fix/conv.py:142-179)
        fan_in = self.in_channels * int(math.prod(self.kernel_size))
        gain = math.sqrt(2.0 / (1 + 5.0))
        bound = math.sqrt(3.0) * gain / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def packed_weight(self):
        """Kernel-layout copy of the weight, re-packed only when the parameter changes.
        ``CONV_PATH`` selects the contraction kernel: 'tc' = tcgen05 tensor cores (3xTF32,
        default), 'simt' = exact-fp32 FFMA kernel."""
        w = self.weight
        kvol = int(math.prod(self.kernel_size))
        use_tc = CONV_PATH == 'tc' and ops.tc_supported(self.out_channels, kvol, self.in_channels)
        mode = ops.TC_MODES[CONV_PRECISION] if use_tc else 0
        key = (_CACHE_EPOCH[0], w.data_ptr(), w._version, w.device, mode)
        if self._packed is None or self._packed_key != key:
            self._packed = ops.pack_weight_tc(w, mode) if use_tc else ops.pack_weight(w)
            self._packed_key = key
        return self._packed

    def forward(self, input):
        return self.forward_fused(input, None, None, relu=False, residual=None)

    def forward_fused(self, input, scale=None, shift=None, relu=False, residual=None):
        """conv (+ optional per-channel scale/shift, residual add, ReLU) in one launch."""
        assert isinstance(input, SparseConvTensor)
        assert input.features.shape[1] == self.in_channels, 'channel size mismatch'
        features = input.features
        # train step (config 5): differentiate through the convolution; the epilogue terms become
        # plain torch ops so that autograd sees them
        with_grad = torch.is_grad_enabled() and (features.requires_grad or self.weight.requires_grad)
        if self.conv1x1:
            out = torch.mm(features, self.weight.view(self.out_channels, self.in_channels).T)
            if self.bias is not None:
                out = out + self.bias
            if scale is not None:
                out = out * scale + shift
            if residual is not None:
                out = out + residual
            if relu:
                out = torch.relu(out)
            t = input.replace_feature(out)
            return t
        iset = _iset_of(input)
        indice_dict = input.indice_dict.copy()
        bias_grad = self.bias is not None and with_grad and self.bias.requires_grad
        if self.bias is not None and not bias_grad:
            # y = (conv + b)*scale + shift  ==  conv*scale + (b*scale + shift)
            b = self.bias.detach().float()
            if scale is None:
                scale, shift = torch.ones_like(b), b.clone()
            else:
                shift = shift + b * scale
        if self.subm:
            datas = input.find_indice_pair(self.indice_key)
            if datas is not None:
                self._check_subm_reuse_valid(input, datas)
            pair = iset.subm_pairs(self.kernel_size, self.dilation)
            out_iset = iset
            out_shape = input.spatial_shape
            if self.indice_key is not None and datas is None:
                indice_dict[self.indice_key] = IndiceData(
                    iset.indices, iset.indices, pair, True, input.spatial_shape, out_shape,
                    self.kernel_size, self.stride, self.padding, self.dilation)
        else:
            if self.indice_key is not None:
                assert self.indice_key not in indice_dict, \
                    f'your indice key {self.indice_key} already exists in this sparse tensor.'
            out_indices, pair, out_grid = ops.rulebook_conv(
                iset.indices, iset.grid, self.kernel_size, self.stride, self.padding, self.dilation)
            out_shape = out_grid.spatial_shape
            out_iset = IndexSet(out_indices, out_shape, input.batch_size, grid=out_grid, unique=True)
            if self.indice_key is not None:
                indice_dict[self.indice_key] = IndiceData(
                    out_indices, iset.indices, pair, False, input.spatial_shape, out_shape,
                    self.kernel_size, self.stride, self.padding, self.dilation)
        if with_grad:
            rb = iset.rulebook_record(pair, self.subm, self.kernel_size, self.dilation, CONV_PATH)
            if self.subm:   # the data gradient reuses pair_fwd with the offsets mirrored: k <-> K-1-k needs odd sizes
                assert all(k % 2 == 1 for k in self.kernel_size), 'SubM backward: odd kernel sizes only'
            out_features = _ag.SparseConvFunction.apply(features, self.weight, self.packed_weight(), rb)
            if bias_grad:   # a trainable bias stays a torch op, so that autograd reaches it
                out_features = out_features + self.bias
            if scale is not None:
                out_features = out_features * scale + shift
            if residual is not None:
                out_features = out_features + residual
            if relu:
                out_features = torch.relu(out_features)
        else:
            packed = self.packed_weight()
            if MASK_SORT and self.subm and pair.shape[0] == 27 and isinstance(packed, ops.TcWeight) \
                    and pair.shape[1] > 0:
                row_perm, pair_sorted = iset.subm_pairs_sorted(self.kernel_size, self.dilation)
                out_features = ops.spconv_fwd_tc(features, packed, pair_sorted, scale, shift, residual, relu,
                                                 row_perm=row_perm)
            else:
                out_features = ops.spconv_fwd(features, packed, pair, scale, shift, residual, relu)
        out = SparseConvTensor(out_features, out_iset.indices, out_shape, input.batch_size,
                               indice_dict=indice_dict, benchmark=input.benchmark)
        out.benchmark_record = input.benchmark_record
        out._timer = input._timer
        _attach_iset(out, out_iset)
        return out

    def _check_subm_reuse_valid(self, inp, datas):
        assert datas.is_subm, 'only support reuse subm indices'
        if self.kernel_size != datas.ksize:
            raise ValueError(f'subm with same indice_key must have same kernel size, expect '
                             f'{datas.ksize}, this layer {self.kernel_size}')
        if self.dilation != datas.dilation:
            raise ValueError(f'subm with same indice_key must have same dilation, expect '
                             f'{datas.dilation}, this layer {self.dilation}')
        if inp.spatial_shape != datas.spatial_shape:
            raise ValueError(f'subm with same indice_key must have same spatial structure, expect '
                             f'{datas.spatial_shape}, input {inp.spatial_shape}')
        if inp.indices.shape[0] != datas.indices.shape[0]:
            raise ValueError(f'subm with same indice_key must have same num of indices, expect '
                             f'{datas.indices.shape[0]}, input {inp.indices.shape[0]}')


@CONV_LAYERS.register_module()
class SubMConv3d(SparseConvolution):
    """bug_fix/conv.py:545-572"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, True, indice_key=indice_key, algo=algo, fp32_accum=fp32_accum,
                         name=name)


@CONV_LAYERS.register_module()
class SparseConv3d(SparseConvolution):
    """bug_fix/conv.py:871-899"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias, indice_key=indice_key, algo=algo, fp32_accum=fp32_accum, name=name)
