"""CPU tests that RUN the SIMT CUDA kernels of the train step on a host emulation of the CUDA execution
model (tests/tools/cuda_emul: one std::thread per CUDA thread, std::barrier for __syncthreads, warp
shuffles through an exchange buffer; the .cu sources are compiled as they are, launches rewritten by a
script).  Purpose: the backward kernels were written in a session without GPU time, so their indexing,
barrier placement and host-side launch logic are checked here against the oracle; the emulator itself
is calibrated on the forward SIMT kernel, which is verified on the B200 by the ``-m gpu`` tests.
The tensor-core kernels (tcgen05 / TMEM / bulk copy, csrc/spconv_tc.cu) run on a functional model of
what csrc/tc.cuh wraps (tests/tools/cuda_emul/tc_emul.h), calibrated the same way on the GPU-verified
variants before it is trusted with the paths that have not run on hardware (mask-sorted tiles).  Cluster
kernels (FPS) are not covered."""
import ctypes
import importlib.util
import os

import numpy as np
import pytest

from oracle import cpu

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def emu():
    global _LIB
    if _LIB is None:
        spec = importlib.util.spec_from_file_location('emul_build', os.path.join(HERE, 'tools', 'cuda_emul', 'build.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _LIB = ctypes.CDLL(mod.build())
        _LIB.emu_last_error.restype = ctypes.c_char_p
        _LIB.emu_msmd_spconv_bwd_weight_workspace.restype = ctypes.c_size_t
    return _LIB


_TC = None


def tc_emu():
    """Host-emulated csrc/spconv_tc.cu (its own library: the tcgen05 model hooks the CTA launch)."""
    global _TC
    if _TC is None:
        spec = importlib.util.spec_from_file_location('emul_build', os.path.join(HERE, 'tools', 'cuda_emul', 'build.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _TC = ctypes.CDLL(mod.build_tc())
        _TC.emu_last_error.restype = ctypes.c_char_p
        _TC.emu_msmd_spconv_tc_packed_floats.restype = ctypes.c_size_t
        _TC.emu_msmd_spconv_tc_workspace.restype = ctypes.c_size_t
        _TC.emu_msmd_spconv_tc16_packed_bytes.restype = ctypes.c_size_t
        _TC.emu_msmd_spconv_bwd_weight_tc_workspace.restype = ctypes.c_size_t
        _TC.emu_msmd_spconv_tc16_workspace.restype = ctypes.c_size_t
        _TC.emu_msmd_spconv_sb_packed_bytes.restype = ctypes.c_size_t
    return _TC


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def ok(status):
    assert status == 0, emu().emu_last_error()


def random_sparse(seed, batch, shape, n, c):
    rng = np.random.default_rng(seed)
    D, H, W = shape
    lin = rng.choice(batch * D * H * W, size=n, replace=False)
    idx = np.stack([lin // (D * H * W), (lin // (H * W)) % D, (lin // W) % H, lin % W], 1).astype(np.int32)
    return idx, rng.standard_normal((n, c)).astype(np.float32)


def rel(a, b):
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def emu_pack(w):
    cout, cin = w.shape[0], w.shape[-1]
    kvol = w.size // (cout * cin)
    packed = np.empty((kvol, cin, cout), np.float32)
    ok(emu().emu_msmd_spconv_pack_weight(P(w), cout, kvol, cin, P(packed), None))
    return packed


def emu_fwd(feat, w, pair, scale=None, shift=None, residual=None, relu=0):
    cout, cin = w.shape[0], w.shape[-1]
    kvol, n_out = pair.shape
    out = np.full((n_out, cout), np.nan, np.float32)
    ok(emu().msmd_spconv_fwd(P(feat), feat.shape[0], P(emu_pack(w)), P(pair), n_out, cin, cout, kvol, P(scale),
                             P(shift), P(residual), relu, P(out), None))
    return out


@pytest.mark.parametrize('cin,cout', [(5, 16), (16, 32), (8, 72)])
def test_emulator_calibration_forward_simt_kernel(cin, cout):
    """The GPU-verified forward kernel, run on the emulator, reproduces the oracle (incl. the fused
    scale / shift / residual / ReLU epilogue) -- i.e. the emulator executes this code base faithfully."""
    shape, batch = [7, 14, 14], 2
    idx, feat = random_sparse(0, batch, shape, 500, cin)
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    ref = cpu.spconv_fwd(feat, w, pair)
    assert rel(emu_fwd(feat, w, pair), ref) < 1e-5
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal(ref.shape).astype(np.float32)
    got = emu_fwd(feat, w, pair, scale, shift, res, 1)
    assert rel(got, np.maximum(ref * scale + shift + res, 0)) < 1e-5


@pytest.mark.parametrize('cin,cout,ksize,subm,n', [(8, 12, 3, True, 700), (5, 6, 3, True, 300),
                                                    (16, 16, 3, False, 900), (68, 72, (3, 1, 1), True, 600),
                                                    (4, 4, 3, True, 40)])
def test_wgrad_kernel_on_emulator(cin, cout, ksize, subm, n):
    """msmd_spconv_bwd_weight (row-slice partial tiles + fixed-order reduction): vector and scalar
    staging paths, multi-tile channels, several row slices, a single short chunk."""
    shape, batch = [7, 16, 16], 2
    idx, feat = random_sparse(3, batch, shape, n, cin)
    rng = np.random.default_rng(4)
    ks = cpu._triple(ksize)
    if subm:
        pair = cpu.subm_rulebook(idx, shape, ks, 1)
    else:
        _, pair, _ = cpu.conv_rulebook(idx, shape, ks, 2, 1, 1)
    kvol, n_out = pair.shape
    go = rng.standard_normal((n_out, cout)).astype(np.float32)
    _, ref = cpu.spconv_bwd(feat, np.zeros((cout, *ks, cin), np.float32), pair, go, need_input_grad=False)
    need = emu().emu_msmd_spconv_bwd_weight_workspace(n_out, cin, cout, kvol)
    assert need > 0 and need % (kvol * cin * cout * 4) == 0
    ws = np.full(need // 4, np.nan, np.float32)
    gw = np.full((cout, *ks, cin), np.nan, np.float32)
    ok(emu().emu_msmd_spconv_bwd_weight(P(feat), feat.shape[0], P(go), P(pair), n_out, cin, cout, kvol, P(gw),
                                        P(ws), ctypes.c_size_t(need), None))
    assert np.isfinite(ws).all(), 'every partial tile must be written (the reduction reads all of them)'
    assert rel(gw, ref) < 1e-5
    # too small a workspace is an error, not an overrun
    assert emu().emu_msmd_spconv_bwd_weight(P(feat), feat.shape[0], P(go), P(pair), n_out, cin, cout, kvol, P(gw),
                                            P(ws), ctypes.c_size_t(need - 4), None) == -3


def test_wgrad_empty_inputs_on_emulator():
    gw = np.full((4, 27, 3), np.nan, np.float32)
    ok(emu().emu_msmd_spconv_bwd_weight(None, 0, None, None, 0, 3, 4, 27, P(gw), None, ctypes.c_size_t(0), None))
    assert (gw == 0).all()


@pytest.mark.parametrize('subm', [True, False])
def test_dgrad_path_on_emulator(subm):
    """msmd_rulebook_transpose + msmd_spconv_transpose_weight + msmd_spconv_bwd_data (SIMT layout) ==
    the oracle's data gradient; for SubM the mirrored-offset form needs no second table."""
    shape, batch, cin, cout = [7, 14, 14], 2, 8, 12
    idx, feat = random_sparse(5, batch, shape, 500, cin)
    rng = np.random.default_rng(6)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    if subm:
        pair = cpu.subm_rulebook(idx, shape, 3, 1)
    else:
        _, pair, _ = cpu.conv_rulebook(idx, shape, 3, 2, 1, 1)
    kvol, n_out = pair.shape
    n_in = idx.shape[0]
    go = rng.standard_normal((n_out, cout)).astype(np.float32)
    ref, _ = cpu.spconv_bwd(feat, w, pair, go, need_weight_grad=False)

    wt = np.full((cin, 3, 3, 3, cout), np.nan, np.float32)
    ok(emu().emu_msmd_spconv_transpose_weight(P(w), cout, kvol, cin, int(subm), P(wt), None))
    w3 = w.reshape(cout, kvol, cin)
    expect = (w3[:, ::-1] if subm else w3).transpose(2, 1, 0)
    assert np.array_equal(wt.reshape(cin, kvol, cout), expect)
    if subm:
        pair_bwd = pair
    else:
        pair_bwd = np.full((kvol, n_in), 7, np.int32)
        ok(emu().emu_msmd_rulebook_transpose(P(pair), kvol, n_out, n_in, P(pair_bwd), None))
        assert np.array_equal(pair_bwd, cpu.pair_transpose(pair, n_in))
    packed = emu_pack(wt)                       # "cout" = cin, "cin" = cout
    gi = np.full((n_in, cin), np.nan, np.float32)
    ok(emu().emu_msmd_spconv_bwd_data(P(go), n_out, P(packed), 0, P(pair_bwd), n_in, cin, cout, kvol, P(gi), None,
                                      ctypes.c_size_t(0), None))
    assert rel(gi, ref) < 1e-5


def test_from_dense_and_grid_rows_on_emulator():
    shape, batch, c = [5, 9, 11], 2, 7
    idx, feat = random_sparse(7, batch, shape, 300, c)
    dense = cpu.dense(idx, feat, shape, batch)
    out = np.full((idx.shape[0], c), np.nan, np.float32)
    sh = (ctypes.c_int * 3)(*shape)
    ok(emu().emu_msmd_from_dense(P(idx), P(dense), idx.shape[0], c, batch, sh, P(out), None))
    assert np.array_equal(out, feat)
    # bit grid of a superset (as sparse_add builds it): rank = row in ascending linear order
    extra, _ = random_sparse(8, batch, shape, 200, 1)
    D, H, W = shape
    lin = lambda a: ((a[:, 0].astype(np.int64) * D + a[:, 1]) * H + a[:, 2]) * W + a[:, 3]  # noqa: E731
    union = np.unique(np.concatenate([lin(idx), lin(extra)]))
    cells = batch * D * H * W
    words = (cells + 31) // 32
    bits = np.zeros(words, np.uint32)
    np.bitwise_or.at(bits, union >> 5, (np.uint32(1) << (union & 31).astype(np.uint32)))
    pop = np.array([bin(int(b)).count('1') for b in bits], np.int32)
    prefix = (np.cumsum(pop) - pop).astype(np.int32)
    probe = np.concatenate([idx, np.array([[0, 0, 0, 0], [batch, 0, 0, 0], [1, D - 1, H - 1, W - 1]], np.int32)])
    rows = np.full(probe.shape[0], -7, np.int32)
    ok(emu().emu_msmd_grid_rows(P(probe), probe.shape[0], batch, sh, P(bits), P(prefix), P(rows), None))
    pl = lin(probe)
    pos = np.searchsorted(union, pl)
    expect = np.where((probe[:, 0] < batch) & (pos < union.size) & (union[np.minimum(pos, union.size - 1)] == pl), pos, -1)
    assert np.array_equal(rows, expect.astype(np.int32))


# --------------------------------------------------------------------------------------
# the REAL tensor-level wrappers (msmdfusion_b200/ops.py, with the ctypes signatures of _cabi.py)
# driving the emulated kernels: wrapper argument order / types + kernels, end to end on the CPU
# --------------------------------------------------------------------------------------
class _EmuLib:
    """Looks like ``_cabi.lib()``: msmd_X resolves to the emulated entry with _cabi's own argtypes."""

    def __init__(self, cabi):
        self._cabi, self._cache = cabi, {}

    TC_UNIT = ('msmd_spconv_tc_', 'msmd_spconv_fwd_tc', 'msmd_spconv_tc16_', 'msmd_spconv_bwd_weight_tc', 'msmd_spconv_sb', 'msmd_spconv_fwd_sb', 'msmd_split_', 'msmd_rulebook_tile_masks')

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            if name == 'msmd_spconv_bwd_data':
                return self._bwd_data
            if name == 'msmd_spconv_set_wgrad_tc':
                return self._set_wgrad_tc
            if name in ('msmd_spconv_bwd_weight', 'msmd_spconv_bwd_weight_workspace') and self._wgrad_tc:
                name = name.replace('bwd_weight', 'bwd_weight_tc')   # what csrc/spconv_bwd.cu does when switched on
            L = tc_emu() if name.startswith(self.TC_UNIT) else emu()
            fn = getattr(L, name) if name == 'msmd_spconv_fwd' else getattr(L, 'emu_' + name)
            fn.restype, fn.argtypes = self._cabi.SIGNATURES[name]
            if not self._wgrad_tc:
                self._cache[name] = fn
        return fn

    _wgrad_tc = False

    def _set_wgrad_tc(self, enable):
        self._wgrad_tc = bool(enable)
        self._cache.clear()
        return 0

    def _bwd_data(self, go, n_out, wt, weight_tc, pair_bwd, n_in, cin, cout, kvol, gi, ws, ws_bytes, stream):
        """csrc/spconv_bwd.cu:msmd_spconv_bwd_data restated (three-way dispatch onto the forward entry points;
        the SIMT emulation unit cannot link the tensor-core one)."""
        if weight_tc in (2, 3):
            return self.msmd_spconv_fwd_tc16(go, n_out, wt, pair_bwd, None, n_in, cout, cin, kvol, int(weight_tc == 2),
                                             None, None, None, 0, gi, stream)
        if weight_tc:
            return self.msmd_spconv_fwd_tc_ws(go, n_out, wt, pair_bwd, n_in, cout, cin, kvol, None, None, None, 0, gi,
                                              ws, ws_bytes, stream)
        return self.msmd_spconv_fwd(go, n_out, wt, pair_bwd, n_in, cout, cin, kvol, None, None, None, 0, gi, stream)


@pytest.fixture()
def ops_on_emulator(monkeypatch):
    import torch
    from msmdfusion_b200 import _cabi, ops
    shim = _EmuLib(_cabi)

    def ptr(t):
        if t is None:
            return None
        assert t.is_contiguous()
        return ctypes.c_void_p(t.data_ptr())

    class Scratch:
        def get(self, device, nbytes, slot='ws'):
            return torch.empty(max(int(nbytes), 16), dtype=torch.uint8)
    for mod in (_cabi, ops):
        monkeypatch.setattr(mod, 'lib', lambda: shim)
        monkeypatch.setattr(mod, 'ptr', ptr)
        monkeypatch.setattr(mod, 'stream', lambda device=None: None)
        monkeypatch.setattr(mod, 'scratch', Scratch())
    monkeypatch.setattr(ops, 'tc_supported', lambda *a: False)
    return ops


def test_ops_wrappers_drive_emulated_kernels(ops_on_emulator):
    import torch
    ops = ops_on_emulator
    shape, batch, cin, cout = [7, 14, 14], 2, 8, 12
    idx, feat = random_sparse(9, batch, shape, 400, cin)
    rng = np.random.default_rng(10)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    oi, pair, oshape = cpu.conv_rulebook(idx, shape, 3, 2, 1, 1)
    go = rng.standard_normal((oi.shape[0], cout)).astype(np.float32)
    t = torch.from_numpy
    out = ops.spconv_fwd(t(feat), ops.pack_weight(t(w)), t(pair))
    assert rel(out.numpy(), cpu.spconv_fwd(feat, w, pair)) < 1e-5
    ri, rw = cpu.spconv_bwd(feat, w, pair, go)
    pair_bwd = ops.rulebook_transpose(t(pair), idx.shape[0])
    gi = ops.spconv_bwd_data(t(go), ops.pack_weight(ops.transpose_weight(t(w))), pair_bwd)
    assert rel(gi.numpy(), ri) < 1e-5
    gw = ops.spconv_bwd_weight(t(feat), t(go), t(pair), w.shape)
    assert tuple(gw.shape) == w.shape and rel(gw.numpy(), rw) < 1e-5
    d = ops.to_dense(t(oi), out, oshape, batch)
    assert np.array_equal(d.numpy(), cpu.dense(oi, out.numpy(), oshape, batch))
    assert np.array_equal(ops.from_dense(t(oi), d, oshape, batch).numpy(), out.numpy())


def test_conv_autograd_function_on_emulated_kernels(ops_on_emulator, monkeypatch):
    """SubMConv3d / SparseConv3d modules -> autograd.SparseConvFunction -> real wrappers -> emulated
    kernels, against the oracle's gradients (rulebooks come from the oracle: the bit-grid kernels use a
    decoupled look-back scan that a block-sequential emulator cannot run)."""
    import torch
    import _cpu_ops
    from msmdfusion_b200 import spconv
    ops = ops_on_emulator
    for name in ('grid_build', 'rulebook_subm', 'rulebook_conv'):
        monkeypatch.setattr(ops, name, _cpu_ops.STANDINS[name])
    shape, batch, cin, cout = [7, 12, 12], 1, 4, 8
    idx, feat = random_sparse(11, batch, shape, 250, cin)
    idx = np.concatenate([idx, idx[:9]])                 # duplicate coordinates in the SubM case
    feat = np.concatenate([feat, feat[:9] * 2])
    for cls, kw in ((spconv.SubMConv3d, dict(padding=1)), (spconv.SparseConv3d, dict(stride=2, padding=1))):
        torch.manual_seed(0)
        conv = cls(cin, cout, 3, bias=False, **kw)
        f = torch.from_numpy(feat).clone().requires_grad_(True)
        out = conv(spconv.SparseConvTensor(f, torch.from_numpy(idx), shape, batch))
        g = torch.randn(out.features.shape, generator=torch.Generator().manual_seed(1))
        (out.features * g).sum().backward()
        w = conv.weight.detach().numpy()
        if conv.subm:
            pair = cpu.subm_rulebook(idx, shape, 3, 1)
        else:
            _, pair, _ = cpu.conv_rulebook(idx, shape, 3, 2, 1, 1)
        ri, rw = cpu.spconv_bwd(feat, w, pair, g.numpy())
        assert rel(out.features.detach().numpy(), cpu.spconv_fwd(feat, w, pair)) < 1e-5
        assert rel(f.grad.numpy(), ri) < 1e-5
        assert rel(conv.weight.grad.numpy(), rw) < 1e-5
    # a trainable bias and a frozen-statistics BatchNorm with trainable affine stay torch ops under autograd, so
    # they receive their gradients (they are folded into the conv epilogue only when nothing asks for one)
    torch.manual_seed(0)
    block = spconv.SparseSequential(spconv.SubMConv3d(cin, cout, 3, padding=1, bias=True),
                                    torch.nn.BatchNorm1d(cout), torch.nn.ReLU())
    block[1].running_mean.normal_()
    block[1].running_var.uniform_(0.5, 2.0)
    block.eval()
    f = torch.from_numpy(feat).clone()
    out = block(spconv.SparseConvTensor(f, torch.from_numpy(idx), shape, batch))
    g = torch.randn(out.features.shape, generator=torch.Generator().manual_seed(2))
    (out.features * g).sum().backward()
    conv, bn = block[0], block[1]
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    pre = torch.from_numpy(cpu.spconv_fwd(feat, conv.weight.detach().numpy(), pair)) + conv.bias.detach()
    inv = torch.rsqrt(bn.running_var + bn.eps)
    xhat = (pre - bn.running_mean) * inv
    gate = ((xhat * bn.weight.detach() + bn.bias.detach()) > 0).float() * g
    assert rel(bn.bias.grad.numpy(), gate.sum(0).numpy()) < 1e-5
    assert rel(bn.weight.grad.numpy(), (gate * xhat).sum(0).numpy()) < 1e-5
    assert rel(conv.bias.grad.numpy(), (gate * bn.weight.detach() * inv).sum(0).numpy()) < 1e-5
    with torch.no_grad():   # and the fused inference form of the same block agrees with it
        fused = block(spconv.SparseConvTensor(f, torch.from_numpy(idx), shape, batch))
    assert rel(fused.features.numpy(), out.features.detach().numpy()) < 1e-5


# --------------------------------------------------------------------------------------
# mask-sorted tiles: digest + stable radix sort + table permutation (csrc/fusion.cu, sort.cuh, scan.cuh)
# --------------------------------------------------------------------------------------
def digest_key(mask):
    """The 15-bit structural digest (tests/tools/mask_sort_estimate.py): own z plane (9 bits), then one
    'any neighbour in this ky row' bit per row of the plane above, then of the plane below."""
    lower, centre, upper = mask & 0x1FF, (mask >> 9) & 0x1FF, (mask >> 18) & 0x1FF
    key = centre.copy()
    for plane, base in ((upper, 9), (lower, 12)):
        for row in range(3):
            key |= (((plane >> (3 * row)) & 7) != 0).astype(np.int64) << (base + row)
    return key


@pytest.mark.parametrize('case', ['plain', 'collisions'])
def test_modality_split_hash_and_sort_paths_on_emulator(case):
    """voxel_modality_split on the emulator, both implementations against the oracle bit for bit: the hash path
    (table on the float key + shared-memory sort of the paired rows; the default) and the sort path (stable radix
    sorts + binary-search merge; the GPU-verified round-1 code, also the overflow fall-back).  'collisions': z >= 17
    and x >= 1000, where distinct voxels share a float32 key (runs of several rows per key in BOTH sets)."""
    rng = np.random.default_rng(0 if case == 'plain' else 7)
    if case == 'plain':
        shape = [41, 60, 60]
        def coords(n):
            lin = rng.choice(int(np.prod(shape)), size=n, replace=False)
            return np.stack([np.zeros(n, np.int64), lin // 3600, (lin // 60) % 60, lin % 60], 1).astype(np.int32)
        n3, n2 = 2500, 3100
        c3, c2 = coords(n3), coords(n2)
        c2[:800] = c3[rng.choice(n3, 800, replace=False)]
    else:
        def coords(n):   # dense little patches high up and far right: neighbouring x collide, x >= 1000 wraps into y + 1
            z = rng.integers(17, 41, n)
            y = rng.integers(100, 104, n)
            x = rng.integers(990, 1040, n)
            c = np.unique(np.stack([np.zeros(n, np.int64), z, y, x], 1), axis=0)
            return c[rng.permutation(c.shape[0])].astype(np.int32)
        c3, c2 = coords(1500), coords(1800)
        n3, n2 = c3.shape[0], c2.shape[0]
    e3, e2, es3, es2 = cpu.voxel_modality_split(c3, c2, 1)
    L = emu()
    L.emu_msmd_modality_split_workspace.restype = ctypes.c_size_t
    need = L.emu_msmd_modality_split_workspace(n3, n2)
    for fn in (L.emu_msmd_modality_split, L.emu_msmd_modality_split_sort):
        m3, m2 = np.full(n3, -1, np.int32), np.full(n2, -1, np.int32)
        s3, s2 = np.full(min(n3, n2), -1, np.int64), np.full(min(n3, n2), -1, np.int64)
        cnt = np.zeros(2, np.int32)
        ws = np.zeros(need, np.uint8)
        ok(fn(P(c3), n3, P(c2), n2, ctypes.c_longlong(0), ctypes.c_longlong(0), P(m3), P(m2), P(s3), P(s2), P(cnt),
              P(ws), ctypes.c_size_t(need), None))
        assert cnt[1] == 0, 'unexpected overflow'
        assert np.array_equal(m3, e3[:, 1]) and np.array_equal(m2, e2[:, 1])
        p = int(cnt[0])
        assert p == es3.shape[0] and np.array_equal(s3[:p], es3) and np.array_equal(s2[:p], es2)
    if case == 'collisions':
        assert es3.shape[0] > 0 and (e3[:, 1].sum() != np.isin(c3.view([('', c3.dtype)] * 4), c2.view([('', c2.dtype)] * 4)).sum())


@pytest.mark.parametrize('n', [700, 5000])
def test_mask_sort_on_emulator(n):
    from msmdfusion_b200 import synthetic
    pts = synthetic.lidar_scene(seed=4, sweeps=1)[:n * 2]
    _, c, _ = cpu.hard_voxelize(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    idx = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], 1)[:n]
    n = idx.shape[0]
    pair = cpu.subm_rulebook(idx, [41, 1440, 1440], 3, 1)
    L = emu()
    L.emu_msmd_rulebook_mask_sort_workspace.restype = ctypes.c_size_t
    need = L.emu_msmd_rulebook_mask_sort_workspace(n)
    ws = np.zeros(need, np.uint8)
    perm = np.full(n, -1, np.int32)
    pair_sorted = np.full_like(pair, -9)
    ok(L.emu_msmd_rulebook_mask_sort(P(pair), 27, n, P(perm), P(pair_sorted), P(ws), ctypes.c_size_t(need), None))
    used = pair >= 0
    mask = (used.astype(np.int64) * (1 << np.arange(27))[:, None]).sum(0)
    expect = np.argsort(digest_key(mask), kind='stable')
    assert np.array_equal(perm, expect.astype(np.int32))
    assert np.array_equal(pair_sorted, pair[:, perm])
    # what the sort is for: fewer (tile, kernel offset) products than the index-set order
    def products(order):
        nt = (n + 127) // 128
        u = np.concatenate([used[:, order], np.zeros((27, nt * 128 - n), bool)], 1).reshape(27, nt, 128).any(2)
        return int(u.sum())
    assert products(perm) < products(np.arange(n))
    assert L.emu_msmd_rulebook_mask_sort(P(pair), 27, n, P(perm), P(pair_sorted), P(ws), ctypes.c_size_t(64), None) == -3
    assert L.emu_msmd_rulebook_mask_sort(P(pair), 9, n, P(perm), P(pair_sorted), P(ws), ctypes.c_size_t(need), None) == -1


# --------------------------------------------------------------------------------------
# tensor-core kernels (tcgen05 / TMEM / bulk copy) on the functional model of tc.cuh
# --------------------------------------------------------------------------------------
def tc_pack(w):
    L = tc_emu()
    cout, cin = w.shape[0], w.shape[-1]
    kvol = w.size // (cout * cin)
    packed = np.full(L.emu_msmd_spconv_tc_packed_floats(cout, kvol, cin), np.nan, np.float32)
    assert L.emu_msmd_spconv_tc_pack_weight(P(w), cout, kvol, cin, P(packed), None) == 0, L.emu_last_error()
    return packed


def tc_fwd(feat, w, pair, variant, split=False, scale=None, shift=None, residual=None, relu=0, row_perm=None):
    L = tc_emu()
    cout, cin = w.shape[0], w.shape[-1]
    kvol, n_out = pair.shape
    assert L.emu_msmd_spconv_tc_set_variant(variant) == 0
    need = L.emu_msmd_spconv_tc_workspace(n_out, cout) if split else 0
    ws = np.zeros(need // 4 + 64, np.float32) if need else None
    out = np.full((n_out, cout), np.nan, np.float32)
    args = (P(feat), feat.shape[0], tc_pack(w).ctypes.data_as(ctypes.c_void_p), P(pair))
    tail = (n_out, cin, cout, kvol, P(scale), P(shift), P(residual), relu, P(out), P(ws), ctypes.c_size_t(need), None)
    if row_perm is None:
        st = L.emu_msmd_spconv_fwd_tc_ws(*args, *tail)
    else:
        st = L.emu_msmd_spconv_fwd_tc_sorted(*args, P(row_perm), *tail)
    L.emu_msmd_spconv_tc_set_variant(0)
    assert st == 0, L.emu_last_error()
    return out, need


TC_CASES = [  # cin, cout, n, variant, split-K pairs
    (16, 16, 300, 0, False),    # variant 2, concatenated-B mode, vector gather
    (5, 16, 200, 0, False),     # scalar gather (cin % 4 != 0), four kernel offsets per K chunk
    (32, 64, 200, 2, False),    # variant 2, concatenated-B at N = 64
    (20, 144, 150, 2, False),   # variant 2, three MMAs per k-step (2N > 256), padded N
    (16, 128, 200, 0, False),   # variant 3 (A operand in tensor memory)
    (16, 128, 200, 0, True),    # variant 3, split-K CTA pairs + hand-off through the workspace
    (8, 40, 130, 3, False),     # variant 3 forced at a small N (three A stages), padded N, ragged last tile
]


@pytest.mark.parametrize('cin,cout,n,variant,split', TC_CASES)
def test_emulator_calibration_tc_kernels(cin, cout, n, variant, split):
    """The GPU-verified tcgen05 kernels (variant 2 / concatenated-B / variant 3 / split-K) run on the
    functional tcgen05 model and reproduce the oracle at 3xTF32 accuracy, with and without the fused
    epilogue -- i.e. descriptors, the 128-byte swizzle, the TMEM layout and the mbarrier protocol are
    modelled the way the hardware executes this code."""
    shape, batch = [5, 12, 12], 1
    idx, feat = random_sparse(0, batch, shape, n, cin)
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    ref = cpu.spconv_fwd(feat, w, pair)
    got, need = tc_fwd(feat, w, pair, variant, split)
    assert (need > 0) == split
    assert rel(got, ref) < 5e-6
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal(ref.shape).astype(np.float32)
    got, _ = tc_fwd(feat, w, pair, variant, split, scale, shift, res, 1)
    assert rel(got, np.maximum(ref * scale + shift + res, 0)) < 5e-6


def test_tc_kernel_on_emulator_strided_rulebook_and_empty_tiles():
    """A strided (27-offset, n_out != n_in) rulebook and a tile without any pair (all-zero rows + shift)."""
    shape = [7, 14, 14]
    idx, feat = random_sparse(5, 1, shape, 400, 16)
    _, pair, _ = cpu.conv_rulebook(idx, shape, (3, 3, 3), 2, 1, 1)
    rng = np.random.default_rng(6)
    w = (rng.standard_normal((32, 3, 3, 3, 16)) * 0.2).astype(np.float32)
    got, _ = tc_fwd(feat, w, pair, 0)
    assert rel(got, cpu.spconv_fwd(feat, w, pair)) < 5e-6
    empty = np.full((27, 140), -1, np.int32)
    empty[13, 130] = 7   # second tile: one pair; first tile: none
    shift = rng.standard_normal(32).astype(np.float32)
    scale = np.ones(32, np.float32)
    for variant in (2, 3):
        got, _ = tc_fwd(feat, w, empty, variant, scale=scale, shift=shift)
        assert rel(got, cpu.spconv_fwd(feat, w, empty) + shift) < 5e-6


@pytest.mark.parametrize('cin,cout,variant,split', [(16, 16, 0, False), (16, 72, 2, False), (16, 128, 0, True)])
def test_mask_sorted_tc_path_on_emulator(cin, cout, variant, split):
    """The opt-in mask-sorted path end to end, which has not run on a GPU yet: msmd_rulebook_mask_sort
    (SIMT emulation) -> permuted pair table + slot -> row map -> msmd_spconv_fwd_tc_sorted (tcgen05 model)
    writes every output row, residual included, where the unsorted kernel does."""
    from msmdfusion_b200 import synthetic
    pts = synthetic.lidar_scene(seed=4, sweeps=1)[:1200]
    _, c, _ = cpu.hard_voxelize(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    idx = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], 1)[:600]
    n = idx.shape[0]
    pair = cpu.subm_rulebook(idx, [41, 1440, 1440], 3, 1)
    L = emu()
    L.emu_msmd_rulebook_mask_sort_workspace.restype = ctypes.c_size_t
    need = L.emu_msmd_rulebook_mask_sort_workspace(n)
    ws = np.zeros(need, np.uint8)
    perm = np.full(n, -1, np.int32)
    pair_sorted = np.full_like(pair, -9)
    ok(L.emu_msmd_rulebook_mask_sort(P(pair), 27, n, P(perm), P(pair_sorted), P(ws), ctypes.c_size_t(need), None))
    rng = np.random.default_rng(2)
    feat = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal((n, cout)).astype(np.float32)
    ref = np.maximum(cpu.spconv_fwd(feat, w, pair) * scale + shift + res, 0)
    T = tc_emu()
    T.emu_tc_mma_count.restype = ctypes.c_longlong
    c0 = T.emu_tc_mma_count()
    got, _ = tc_fwd(feat, w, pair_sorted, variant, split, scale, shift, res, 1, row_perm=perm)
    c1 = T.emu_tc_mma_count()
    plain, _ = tc_fwd(feat, w, pair, variant, split, scale, shift, res, 1)
    c2 = T.emu_tc_mma_count()
    assert rel(got, ref) < 5e-6
    # per-row accumulation order is the kernel-offset order in both cases; only chunk skipping differs
    assert rel(got, plain) < 5e-6
    assert (c1 - c0) < (c2 - c1)   # fewer MMAs issued: the reason for the sort


# --------------------------------------------------------------------------------------
# 16-bit operand kernels (csrc/spconv_tc16.cu: bf16 / bf16x3) -- not yet run on hardware
# --------------------------------------------------------------------------------------
def bf16_round(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def tc16_fwd(feat, w, pair, x3, scale=None, shift=None, residual=None, relu=0, row_perm=None, variant=2, split=False):
    L = tc_emu()
    if variant != 2 or split:
        return _tc16_fwd_v3(feat, w, pair, x3, scale, shift, residual, relu, row_perm, variant, split)
    cout, cin = w.shape[0], w.shape[-1]
    kvol, n_out = pair.shape
    packed = np.full(L.emu_msmd_spconv_tc16_packed_bytes(cout, kvol, cin, x3) // 2, 0x7FC0, np.uint16)  # NaN fill
    assert L.emu_msmd_spconv_tc16_pack_weight(P(w), cout, kvol, cin, x3, P(packed), None) == 0, L.emu_last_error()
    out = np.full((n_out, cout), np.nan, np.float32)
    st = L.emu_msmd_spconv_fwd_tc16(P(feat), feat.shape[0], P(packed), P(pair), P(row_perm), n_out, cin, cout, kvol,
                                    x3, P(scale), P(shift), P(residual), relu, P(out), None)
    assert st == 0, L.emu_last_error()
    return out


@pytest.mark.parametrize('x3', [0, 1])
@pytest.mark.parametrize('cin,cout,n', [(16, 16, 300),    # four kernel offsets per 64-element chunk, vector gather
                                        (5, 16, 200),     # scalar gather, eight offsets per chunk
                                        (20, 144, 150),   # piece straddling cin (cin % 8 == 4), padded N, 3-MMA x3 mode
                                        (64, 128, 140)])  # one offset per chunk, concatenated-B x3 mode at 2N = 256
def test_tc16_kernels_on_emulator(x3, cin, cout, n):
    """bf16: equals the convolution of the bf16-ROUNDED operands (fp32 accumulate) -- the rounding is the
    only difference from the fp32 path.  bf16x3: within 2e-5 of the fp32 oracle (hi/lo split, lo*lo
    dropped).  Fused epilogue on both."""
    shape = [5, 12, 12]
    idx, feat = random_sparse(0, 1, shape, n, cin)
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    if x3:
        ref, tol = cpu.spconv_fwd(feat, w, pair), 2e-5
    else:
        ref, tol = cpu.spconv_fwd(bf16_round(feat), bf16_round(w), pair), 2e-6
    assert rel(tc16_fwd(feat, w, pair, x3), ref) < tol
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal(ref.shape).astype(np.float32)
    got = tc16_fwd(feat, w, pair, x3, scale, shift, res, 1)
    assert rel(got, np.maximum(ref * scale + shift + res, 0)) < tol
    if not x3:   # and it IS a different arithmetic from fp32: the parity bound does not hold
        assert rel(tc16_fwd(feat, w, pair, 0), cpu.spconv_fwd(feat, w, pair)) > 1e-4


def test_tc16_strided_rulebook_mask_sorted_and_empty_tiles_on_emulator():
    shape = [7, 14, 14]
    idx, feat = random_sparse(5, 1, shape, 400, 16)
    rng = np.random.default_rng(6)
    w = (rng.standard_normal((32, 3, 3, 3, 16)) * 0.2).astype(np.float32)
    _, pair, _ = cpu.conv_rulebook(idx, shape, (3, 3, 3), 2, 1, 1)
    assert rel(tc16_fwd(feat, w, pair, 1), cpu.spconv_fwd(feat, w, pair)) < 2e-5
    # mask-sorted SubM table through the slot -> row map
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    n = pair.shape[1]
    L = emu()
    L.emu_msmd_rulebook_mask_sort_workspace.restype = ctypes.c_size_t
    need = L.emu_msmd_rulebook_mask_sort_workspace(n)
    ws, perm, pair_sorted = np.zeros(need, np.uint8), np.full(n, -1, np.int32), np.full_like(pair, -9)
    ok(L.emu_msmd_rulebook_mask_sort(P(pair), 27, n, P(perm), P(pair_sorted), P(ws), ctypes.c_size_t(need), None))
    res = rng.standard_normal((n, 32)).astype(np.float32)
    one = np.ones(32, np.float32)
    for x3 in (0, 1):
        plain = tc16_fwd(feat, w, pair, x3, one, 0 * one, res, 1)
        assert np.array_equal(tc16_fwd(feat, w, pair_sorted, x3, one, 0 * one, res, 1, row_perm=perm), plain)
    # a tile without any pair: rows = shift
    empty = np.full((27, 140), -1, np.int32)
    empty[13, 130] = 7
    shift = rng.standard_normal(32).astype(np.float32)
    got = tc16_fwd(feat, w, empty, 0, one, shift)
    assert rel(got, cpu.spconv_fwd(bf16_round(feat), bf16_round(w), empty) + shift) < 2e-6


# --------------------------------------------------------------------------------------
# tensor-core weight gradient (csrc/spconv_wgrad_tc.cu) -- not yet run on hardware
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize('cin,cout,ksize,subm,n', [(16, 16, 3, True, 700),      # one co tile, N = 16
                                                    (8, 12, 3, True, 300),       # padded N (8 -> 16), cout % 16 != 0
                                                    (68, 72, (3, 1, 1), True, 600),   # N = 80, three offsets
                                                    (32, 160, 3, False, 2500),   # strided rulebook, two co tiles
                                                    (192, 64, 3, True, 1500)])   # widest ci on the path, two row slices
def test_wgrad_tc_kernel_on_emulator(cin, cout, ksize, subm, n):
    """msmd_spconv_bwd_weight_tc on the tcgen05 host model against the oracle's float64-accumulated
    gradient: compaction, raw-tile transposition into TMEM (A) / swizzled shared memory (B), 3xTF32
    accumulation across chunks and scans, partial tiles + fixed-order slice reduction."""
    L = tc_emu()
    shape, batch = [7, 16, 16], 2
    idx, feat = random_sparse(3, batch, shape, n, cin)
    rng = np.random.default_rng(4)
    ks = cpu._triple(ksize)
    if subm:
        pair = cpu.subm_rulebook(idx, shape, ks, 1)
    else:
        _, pair, _ = cpu.conv_rulebook(idx, shape, ks, 2, 1, 1)
    kvol, n_out = pair.shape
    go = rng.standard_normal((n_out, cout)).astype(np.float32)
    _, ref = cpu.spconv_bwd(feat, np.zeros((cout, *ks, cin), np.float32), pair, go, need_input_grad=False)
    assert L.emu_msmd_spconv_bwd_weight_tc_supported(cin, cout, kvol) == 1
    need = L.emu_msmd_spconv_bwd_weight_tc_workspace(n_out, cin, cout, kvol)
    assert need > 0 and need % (kvol * cin * cout * 4) == 0
    ws = np.full(need // 4, np.nan, np.float32)
    gw = np.full((cout, *ks, cin), np.nan, np.float32)
    st = L.emu_msmd_spconv_bwd_weight_tc(P(feat), feat.shape[0], P(go), P(pair), n_out, cin, cout, kvol, P(gw),
                                         P(ws), ctypes.c_size_t(need), None)
    assert st == 0, L.emu_last_error()
    assert rel(gw, ref) < 5e-6
    assert L.emu_msmd_spconv_bwd_weight_tc(P(feat), feat.shape[0], P(go), P(pair), n_out, cin, cout, kvol, P(gw),
                                           P(ws), ctypes.c_size_t(need - 4), None) == -3   # MSMD_ERR_WORKSPACE


def test_wgrad_tc_edge_cases_on_emulator():
    L = tc_emu()
    gw = np.full((4, 27, 8), np.nan, np.float32)
    assert L.emu_msmd_spconv_bwd_weight_tc(None, 0, None, None, 0, 8, 4, 27, P(gw), None, ctypes.c_size_t(0), None) == 0
    assert np.all(gw == 0)
    assert L.emu_msmd_spconv_bwd_weight_tc_supported(5, 16, 27) == 0      # the SIMT kernel keeps these
    assert L.emu_msmd_spconv_bwd_weight_tc_supported(260, 16, 27) == 0
    # an offset without any pair: its gradient slab is exactly zero
    idx, feat = random_sparse(9, 1, [5, 9, 9], 60, 8)
    pair = cpu.subm_rulebook(idx, [5, 9, 9], 3, 1)
    pair[3, :] = -1
    go = np.random.default_rng(0).standard_normal((60, 8)).astype(np.float32)
    need = L.emu_msmd_spconv_bwd_weight_tc_workspace(60, 8, 8, 27)
    ws = np.full(need // 4, np.nan, np.float32)
    gw = np.full((8, 3, 3, 3, 8), np.nan, np.float32)
    assert L.emu_msmd_spconv_bwd_weight_tc(P(feat), 60, P(go), P(pair), 60, 8, 8, 27, P(gw), P(ws),
                                           ctypes.c_size_t(need), None) == 0, L.emu_last_error()
    _, ref = cpu.spconv_bwd(feat, np.zeros_like(gw), pair, go, need_input_grad=False)
    assert rel(gw, ref) < 5e-6 and np.all(gw.reshape(8, 27, 8)[:, 3] == 0)


# --------------------------------------------------------------------------------------
# the real wrappers / modules / autograd over the EMULATED tensor-core kernels: the Python glue of the
# paths that have not run on hardware (precision modes, mask-sorted tables, tc dgrad, tc wgrad)
# --------------------------------------------------------------------------------------
@pytest.fixture()
def tc_ops_on_emulator(ops_on_emulator, monkeypatch):
    import _cpu_ops
    ops = ops_on_emulator
    monkeypatch.setattr(ops, 'tc_supported', lambda cout, kvol, cin: cout <= 256 and kvol <= 32)
    for name in ('grid_build', 'rulebook_subm', 'rulebook_conv'):
        monkeypatch.setattr(ops, name, _cpu_ops.STANDINS[name])
    yield ops
    ops.set_wgrad_tc(False)


def test_tc_wrappers_drive_emulated_kernels(tc_ops_on_emulator):
    import torch
    ops = tc_ops_on_emulator
    shape, cin, cout = [7, 14, 14], 16, 24
    idx, feat = random_sparse(9, 1, shape, 300, cin)
    rng = np.random.default_rng(10)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    ref = cpu.spconv_fwd(feat, w, pair)
    t = torch.from_numpy
    row_perm, pair_sorted = ops.rulebook_mask_sort(t(pair))
    for mode, tol in (('tf32x3', 5e-6), ('bf16x3', 2e-5), ('bf16x3c', 2e-5), ('bf16', 1e-2)):
        tcw = ops.pack_weight_tc(t(w), ops.TC_MODES[mode])
        assert tcw.mode == ops.TC_MODES[mode] and (tcw.packed.dtype == torch.int16) == (mode != 'tf32x3')
        out = ops.spconv_fwd(t(feat), tcw, t(pair))
        assert rel(out.numpy(), ref) < tol
        if mode == 'bf16x3c':   # the result carries its split image for the next convolution; re-use is by identity
            assert out._msmd_split[1].shape == (ref.shape[0], 2 * 24) and ops.split_bf16(out) is out._msmd_split[1]
        else:
            srt = ops.spconv_fwd_tc(t(feat), tcw, pair_sorted, row_perm=row_perm)
            assert rel(srt.numpy(), out.numpy()) < 5e-6
        # data gradient through the same precision: forward contraction with the mirrored weight
        go = rng.standard_normal(ref.shape).astype(np.float32)
        gi = ops.spconv_bwd_data(t(go), ops.pack_weight_tc(ops.transpose_weight(t(w), flip_k=True), tcw.mode), t(pair))
        ri, _ = cpu.spconv_bwd(feat, w, pair, go, need_weight_grad=False)
        assert rel(gi.numpy(), ri) < tol
    # weight gradient: SIMT by default, tensor-core kernel when switched on, same wrapper
    _, rw = cpu.spconv_bwd(feat, w, pair, go, need_input_grad=False)
    simt = ops.spconv_bwd_weight(t(feat), t(go), t(pair), w.shape)
    ops.set_wgrad_tc(True)
    tcg = ops.spconv_bwd_weight(t(feat), t(go), t(pair), w.shape)
    ops.set_wgrad_tc(False)
    assert rel(simt.numpy(), rw) < 1e-5 and rel(tcg.numpy(), rw) < 5e-6 and not np.array_equal(simt.numpy(), tcg.numpy())


@pytest.mark.parametrize('precision', ['bf16x3', 'bf16x3c', 'bf16'])
def test_modules_and_autograd_in_16bit_modes_on_emulated_kernels(tc_ops_on_emulator, monkeypatch, precision):
    """SubMConv3d / SparseConv3d with spconv.CONV_PRECISION set: inference path (fused epilogue, packed-weight
    cache keyed on the mode) and the autograd path (forward + dgrad in the forward's precision, fp32 wgrad)."""
    import torch
    from msmdfusion_b200 import spconv
    monkeypatch.setattr(spconv, 'CONV_PRECISION', precision)
    tol = 2e-5 if precision.startswith('bf16x3') else 1e-2
    shape, cin, cout = [7, 12, 12], 8, 16
    idx, feat = random_sparse(11, 1, shape, 250, cin)
    for cls, kw in ((spconv.SubMConv3d, dict(padding=1)), (spconv.SparseConv3d, dict(stride=2, padding=1))):
        torch.manual_seed(0)
        conv = cls(cin, cout, 3, bias=False, **kw)
        w = conv.weight.detach().numpy()
        if conv.subm:
            pair = cpu.subm_rulebook(idx, shape, 3, 1)
        else:
            _, pair, _ = cpu.conv_rulebook(idx, shape, 3, 2, 1, 1)
        ref = cpu.spconv_fwd(feat, w, pair)
        with torch.no_grad():
            out = conv(spconv.SparseConvTensor(torch.from_numpy(feat), torch.from_numpy(idx), shape, 1))
        assert conv.packed_weight().mode == tc_ops_on_emulator.TC_MODES[precision]
        assert rel(out.features.numpy(), ref) < tol
        monkeypatch.setattr(spconv, 'CONV_PRECISION', 'tf32x3')      # the cache follows the setting
        assert conv.packed_weight().mode == 1
        monkeypatch.setattr(spconv, 'CONV_PRECISION', precision)
        f = torch.from_numpy(feat).clone().requires_grad_(True)
        out = conv(spconv.SparseConvTensor(f, torch.from_numpy(idx), shape, 1))
        g = torch.randn(out.features.shape, generator=torch.Generator().manual_seed(1))
        (out.features * g).sum().backward()
        ri, rw = cpu.spconv_bwd(feat, w, pair, g.numpy())
        assert rel(out.features.detach().numpy(), ref) < tol
        assert rel(f.grad.numpy(), ri) < tol and rel(conv.weight.grad.numpy(), rw) < 1e-5


# --------------------------------------------------------------------------------------
# the spconv-2.x functional boundary (msmdfusion_b200/spconv_v2_api.py): the reference's OWN patched
# convolution module (bug_fix/conv.py), compiled from its source in place, runs on the two mirrored functions
# --------------------------------------------------------------------------------------
def _reference_sparse_convolution():
    import contextlib
    import math
    import sys
    import time
    from typing import List, Optional, Tuple, Union
    import torch
    from torch import nn
    from torch.nn import init
    from torch.nn.init import calculate_gain
    from torch.nn.parameter import Parameter
    from msmdfusion_b200 import spconv, spconv_v2_api as api
    from oracle.ref_inplace import load_def
    ns = dict(math=math, time=time, sys=sys, np=np, torch=torch, nn=nn, init=init, Parameter=Parameter,
              calculate_gain=calculate_gain, List=List, Optional=Optional, Tuple=Tuple, Union=Union,
              SparseModule=spconv.SparseModule, SparseConvTensor=spconv.SparseConvTensor, expand_nd=spconv.expand_nd,
              IndiceData=spconv.IndiceData, ImplicitGemmIndiceData=api.ImplicitGemmIndiceData, ConvAlgo=api.ConvAlgo,
              ops=api.ops, Fsp=api.Fsp, CPU_ONLY_BUILD=False, FILTER_HWIO=False, nullcontext=contextlib.nullcontext,
              spconv_save_debug_data=lambda *a, **k: None)
    return load_def('bug_fix/conv.py', 'SparseConvolution', ns, keyword='class')


@pytest.mark.skipif(not os.path.isdir('/root/reference/bug_fix'), reason='reference tree not mounted')
def test_reference_conv_module_runs_on_the_v2_functional_boundary(tc_ops_on_emulator):
    """bug_fix/conv.py:SparseConvolution (the module the reference tells users to copy over spconv's) with
    ``ops.get_indice_pairs_implicit_gemm`` / ``Fsp.implicit_gemm`` resolved to msmdfusion_b200.spconv_v2_api:
    forward (SubM with a shared indice_key, strided) and gradients equal this package's own modules / the oracle;
    the 9-tuple has spconv's structure."""
    import torch
    from msmdfusion_b200 import spconv, spconv_v2_api as api
    Ref = _reference_sparse_convolution()
    shape, cin, cout = [7, 12, 12], 8, 16
    idx, feat = random_sparse(21, 1, shape, 260, cin)
    ti, tf = torch.from_numpy(idx), torch.from_numpy(feat)
    # the tuple itself
    res = api.get_indice_pairs_implicit_gemm(ti, 1, shape, api.ConvAlgo.MaskImplicitGemm, [3, 3, 3], [2, 2, 2], [1, 1, 1],
                                             [1, 1, 1], [0, 0, 0], subm=False, transpose=False, is_train=True)
    oi, pair, _ = cpu.conv_rulebook(idx, shape, 3, 2, 1, 1)
    assert len(res) == 9 and np.array_equal(res[0].numpy(), oi) and np.array_equal(res[2].numpy(), pair)
    assert np.array_equal(res[1].numpy(), (pair >= 0).sum(1)) and np.array_equal(res[3].numpy(), cpu.pair_transpose(pair, 260))
    m = res[4][0].numpy().astype(np.int64) & 0xFFFFFFFF
    assert np.array_equal(m, ((pair >= 0).astype(np.int64) << np.arange(27)[:, None]).sum(0))
    assert np.array_equal(np.sort(res[6][0].numpy()), np.arange(pair.shape[1])) and res[8][0].dtype == np.uint32
    sub = api.get_indice_pairs_implicit_gemm(ti, 1, shape, api.ConvAlgo.MaskImplicitGemm, 3, 1, 1, 1, 0, subm=True,
                                             is_train=False)
    assert sub[0] is ti and sub[3].numel() == 0 and np.array_equal(sub[2].numpy(), cpu.subm_rulebook(idx, shape, 3, 1))
    # the reference module on top of it
    for subm in (True, False):
        torch.manual_seed(0)
        kw = dict(padding=1) if subm else dict(stride=2, padding=1)
        ref = Ref(3, cin, cout, 3, bias=False, subm=subm, indice_key='k' if subm else None, **kw)
        ours = (spconv.SubMConv3d if subm else spconv.SparseConv3d)(cin, cout, 3, bias=False, **kw)
        assert ref.weight.shape == ours.weight.shape
        with torch.no_grad():
            ours.weight.copy_(ref.weight)
            a = ref(spconv.SparseConvTensor(tf, ti, shape, 1))
            b = ours(spconv.SparseConvTensor(tf, ti, shape, 1))
            assert torch.equal(a.indices, b.indices) and a.spatial_shape == b.spatial_shape
            assert rel(a.features.numpy(), b.features.numpy()) < 1e-6
            if subm:   # the stored ImplicitGemmIndiceData is reused by a second layer with the same key
                ref2 = Ref(3, cout, cout, 3, bias=False, subm=True, indice_key='k', padding=1)
                c = ref2(a)
                assert c.indice_dict['k'] is a.indice_dict['k'] and c.features.shape == (260, cout)
        f = tf.clone().requires_grad_(True)
        out = ref(spconv.SparseConvTensor(f, ti, shape, 1))
        g = torch.randn(out.features.shape, generator=torch.Generator().manual_seed(1))
        (out.features * g).sum().backward()
        pr = cpu.subm_rulebook(idx, shape, 3, 1) if subm else pair
        ri, rw = cpu.spconv_bwd(feat, ref.weight.detach().numpy(), pr, g.numpy())
        assert rel(f.grad.numpy(), ri) < 1e-5 and rel(ref.weight.grad.numpy(), rw) < 1e-5


# --------------------------------------------------------------------------------------
# the NATIVE EXECUTOR (csrc/executor.cu) on the emulator, driven by the real Python plan builder: bit grids,
# rulebooks, mask sort, SIMT / tensor-core convolutions, arena, activation descriptors -- in the three
# configurations whose integration has not run on hardware (default is GPU-verified and calibrates)
# --------------------------------------------------------------------------------------
_EXEC = None


def exec_emu():
    global _EXEC
    if _EXEC is None:
        spec = importlib.util.spec_from_file_location('emul_build', os.path.join(HERE, 'tools', 'cuda_emul', 'build.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _EXEC = ctypes.CDLL(mod.build_exec())
        _EXEC.emu_last_error.restype = ctypes.c_char_p
    return _EXEC


class _ExecLib:
    """``_cabi.lib()`` look-alike over the emulated executor library (all units in one image)."""

    def __init__(self, cabi):
        self._cabi, self._cache = cabi, {}

    def msmd_last_error(self):
        return exec_emu().emu_last_error()

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            fn = getattr(exec_emu(), 'emu_' + name)
            fn.restype, fn.argtypes = self._cabi.SIGNATURES[name]
            self._cache[name] = fn
        return fn


@pytest.fixture()
def executor_on_emulator(monkeypatch):
    import torch
    from msmdfusion_b200 import _cabi, executor, ops
    shim = _ExecLib(_cabi)

    def ptr(t):
        if t is None:
            return None
        assert t.is_contiguous()
        return ctypes.c_void_p(t.data_ptr())

    class Scratch:
        def get(self, device, nbytes, slot='ws'):
            return torch.empty(max(int(nbytes), 16), dtype=torch.uint8)
    for mod in (_cabi, ops, executor):
        monkeypatch.setattr(mod, 'lib', lambda: shim)
        monkeypatch.setattr(mod, 'ptr', ptr)
        monkeypatch.setattr(mod, 'stream', lambda device=None: None)
    for mod in (_cabi, ops):
        monkeypatch.setattr(mod, 'scratch', Scratch())
    yield shim
    shim.msmd_spconv_set_mask_sort(0)


def test_native_executor_on_emulator(executor_on_emulator, monkeypatch):
    """A SparseEncoder (basic blocks: SubM chains with residuals, three strided convs, conv_out) through
    executor.SparseNetPlan -> msmd_sparse_net_forward, all kernels emulated, against the oracle's encoder:
    default (calibration: this configuration is GPU-verified), mask-sorted, bf16x3, bf16x3 + mask-sorted -- the last
    three are the integrations that have not run on hardware."""
    import torch
    import msmdfusion_b200 as m
    from msmdfusion_b200 import ops, spconv
    from oracle import model as omodel
    sys_path_fix = os.path.join(HERE)
    if sys_path_fix not in __import__('sys').path:
        __import__('sys').path.insert(0, sys_path_fix)
    from _fixtures import randomize_bn
    cfg = dict(type='SparseEncoder', in_channels=5, sparse_shape=[17, 48, 48], output_channels=32, order=('conv', 'norm', 'act'),
               encoder_channels=((16, 16, 32), (32, 32, 48), (48, 48, 112), (112, 112)),   # 112: variant 3 (N >= 96)
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')
    torch.manual_seed(0)
    enc = m.registry.build_middle_encoder(dict(cfg)).eval()
    randomize_bn(enc, 1)
    idx, feat = random_sparse(2, 2, [17, 48, 48], 600, 5)
    ref_sp, ref_feats, _ = omodel.sparse_encoder(enc.state_dict(), dict(cfg), feat, idx, 2)
    tf, ti = torch.from_numpy(feat), torch.from_numpy(idx)

    def run():
        enc._plan = None
        plan, marks = enc._plan_for()
        with torch.no_grad():
            acts = plan.run(tf, ti, enc.sparse_shape, 2)
        return [acts[i] for i in marks]

    for mask_sort, precision, tol in ((0, 'tf32x3', 1e-5), (1, 'tf32x3', 1e-5), (0, 'bf16x3', 1e-4), (1, 'bf16x3', 1e-4),
                                      (0, 'bf16x3c', 1e-4)):
        executor_on_emulator.msmd_spconv_set_mask_sort(mask_sort)
        monkeypatch.setattr(spconv, 'CONV_PRECISION', precision)
        outs = run()
        assert {L['weight_tc'] for L in enc._plan[1].layers} == {ops.TC_MODES[precision]}
        for (f, ix, shape), r in zip(outs[:-1], ref_feats):
            assert np.array_equal(ix.numpy(), r.indices) and list(shape) == list(r.spatial_shape)
            assert rel(f.numpy(), r.features) < tol, (mask_sort, precision)
        f, ix, shape = outs[-1]
        dense = cpu.dense(ix.numpy(), f.numpy(), shape, 2)
        assert rel(dense.reshape(ref_sp.shape), ref_sp) < tol


# --------------------------------------------------------------------------------------
# 16-bit modes, variant 3: A operand in TENSOR memory (two K elements per 32-bit column) + split-K pairs
# --------------------------------------------------------------------------------------
def _tc16_fwd_v3(feat, w, pair, x3, scale, shift, residual, relu, row_perm, variant, split):
    L = tc_emu()
    cout, cin = w.shape[0], w.shape[-1]
    kvol, n_out = pair.shape
    packed = np.full(L.emu_msmd_spconv_tc16_packed_bytes(cout, kvol, cin, x3) // 2, 0x7FC0, np.uint16)
    assert L.emu_msmd_spconv_tc16_pack_weight(P(w), cout, kvol, cin, x3, P(packed), None) == 0, L.emu_last_error()
    out = np.full((n_out, cout), np.nan, np.float32)
    assert L.emu_msmd_spconv_tc16_set_variant(variant) == 0
    try:
        need = L.emu_msmd_spconv_tc16_workspace(n_out, cout) if split else 0
        assert (need > 0) == bool(split)
        ws = np.zeros(need // 4 + 64, np.float32) if need else None
        st = L.emu_msmd_spconv_fwd_tc16_ws(P(feat), feat.shape[0], P(packed), P(pair), P(row_perm), n_out, cin, cout,
                                           kvol, x3, P(scale), P(shift), P(residual), relu, P(out), P(ws),
                                           ctypes.c_size_t(need), None)
    finally:
        L.emu_msmd_spconv_tc16_set_variant(2)
    assert st == 0, L.emu_last_error()
    return out


@pytest.mark.parametrize('x3,cin,cout,n,split', [(1, 16, 16, 300, False),    # three A stages (N <= 64), vector gather
                                                 (0, 5, 24, 200, False),     # scalar gather, padded N, bf16
                                                 (1, 20, 144, 150, False),   # N = 144, two A stages
                                                 (1, 32, 128, 200, True),    # split-K CTA pairs + hand-off
                                                 (0, 32, 128, 200, True)])
def test_tc16_variant3_on_emulator(x3, cin, cout, n, split):
    """spconv_fwd_tc16t_kernel on the host model (16-bit A operand in tensor memory: element 2c in the low half
    of column c): same results as variant 2 of the same mode; epilogue, mask-sorted table."""
    shape = [5, 12, 12]
    idx, feat = random_sparse(0, 1, shape, n, cin)
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    if x3:
        ref, tol = cpu.spconv_fwd(feat, w, pair), 2e-5
    else:
        ref, tol = cpu.spconv_fwd(bf16_round(feat), bf16_round(w), pair), 2e-6
    got = tc16_fwd(feat, w, pair, x3, variant=3, split=split)
    assert rel(got, ref) < tol
    assert rel(got, tc16_fwd(feat, w, pair, x3)) < 2e-6       # variant 2 of the same mode
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal(ref.shape).astype(np.float32)
    perm = rng.permutation(n).astype(np.int32)      # any permutation is a valid slot -> row map
    got = tc16_fwd(feat, w, np.ascontiguousarray(pair[:, perm]), x3, scale, shift, res, 1, row_perm=perm, variant=3,
                   split=split)
    assert rel(got, np.maximum(ref * scale + shift + res, 0)) < tol


@pytest.mark.parametrize('x3', [0, 1])
@pytest.mark.parametrize('cin,cout,n', [(16, 16, 300),    # 7 chunks: the last stage holds one chunk block only
                                        (20, 80, 150)])   # three-MMA x3 mode (2N > 256 is not needed: cat off at N > 128)
def test_tc16_two_chunk_blocks_per_stage_on_emulator(x3, cin, cout, n):
    """A/B switch [3] = 2 (MSMD_TC_TUNE=cps=2): a pipeline stage of the 16-bit kernel holds two chunk blocks --
    half the mbarrier round trips per K element; bit-identical to one block per stage (same MMA order)."""
    L = tc_emu()
    shape = [5, 12, 12]
    idx, feat = random_sparse(0, 1, shape, n, cin)
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    pair[3:9, : n // 2] = -1          # some tiles skip chunks: odd / even active counts both occur
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal((n, cout)).astype(np.float32)
    one = tc16_fwd(feat, w, pair, x3, scale, shift, res, 1)
    assert L.emu_msmd_spconv_tc_set_tuning(3, 2) == 0
    try:
        two = tc16_fwd(feat, w, pair, x3, scale, shift, res, 1)
    finally:
        L.emu_msmd_spconv_tc_set_tuning(3, 0)
    assert np.array_equal(one, two)
    ref = cpu.spconv_fwd(feat, w, pair) if x3 else cpu.spconv_fwd(bf16_round(feat), bf16_round(w), pair)
    assert rel(two, np.maximum(ref * scale + shift + res, 0)) < (2e-5 if x3 else 2e-6)


# --------------------------------------------------------------------------------------
# bf16x3 through the split-bf16 operand cache (csrc/spconv_sb.cu): cp.async gather of pre-split activations
# --------------------------------------------------------------------------------------
def bf16_split(x):
    hi = bf16_round(x)
    return hi, bf16_round(np.ascontiguousarray(x, np.float32) - hi)


def sb_split(feat):
    L = tc_emu()
    n, c = feat.shape
    width = L.emu_msmd_split_width(c)
    xs = np.full((max(n, 1), width), 0x7FC0, np.uint16)[:n]   # NaN fill: every element must be written
    assert L.emu_msmd_split_bf16(P(feat), n, c, P(xs), None) == 0, L.emu_last_error()
    return xs


def sb_fwd(feat, w, pair, scale=None, shift=None, residual=None, relu=0, want_out=True, want_split=True, xs=None):
    L = tc_emu()
    cout, cin = w.shape[0], w.shape[-1]
    kvol, n_out = pair.shape
    packed = np.full(L.emu_msmd_spconv_sb_packed_bytes(cout, kvol, cin) // 2, 0x7FC0, np.uint16)
    assert L.emu_msmd_spconv_sb_pack_weight(P(w), cout, kvol, cin, P(packed), None) == 0, L.emu_last_error()
    xs = sb_split(feat) if xs is None else xs
    out = np.full((n_out, cout), np.nan, np.float32) if want_out else None
    out_s = np.full((n_out, L.emu_msmd_split_width(cout)), 0x7FC0, np.uint16) if want_split else None
    st = L.emu_msmd_spconv_fwd_sb(P(xs), feat.shape[0], P(packed), P(pair), n_out, cin, cout, kvol, P(scale), P(shift),
                                  P(residual), relu, P(out), P(out_s), None)
    assert st == 0, L.emu_last_error()
    return out, out_s


@pytest.fixture(params=['tile_per_cta', 'persistent', 'persistent_two_per_sm'])
def sb_schedule(request):
    """The three schedules of csrc/spconv_sb.cu: one tile per CTA (r02c kernel), the persistent work-balanced kernel
    with one CTA per SM (8 epilogue warps), and its two-CTAs-per-SM instantiation (4 epilogue warps).  The small
    test shapes give the persistent kernel ranges of ~4 chunks, so nearly every tile is split over 2-3 CTAs."""
    L = tc_emu()
    assert L.emu_msmd_spconv_sb_set_variant({'tile_per_cta': 1}.get(request.param, 2)) == 0
    assert L.emu_msmd_spconv_tc_set_tuning(0, 2 if request.param == 'persistent_two_per_sm' else 1) == 0
    try:
        yield request.param
    finally:
        L.emu_msmd_spconv_sb_set_variant(0)
        L.emu_msmd_spconv_tc_set_tuning(0, 0)


def split_to_float(xs, c):
    """[hi | lo] bf16 image -> hi + lo as fp32 (first c channels) and the padding channels."""
    c8 = xs.shape[1] // 2
    f = (xs.astype(np.uint32) << 16).view(np.float32)
    return f[:, :c] + f[:, c8:c8 + c], f[:, c:c8], f[:, c8 + c:]


def test_split_bf16_image_on_emulator():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((37, 5)) * 3).astype(np.float32)
    xs = sb_split(x)
    assert xs.shape == (37, 16)
    hi, lo = bf16_split(x)
    f = (xs.astype(np.uint32) << 16).view(np.float32)
    assert np.array_equal(f[:, :5], hi) and np.array_equal(f[:, 8:13], lo)
    assert not f[:, 5:8].any() and not f[:, 13:].any()          # channel padding is zero
    assert np.abs(f[:, :5] + f[:, 8:13] - x).max() < 2.0 ** -16 * np.abs(x).max()


@pytest.mark.parametrize('cin,cout,n', [(16, 16, 300),    # four kernel offsets per 64-element chunk, concatenated-B
                                        (5, 16, 200),     # input channels padded 5 -> 8: eight offsets per chunk
                                        (24, 144, 150),   # chunk boundaries inside a kernel offset, padded N, 3-MMA mode
                                        (64, 128, 140),   # one offset per chunk, concatenated-B at 2N = 256
                                        (80, 96, 260),    # fusion-encoder widths, ragged last tile, several tiles
                                        (16, 20, 150)])   # Cout % 8 == 4: zero padding of the split image from the row-contiguous epilogue
def test_sb_kernel_on_emulator(cin, cout, n, sb_schedule):
    """csrc/spconv_sb.cu on the tcgen05 model with late-as-possible asynchronous copies: the result equals the
    bf16x3 arithmetic (within 2e-5 of the fp32 oracle), the split image the epilogue writes IS the split of the fp32
    result it writes (bit for bit, padding channels zero), fused epilogue, fp32-only and split-only outputs."""
    shape = [5, 12, 12]
    idx, feat = random_sparse(0, 1, shape, n, cin)
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    ref = cpu.spconv_fwd(feat, w, pair)
    out, out_s = sb_fwd(feat, w, pair)
    assert rel(out, ref) < 2e-5
    hi, lo = bf16_split(out)
    f = (out_s.astype(np.uint32) << 16).view(np.float32)
    c8 = out_s.shape[1] // 2
    assert np.array_equal(f[:, :cout], hi) and np.array_equal(f[:, c8:c8 + cout], lo)
    assert not f[:, cout:c8].any() and not f[:, c8 + cout:].any()
    # it is the bf16x3 arithmetic of spconv_tc16.cu, operand for operand
    assert rel(out, tc16_fwd(feat, w, pair, 1)) < 2e-6
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal(ref.shape).astype(np.float32)
    want = np.maximum(ref * scale + shift + res, 0)
    got, _ = sb_fwd(feat, w, pair, scale, shift, res, 1, want_split=False)
    assert rel(got, want) < 2e-5
    _, only_s = sb_fwd(feat, w, pair, scale, shift, res, 1, want_out=False)
    assert np.array_equal(only_s, sb_split(got))


def test_sb_kernel_chain_strided_rulebook_and_empty_tiles_on_emulator(sb_schedule):
    """Two layers chained through the split image only (the second never sees fp32 activations), a strided
    rulebook, and a tile without any pair."""
    shape = [7, 14, 14]
    idx, feat = random_sparse(5, 1, shape, 400, 16)
    rng = np.random.default_rng(6)
    w1 = (rng.standard_normal((32, 3, 3, 3, 16)) * 0.2).astype(np.float32)
    w2 = (rng.standard_normal((32, 3, 3, 3, 32)) * 0.1).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, 3, 1)
    y1, y1_s = sb_fwd(feat, w1, pair, relu=1)
    y2, _ = sb_fwd(y1, w2, pair, xs=y1_s)
    ref1 = np.maximum(cpu.spconv_fwd(feat, w1, pair), 0)
    assert rel(y1, ref1) < 2e-5 and rel(y2, cpu.spconv_fwd(ref1, w2, pair)) < 4e-5
    _, spair, _ = cpu.conv_rulebook(idx, shape, (3, 3, 3), 2, 1, 1)
    got, _ = sb_fwd(feat, w1, spair)
    assert rel(got, cpu.spconv_fwd(feat, w1, spair)) < 2e-5
    empty = np.full((27, 140), -1, np.int32)
    empty[13, 130] = 7
    one = np.ones(32, np.float32)
    shift = rng.standard_normal(32).astype(np.float32)
    got, _ = sb_fwd(feat, w1, empty, one, shift)
    assert rel(got, cpu.spconv_fwd(feat, w1, empty) + shift) < 2e-5


def sb_fwd_ex(feat, w, pair, row_perm=None, masks=True, scale=None, shift=None, residual=None, relu=0):
    """msmd_spconv_fwd_sb_ex: the persistent kernel with tile masks (weighted work shares, skipped chunks) and / or a
    mask-sorted pair table."""
    L = tc_emu()
    cout, cin = w.shape[0], w.shape[-1]
    kvol, n_out = pair.shape
    packed = np.full(L.emu_msmd_spconv_sb_packed_bytes(cout, kvol, cin) // 2, 0x7FC0, np.uint16)
    assert L.emu_msmd_spconv_sb_pack_weight(P(w), cout, kvol, cin, P(packed), None) == 0, L.emu_last_error()
    xs = sb_split(feat)
    tm = None
    if masks:
        tm = np.full((n_out + 127) // 128, 0xFFFFFFFF, np.uint32)
        assert L.emu_msmd_rulebook_tile_masks(P(pair), kvol, n_out, P(tm), None) == 0, L.emu_last_error()
        used = np.concatenate([pair >= 0, np.zeros((kvol, len(tm) * 128 - n_out), bool)], 1).reshape(kvol, len(tm), 128).any(2)
        assert np.array_equal(tm, (used.astype(np.uint64) << np.arange(kvol, dtype=np.uint64)[:, None]).sum(0).astype(np.uint32))
    out = np.full((n_out, cout), np.nan, np.float32)
    out_s = np.full((n_out, L.emu_msmd_split_width(cout)), 0x7FC0, np.uint16)
    st = L.emu_msmd_spconv_fwd_sb_ex(P(xs), feat.shape[0], P(packed), P(pair), P(row_perm), P(tm), n_out, cin, cout, kvol,
                                     P(scale), P(shift), P(residual), relu, P(out), P(out_s), None)
    assert st == 0, L.emu_last_error()
    return out, out_s


@pytest.mark.parametrize('cin,cout', [(16, 32), (64, 64), (80, 96)])
def test_sb_persistent_kernel_with_tile_masks_and_mask_sorted_rows_on_emulator(cin, cout):
    """The persistent schedule with its two side tables, on a LiDAR-like index set (many tiles miss most kernel
    offsets): tile masks only (work shares counted in chunks that have a pair, the others skipped), and tile masks
    of a mask-sorted table with the row permutation; a table with tiles that have no pair at all."""
    from msmdfusion_b200 import synthetic
    pts = synthetic.lidar_scene(seed=4, sweeps=1)[:2400]
    _, c, _ = cpu.hard_voxelize(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    idx = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], 1)[:900]
    n = idx.shape[0]
    rng = np.random.default_rng(2)
    feat = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) * 0.2).astype(np.float32)
    pair = cpu.subm_rulebook(idx, [41, 1440, 1440], 3, 1)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal((n, cout)).astype(np.float32)
    want = np.maximum(cpu.spconv_fwd(feat, w, pair) * scale + shift + res, 0)
    plain, plain_s = sb_fwd(feat, w, pair, scale, shift, res, 1)
    assert rel(plain, want) < 2e-5
    got, got_s = sb_fwd_ex(feat, w, pair, None, True, scale, shift, res, 1)
    assert rel(got, want) < 2e-5 and rel(got, plain) < 2e-6
    assert np.array_equal(got_s, sb_split(got))
    E = emu()
    E.emu_msmd_rulebook_mask_sort_workspace.restype = ctypes.c_size_t
    need = E.emu_msmd_rulebook_mask_sort_workspace(n)
    ws = np.zeros(need, np.uint8)
    perm = np.full(n, -1, np.int32)
    pair_sorted = np.full_like(pair, -9)
    ok(E.emu_msmd_rulebook_mask_sort(P(pair), 27, n, P(perm), P(pair_sorted), P(ws), ctypes.c_size_t(need), None))
    for masks in (True, False):
        got, got_s = sb_fwd_ex(feat, w, pair_sorted, perm, masks, scale, shift, res, 1)
        assert rel(got, want) < 2e-5 and rel(got, plain) < 2e-6
        assert np.array_equal(got_s, sb_split(got))
    empty = np.full((27, 700), -1, np.int32)     # tiles 0, 2, 3 and 5 have no pair: one unit each, epilogue only
    empty[13, 130] = 7
    empty[2, 600:640] = np.arange(40)
    got, _ = sb_fwd_ex(feat, w, empty, None, True, scale, shift)
    assert rel(got, cpu.spconv_fwd(feat, w, empty) * scale + shift) < 2e-5


# --------------------------------------------------------------------------------------
# one Gated Modality-Aware stage in one C-ABI call (csrc/gma.cu) on the emulated image
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', ['tf32x3', 'bf16x3c'])
def test_native_gma_stage_on_emulator(executor_on_emulator, monkeypatch, precision):
    """msmd_gma_stage_forward (row gather, only-3D chain, fused gates + concatenation, aggregation block, sparse_add
    with the previous stage, strided downscale conv -- all carved from one arena) against the Python module path of
    the same encoder (eager nn.Linear gates, index_select / cat, one executor call per chain), stage after stage:
    identical index sets, features within fp32 rounding; including a stage with NO only-2D voxel and NO mixed pair
    (the all-zero padding voxels) and unassigned only-2D voxels (nn_idx = -1 -> the dummy embedding)."""
    import sys
    import torch
    sys.path.insert(0, HERE)
    import _cpu_ops
    from test_train_host import _encoder_inputs
    from msmdfusion_b200 import functional as Fsp
    from msmdfusion_b200 import fusion_encoder as fe
    from msmdfusion_b200 import ops, spconv
    from _fixtures import randomize_bn
    monkeypatch.setattr(spconv, 'CONV_PRECISION', precision)
    monkeypatch.setattr(ops, 'tc_supported', lambda cout, kvol, cin: cout <= 256 and kvol <= 32)
    torch.manual_seed(12)
    enc = fe.SparseMultiModalEncoderPaint(in_channels_3D=(4, 8, 8, 8), in_channels_2D=(64,) * 4,
                                          out_channels=(8, 8, 8, 8), padding=(1, 1, [0, 1, 1], 0)).eval()
    randomize_bn(enc, 3)
    v3l, v2l, s3l, s2l = _encoder_inputs(31, 1)
    rng = np.random.default_rng(5)
    tol = 1e-5 if precision == 'tf32x3' else 1e-4
    prev_native = prev_python = None
    for stage_id in range(4):
        (i3, f3, shape), (i2, f2, _) = v3l[stage_id], v2l[stage_id]
        syn3, syn2 = s3l[stage_id], s2l[stage_id]
        if stage_id == 2:          # no only-2D voxel, no mixed pair: both groups are the all-zero padding voxel
            keep = torch.ones(i2.shape[0], dtype=torch.bool)
            keep[:] = False
            keep[syn2] = True
            i2, f2 = i2[keep], f2[keep]
            i3 = i3.clone(); i3[:, 1] = 0
            i2 = i2.clone(); i2[:, 1] = 1   # every 2-D voxel is "mixed" by flag, but the pair lists are empty below
            syn3, syn2 = syn3[:0], syn2[:0]
        v3 = spconv.SparseConvTensor(f3.clone(), i3.clone(), shape, 1)
        v2 = spconv.SparseConvTensor(f2.clone(), i2.clone(), shape, 1)
        for t in (v3, v2):
            t._mix = t.indices[:, 1].contiguous().int()
            t._bzyx = t.indices[:, [0, 2, 3, 4]].contiguous()
        only3 = torch.nonzero(v3._mix == 0).flatten()
        only2 = torch.nonzero(v2._mix == 0).flatten()
        if stage_id == 2:
            only2 = only2[:0]
        if only2.shape[0]:
            only2_rows, only2_bzyx = only2, v2._bzyx.index_select(0, only2)
        else:
            only2_rows, only2_bzyx = None, torch.zeros((1, 4), dtype=torch.int32)
        nn_idx = torch.from_numpy(rng.integers(-1, f3.shape[0], only2_bzyx.shape[0])).long()
        assign = dict(only3_rows=only3, only2_rows=only2_rows, only2_bzyx=only2_bzyx, nn_idx=nn_idx)
        rec = enc._stage_plan(stage_id)
        assert rec is not None
        with torch.no_grad():
            torch.manual_seed(100 + stage_id)
            got = enc._stage_native(rec, v3, v2, syn3, syn2, assign, stage_id, prev_native)
            torch.manual_seed(100 + stage_id)
            enc.fused_gates = False
            try:
                out = enc._grouped_sparse_conv_b1(v3, v2, syn3, syn2, stage_id, 6, 6, 20, 13.3, assign=assign)
            finally:
                enc.fused_gates = True
            if prev_python is not None:
                out = Fsp.sparse_add(out, prev_python)
            want = enc._run_chain(('down', stage_id), getattr(enc.downscale_blocks, f'stage_{stage_id + 1}'), out)
        assert got.spatial_shape == want.spatial_shape
        assert np.array_equal(got.indices.numpy(), want.indices.numpy()), stage_id
        assert rel(got.features.numpy(), want.features.numpy()) < tol, stage_id
        prev_native, prev_python = got, want
