"""CPU tests of the host-side mirror: registries, config loading, module/state-dict layout,
and that the C-ABI library loads and exports every symbol include/*.h declares.  No compute
call is made (there is no GPU here)."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest
import torch

import msmdfusion_b200 as m
from msmdfusion_b200 import _cabi, registry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        names |= set(re.findall(r'\b(msmd_[a-z0-9_]+)\s*\(', open(h).read()))
    return sorted(names)


def test_cabi_exports_every_declared_symbol():
    from msmdfusion_b200 import build
    build.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/ but not exported'
        assert s in _cabi.SIGNATURES, f'{s} has no ctypes signature in _cabi.py'
    assert set(_cabi.SIGNATURES) <= set(syms)
    L = _cabi.lib()
    assert L.msmd_abi_version() == 1
    assert L.msmd_scan_workspace() > 0 and L.msmd_hard_voxelize_workspace(1000) > 0


def test_no_cpu_fallback():
    """Ops must refuse CPU tensors instead of silently computing elsewhere."""
    from msmdfusion_b200 import ops
    with pytest.raises(RuntimeError):
        ops.hard_voxelize(torch.zeros(10, 5), [1, 1, 1], [0, 0, 0, 4, 4, 4], 5, 100)
    with pytest.raises(RuntimeError):
        m.hard_voxelize(torch.zeros(10, 5), None, None, None, [1, 1, 1], [0, 0, 0, 4, 4, 4], 5, 100)


def test_product_never_imports_oracle():
    """Only tests/ (incl. tests/tools), __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for f in glob.glob(os.path.join(ROOT, 'msmdfusion_b200', '*.py')) + glob.glob(os.path.join(ROOT, 'tools', '*.py')):
        src = open(f).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f
        assert 'oracle.' not in src.replace('oracle/', ''), f


def test_registries_and_builders():
    assert 'HardSimpleVFE' in registry.VOXEL_ENCODERS
    assert 'SparseEncoder' in registry.MIDDLE_ENCODERS
    assert registry.FUSION_LAYERS.name == 'fusion_layer'
    assert 'SubMConv3d' in registry.CONV_LAYERS and 'SparseConv3d' in registry.CONV_LAYERS
    conv = registry.build_conv_layer(dict(type='SubMConv3d', indice_key='k'), 4, 8, 3, padding=1, bias=False)
    assert tuple(conv.weight.shape) == (8, 3, 3, 3, 4) and conv.bias is None and conv.subm
    conv = registry.build_conv_layer(dict(type='SparseConv3d', indice_key='d'), 4, 8, (3, 1, 1),
                                     stride=(2, 1, 1), padding=0, bias=False)
    assert tuple(conv.weight.shape) == (8, 3, 1, 1, 4) and not conv.subm
    name, bn = registry.build_norm_layer(dict(type='BN1d', eps=1e-3, momentum=0.01), 8)
    assert isinstance(bn, torch.nn.BatchNorm1d) and bn.eps == 1e-3 and name == 'bn'
    with pytest.raises(KeyError):
        registry.build_middle_encoder(dict(type='NoSuchEncoder'))


def test_hotpath_config_builds_encoder_with_reference_state_dict_layout():
    cfg = m.Config.fromfile(os.path.join(ROOT, 'configs', 'msmd_lc_hotpath.py'))
    enc = registry.build_middle_encoder(cfg.hotpath.pts_middle_encoder)
    sd = enc.state_dict()
    # key layout + KRSC weight shapes
    assert tuple(sd['conv_input.0.weight'].shape) == (16, 3, 3, 3, 5)
    assert tuple(sd['encoder_layers.encoder_layer1.0.conv1.weight'].shape) == (16, 3, 3, 3, 16)
    assert tuple(sd['encoder_layers.encoder_layer1.2.0.weight'].shape) == (32, 3, 3, 3, 16)
    assert tuple(sd['encoder_layers.encoder_layer3.2.0.weight'].shape) == (128, 3, 3, 3, 64)
    assert tuple(sd['encoder_layers.encoder_layer4.1.conv2.weight'].shape) == (128, 3, 3, 3, 128)
    assert tuple(sd['conv_out.0.weight'].shape) == (128, 3, 1, 1, 128)
    assert 'encoder_layers.encoder_layer2.1.bn2.running_var' in sd
    n_conv = sum(1 for k in sd if k.endswith('weight') and sd[k].dim() == 5)
    assert n_conv == 21
    assert enc.encoder_layers.encoder_layer3[2][0].padding == [0, 1, 1]
    vfe = registry.build_voxel_encoder(cfg.hotpath.pts_voxel_encoder)
    assert vfe.num_features == 5
    vl = m.Voxelization(**cfg.hotpath.pts_voxel_layer)
    assert vl.max_voxels == (120000, 160000) and vl.grid_size.tolist() == [1440, 1440, 40]


@pytest.mark.skipif(not os.path.isdir('/root/reference/configs'), reason='reference tree not mounted')
def test_reference_config_files_load_unchanged():
    for name in ('MSMDFusion_nusc_voxel_LC.py', 'transfusion_nusc_voxel_L.py'):
        cfg = m.Config.fromfile(os.path.join('/root/reference/configs', name))
        enc = registry.build_middle_encoder(cfg.model.pts_middle_encoder)
        assert enc.sparse_shape == [41, 1440, 1440]
        registry.build_voxel_encoder(cfg.model.pts_voxel_encoder)
        m.Voxelization(**cfg.model.pts_voxel_layer)


def test_hard_simple_vfe_matches_reference_formula():
    vfe = m.HardSimpleVFE(num_features=4)
    feats = torch.rand(7, 10, 5)
    num = torch.randint(1, 10, (7,), dtype=torch.int32)
    out = vfe(feats, num, None)
    assert out.shape == (7, 4)
    assert torch.allclose(out, feats[:, :, :4].sum(1) / num.float().view(-1, 1))


def test_sparse_tensor_surface():
    sp = m.spconv
    f = torch.rand(5, 3)
    idx = torch.zeros(5, 4, dtype=torch.int32)
    t = sp.SparseConvTensor(f, idx, [4, 4, 4], 1)
    t2 = t.replace_feature(f * 2)
    assert t2.indices is t.indices and t2.features is not t.features
    t.indices = torch.zeros(5, 5, dtype=torch.int32)  # MSMDFusion.py:322-323 assigns 5 columns
    assert t.indices.shape[1] == 5
    assert t.find_indice_pair('nope') is None and t.find_indice_pair(None) is None
    seq = sp.SparseSequential(torch.nn.ReLU())
    assert len(seq) == 1 and isinstance(seq[0], torch.nn.ReLU)
    with pytest.raises(ValueError):
        sp._iset_of(t)


def test_executor_plan_matches_reference_layer_table(monkeypatch):
    """The native-executor plan derived from the module tree is the 21-layer table of
    SURVEY Appendix A (mmdet3d/models/middle_encoders/sparse_encoder.py:60-209): SubM/strided
    kinds, channel widths, geometry, the residual links of SparseBasicBlock and the activations
    exported as encode_features."""
    import torch
    import msmdfusion_b200 as m
    from msmdfusion_b200 import executor, registry
    from msmdfusion_b200 import spconv as sp
    cfg = m.Config.fromfile(os.path.join(ROOT, 'configs', 'msmd_lc_hotpath.py')).hotpath
    enc = registry.build_middle_encoder(cfg.pts_middle_encoder).eval()
    # kernel-layout packing needs the CUDA library; the plan STRUCTURE does not
    monkeypatch.setattr(sp.SparseConvolution, 'packed_weight', lambda self: self.weight.detach().reshape(-1))
    plan = executor.SparseNetPlan()
    cur = plan.add(enc.conv_input, 0)
    marks = [cur]
    for layer in enc.encoder_layers:
        cur = plan.add(layer, cur)
        marks.append(cur)
    marks.append(plan.add(enc.conv_out, cur))
    L = plan.layers
    assert len(L) == 21 and marks == [1, 6, 11, 16, 20, 21]
    assert [(x['cin'], x['cout']) for x in L] == (
        [(5, 16)] + [(16, 16)] * 4 + [(16, 32)] + [(32, 32)] * 4 + [(32, 64)] + [(64, 64)] * 4 +
        [(64, 128)] + [(128, 128)] * 4 + [(128, 128)])
    assert [x['subm'] for x in L] == [1] * 5 + [0] + [1] * 4 + [0] + [1] * 4 + [0] + [1] * 4 + [0]
    assert L[5]['stride'] == [2, 2, 2] and L[5]['padding'] == [1, 1, 1]
    assert L[15]['padding'] == [0, 1, 1]                      # encoder_paddings ((0,0,[0,1,1]))
    assert L[20]['ksize'] == [3, 1, 1] and L[20]['stride'] == [2, 1, 1] and L[20]['padding'] == [0, 0, 0]
    # SparseBasicBlock: conv2 adds the block input before the ReLU
    assert [x['residual'] for x in L[1:5]] == [-1, 1, -1, 3]
    assert all(x['relu'] == 1 for x in L) and all(x['scale'] is not None for x in L)
    enc.train()
    with pytest.raises(executor.Unsupported):
        executor.SparseNetPlan().add(enc.conv_input, 0)     # training-mode BN is not foldable


def test_packed_foreground_layout_cpu():
    """One packed upload per batch: order = samples, cameras, points (MSMDFusion.py:189-226)."""
    from msmdfusion_b200 import synthetic
    from msmdfusion_b200.detector import PackedForeground
    pts = synthetic.lidar_scene(0, 1)[:5000]
    metas = [synthetic.camera_scene(0, pts, virtual_per_camera=50, real_per_camera=20, empty_cameras=(2,)),
             synthetic.camera_scene(1, pts, virtual_per_camera=30, real_per_camera=10)]
    pk = PackedForeground(metas, 'cpu')
    assert pk.counts == [250, 180] and pk.ncam == 6
    assert pk.pixels.shape == (430, 3) and pk.points.shape == (430, 15) and pk.lidar2img.shape == (12, 16)
    cam = pk.cam.numpy()
    assert (np.diff(cam) >= 0).all() and set(cam[:250].tolist()) == {0, 1, 3, 4, 5}   # camera 2 is empty
    assert cam[250] == 6
    first = metas[1]['foreground2D_info']['fg_points'][0]
    assert np.array_equal(pk.points[250:280].numpy(), first)
    assert np.allclose(pk.lidar2img[7].numpy(), np.asarray(metas[1]['lidar2img'][1], np.float32).reshape(16))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) prints ONE JSON line
    with the contract's keys and `"impl": "reference"`; non-zero ranks exit silently."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS='4')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['unit'] == 'scenes/s' and line['value'] > 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'scenes/s', 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    silent = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                            capture_output=True, text=True, timeout=120, env=dict(env, RANK='1', WORLD_SIZE='2'),
                            cwd=ROOT)
    assert silent.returncode == 0 and silent.stdout.strip() == ''


def test_bench_step_defaults_depend_on_the_arm(monkeypatch):
    """Without --steps / --warmup: 100 + 30 device steps for this implementation, 3 + 1 whole-scene host passes for
    --impl reference (5-7 s each); explicit flags are taken as given in both arms."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location('bench_for_test', os.path.join(ROOT, 'bench.py'))
    B = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(B)
    for argv, want in ((['bench.py'], (100, 30)), (['bench.py', '--impl', 'reference'], (3, 1)),
                       (['bench.py', '--impl', 'reference', '--steps', '20', '--warmup', '5'], (20, 5)),
                       (['bench.py', '--steps', '7'], (7, 30))):
        monkeypatch.setattr(sys, 'argv', argv)
        a = B.parse()
        assert (a.steps, a.warmup) == want and a.workload == 'LC'


def test_tc_trace_summary_on_fabricated_record():
    """tools/tc_trace.py (GPU debug tool): the record layout of csrc/tc_trace.cuh and the per-chunk
    period / wait-share arithmetic, on a fabricated timeline (2000 cycles per chunk at 2000 MHz)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('tc_trace', os.path.join(ROOT, 'tools', 'tc_trace.py'))
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)
    src = open(os.path.join(ROOT, 'msmdfusion_b200', 'csrc', 'tc_trace.cuh')).read()
    assert 'kTrCtas = %d, kTrRoles = %d, kTrIts = %d, kTrPhases = %d, kTrHead = %d' % (
        T.CTAS, T.ROLES, T.ITS, T.PHASES, T.HEAD) in src
    words = T.HEAD + T.ROLES * T.ITS * T.PHASES
    rec = np.zeros((T.CTAS, words), np.uint64)
    r = rec[3]
    r[0], r[1], r[2], r[4], r[5], r[6], r[7], r[8], r[9], r[10] = 1000, 3000, 23000, 25000, 26000, 5, 10, 42, 10**6, 10**6 + 14000
    ev = r[T.HEAD:].reshape(T.ROLES, T.ITS, T.PHASES)
    for it in range(10):
        t0 = 3000 + 2000 * it
        ev[0, it] = [t0, t0 + 200, t0 + 900, 0]
        ev[1, it] = [t0, t0 + 300, t0 + 950, 0]
        ev[2, it] = [t0, t0 + 100, 0, 0]
        ev[3, it] = [t0 + 500, t0 + 1500, t0 + 1800, 0]
    (d,) = T.summarize(rec, 2000.0)
    assert d['cta'] == 42 and d['n_act'] == 10 and abs(d['period_us'] - 1.0) < 1e-9
    assert abs(d['mma_wait'] - 0.5) < 1e-9 and abs(d['prod0_wait'] - 0.1) < 1e-9 and abs(d['prod0_fill'] - 0.35) < 1e-9
    assert abs(d['b_lead_us'] - 0.2) < 1e-9 and abs(d['setup_us'] - 1.0) < 1e-9 and abs(d['wall_us'] - 14.0) < 1e-9
    assert 'mma_wait_weights' not in d


def test_quick_gpu_check_restatements_agree_with_the_oracle():
    """tools/quick_gpu_check.py (torch-free GPU smoke): its numpy restatements of the rulebook, the convolution,
    the weight gradient and the mirrored-weight data gradient against the oracle."""
    import importlib.util
    from oracle import cpu
    spec = importlib.util.spec_from_file_location('quick_gpu_check', os.path.join(ROOT, 'tools', 'quick_gpu_check.py'))
    Q = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(Q)
    rng = np.random.default_rng(0)
    shape, n, cin, cout = [9, 40, 40], 800, 8, 12
    lin = rng.choice(9 * 1600, size=n, replace=False)
    idx = np.stack([lin * 0, lin // 1600, (lin // 40) % 40, lin % 40], 1).astype(np.int32)
    pair = Q.subm_pairs(idx, shape)
    assert np.array_equal(pair, cpu.subm_rulebook(idx, shape, 3, 1))
    feat = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, 27, cin)) * 0.2).astype(np.float32)
    go = rng.standard_normal((n, cout)).astype(np.float32)
    w5 = w.reshape(cout, 3, 3, 3, cin)
    assert Q.rel(cpu.spconv_fwd(feat, w5, pair), Q.conv_ref(feat, w, pair)) < 1e-6
    gi, gw = cpu.spconv_bwd(feat, w5, pair, go)
    assert Q.rel(gw.reshape(cout, 27, cin), Q.wgrad_ref(feat, go, pair, cout, cin)) < 1e-6
    wt = np.ascontiguousarray(w[:, ::-1, :].transpose(2, 1, 0))
    assert Q.rel(gi, Q.conv_ref(go, wt, pair)) < 1e-6
    assert np.array_equal(Q.bf16_round(np.float32([1.0, 1.00390625, 3.1415927])), np.float32([1.0, 1.0, 3.140625]))
