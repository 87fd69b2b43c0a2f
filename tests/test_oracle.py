"""CPU tests: pin the oracle against the reference's own known answers / golden vectors and
against an independent dense conv3d; nothing here touches the CUDA library."""
import glob
import os
import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from msmdfusion_b200 import synthetic
from oracle import cpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_voxel_generator_known_answer():
    """tests/test_models/test_voxel_encoder/test_voxel_generator.py:6-21 of the reference."""
    np.random.seed(0)
    points = np.random.rand(1000, 4)
    voxels, coors, num = cpu.hard_voxelize(points.astype(np.float32), [0.5, 0.5, 0.5],
                                           [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    expected = np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1], [7, 81, 0], [6, 80, 1],
                         [7, 80, 0], [6, 80, 0]])
    assert np.all(coors == expected)
    assert voxels.shape == (8, 1000, 4)
    assert np.all(num == np.array([120, 121, 127, 134, 115, 127, 125, 131]))


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, 'voxelize_*.npz')))


def golden_points(name):
    from importlib import util
    spec = util.spec_from_file_location('make_golden', os.path.join(GOLDEN, 'make_golden.py'))
    mod = util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    key = os.path.basename(name)[len('voxelize_'):-len('.npz')]
    return mod.CASES[key][0]()


@pytest.mark.parametrize('path', golden_cases(), ids=os.path.basename)
def test_oracle_matches_reference_golden(path):
    """Fixtures were produced by the reference's own CPU hard_voxelize (make_golden.py)."""
    g = np.load(path)
    pts = golden_points(path)
    assert np.uint32(zlib.crc32(pts.tobytes())) == g['points_crc'], 'synthetic generator drifted'
    v, c, n = cpu.hard_voxelize(pts, g['voxel_size'], g['coors_range'], int(g['max_points']),
                                int(g['max_voxels']))
    assert c.shape[0] == int(g['voxel_num'])
    assert np.array_equal(c, g['coors'].astype(np.int32))
    assert np.array_equal(n, g['num_points'].astype(np.int32))
    assert np.uint32(zlib.crc32(np.ascontiguousarray(v).tobytes())) == g['voxels_crc']


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference tree not mounted')
def test_oracle_matches_reference_op_live():
    pts = synthetic.lidar_scene(seed=11, sweeps=1)[:8000]
    a = cpu.hard_voxelize(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 3000)
    b = cpu.hard_voxelize_ref(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 3000)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_fps_known_answer():
    """tests/test_models/test_common_modules/test_pointnet_ops.py:9-23 of the reference."""
    xyz = np.array([[[-0.2748, 1.0020, -1.1674], [0.1015, 1.3952, -1.2681], [-0.8070, 2.4137, -0.5845],
                     [-1.0001, 2.1982, -0.5859], [0.3841, 1.8983, -0.7431]],
                    [[-1.0696, 3.0758, -0.1899], [-0.2559, 3.5521, -0.1402], [0.8164, 4.0081, -0.1839],
                     [-1.1000, 3.0213, -0.8205], [-0.0518, 3.7251, -0.3950]]], np.float32)
    idx = np.stack([cpu.furthest_point_sample(xyz[b], 3) for b in range(2)])
    assert np.array_equal(idx, np.array([[0, 2, 4], [0, 2, 1]]))


BQ_NEW = np.array([[[-0.0740, 1.3147, -1.3625], [-2.2769, 2.7817, -0.2334], [-0.4003, 2.4666, -0.5116],
                    [-0.0740, 1.3147, -1.3625], [-0.0740, 1.3147, -1.3625]],
                   [[-2.0289, 2.4952, -0.1708], [-2.0668, 6.0278, -0.4875], [0.4066, 1.4211, -0.2947],
                    [-2.0289, 2.4952, -0.1708], [-2.0289, 2.4952, -0.1708]]], np.float32)
BQ_XYZ = np.array([[[-0.0740, 1.3147, -1.3625], [0.5555, 1.0399, -1.3634], [-0.4003, 2.4666, -0.5116],
                    [-0.5251, 2.4379, -0.8466], [-0.9691, 1.1418, -1.3733], [-0.2232, 0.9561, -1.3626],
                    [-2.2769, 2.7817, -0.2334], [-0.2822, 1.3192, -1.3645], [0.1533, 1.5024, -1.0432],
                    [0.4917, 1.1529, -1.3496]],
                   [[-2.0289, 2.4952, -0.1708], [-0.7188, 0.9956, -0.5096], [-2.0668, 6.0278, -0.4875],
                    [-1.9304, 3.3092, 0.6610], [0.0949, 1.4332, 0.3140], [-1.2879, 2.0008, -0.7791],
                    [-0.7252, 0.9611, -0.6371], [0.4066, 1.4211, -0.2947], [0.3220, 1.4447, 0.3548],
                    [-0.9744, 2.3856, -1.2000]]], np.float32)
BQ_EXPECT_0 = np.array([[[0, 0, 0, 0, 0], [6, 6, 6, 6, 6], [2, 2, 2, 2, 2], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]],
                        [[0, 0, 0, 0, 0], [2, 2, 2, 2, 2], [7, 7, 7, 7, 7], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]]])
BQ_EXPECT_1 = np.array([[[0, 5, 7, 0, 0], [6, 6, 6, 6, 6], [2, 3, 2, 2, 2], [0, 5, 7, 0, 0], [0, 5, 7, 0, 0]],
                        [[0, 0, 0, 0, 0], [2, 2, 2, 2, 2], [7, 7, 7, 7, 7], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]]])


def test_ball_query_known_answer():
    """test_pointnet_ops.py:26-74 of the reference (plain and dilated ball query)."""
    for (rmin, rmax, exp) in ((0, 0.2, BQ_EXPECT_0), (0.2, 0.4, BQ_EXPECT_1)):
        idx = np.stack([cpu.ball_query(rmin, rmax, 5, BQ_XYZ[b], BQ_NEW[b]) for b in range(2)])
        assert np.array_equal(idx, exp)


def random_sparse(rng, batch, shape, n, c):
    D, H, W = shape
    lin = rng.choice(batch * D * H * W, size=n, replace=False)
    rng.shuffle(lin)
    idx = np.stack([lin // (D * H * W), (lin // (H * W)) % D, (lin // W) % H, lin % W], 1).astype(np.int32)
    feat = rng.standard_normal((n, c)).astype(np.float32)
    return idx, feat


def dense_conv_oracle(idx, feat, shape, batch, weight, stride, padding, dilation):
    """Independent oracle: densify -> F.conv3d -> (B,Cout,oD,oH,oW)."""
    x = torch.from_numpy(cpu.dense(idx, feat, shape, batch)).double()
    w = torch.from_numpy(weight).double().permute(0, 4, 1, 2, 3).contiguous()  # KRSC -> (Co,Ci,kz,ky,kx)
    return F.conv3d(x, w, stride=stride, padding=padding, dilation=dilation).numpy()


@pytest.mark.parametrize('ksize,dilation', [(3, 1), ((3, 1, 1), 1), (3, 2), ((1, 3, 3), 1)])
def test_subm_conv_vs_dense(ksize, dilation):
    rng = np.random.default_rng(0)
    shape, batch = [7, 12, 10], 2
    idx, feat = random_sparse(rng, batch, shape, 300, 6)
    ks = cpu._triple(ksize)
    w = rng.standard_normal((9, *ks, 6)).astype(np.float32)
    pair = cpu.subm_rulebook(idx, shape, ksize, dilation)
    out = cpu.spconv_fwd(feat, w, pair)
    dl = cpu._triple(dilation)
    pad = [(k // 2) * d for k, d in zip(ks, dl)]
    ref = dense_conv_oracle(idx, feat, shape, batch, w, 1, pad, dl)
    got = ref[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    assert np.abs(out - got).max() < 1e-4


@pytest.mark.parametrize('ksize,stride,padding', [(3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0),
                                                  (2, 2, 0), (3, 1, 0), (3, (1, 2, 3), (1, 0, 2))])
def test_strided_conv_vs_dense(ksize, stride, padding):
    rng = np.random.default_rng(1)
    shape, batch = [9, 14, 11], 2
    idx, feat = random_sparse(rng, batch, shape, 400, 5)
    ks = cpu._triple(ksize)
    w = rng.standard_normal((7, *ks, 5)).astype(np.float32)
    out_idx, pair, out_shape = cpu.conv_rulebook(idx, shape, ksize, stride, padding, 1)
    out = cpu.spconv_fwd(feat, w, pair)
    ref = dense_conv_oracle(idx, feat, shape, batch, w, cpu._triple(stride), cpu._triple(padding), 1)
    assert list(ref.shape[2:]) == out_shape
    # ascending linear order, unique
    lin = ((out_idx[:, 0].astype(np.int64) * out_shape[0] + out_idx[:, 1]) * out_shape[1]
           + out_idx[:, 2]) * out_shape[2] + out_idx[:, 3]
    assert np.all(np.diff(lin) > 0)
    got = ref[out_idx[:, 0], :, out_idx[:, 1], out_idx[:, 2], out_idx[:, 3]]
    assert np.abs(out - got).max() < 1e-4
    # every output position the dense conv can reach from an active input is in the set:
    occ = torch.from_numpy(cpu.dense(idx, np.ones((idx.shape[0], 1), np.float32), shape, batch))
    reach = F.conv3d(occ, torch.ones(1, 1, *ks), stride=cpu._triple(stride),
                     padding=cpu._triple(padding)).numpy()[:, 0] > 0
    assert reach.sum() == out_idx.shape[0]
    assert reach[out_idx[:, 0], out_idx[:, 1], out_idx[:, 2], out_idx[:, 3]].all()


def test_sparse_add_and_dense():
    rng = np.random.default_rng(2)
    shape = [5, 8, 9]
    ia, fa = random_sparse(rng, 2, shape, 120, 4)
    ib, fb = random_sparse(rng, 2, shape, 150, 4)
    oi, of = cpu.sparse_add(ia, fa, ib, fb, shape)
    assert np.allclose(cpu.dense(oi, of, shape, 2), cpu.dense(ia, fa, shape, 2) + cpu.dense(ib, fb, shape, 2))
    lin = ((oi[:, 0].astype(np.int64) * shape[0] + oi[:, 1]) * shape[1] + oi[:, 2]) * shape[2] + oi[:, 3]
    assert np.all(np.diff(lin) > 0)


def test_float_key_collisions_reproduced():
    """the float32 key z*1e6+y*1e3+x collapses neighbouring x for z >= 17."""
    c = np.array([[17, 1000, 700], [17, 1000, 701], [3, 10, 7], [3, 10, 8]], np.int32)
    k = cpu.float_key(c)
    assert k[0] == k[1] or abs(float(k[1]) - float(k[0])) == 2.0  # rounded to even spacing
    assert k[2] != k[3]
    ref = (torch.from_numpy(c[:, 0]) * 1e6 + torch.from_numpy(c[:, 1]) * 1e3 + torch.from_numpy(c[:, 2])).numpy()
    assert ref.dtype == np.float32 and np.array_equal(ref, k)


def test_modality_split_consistency():
    rng = np.random.default_rng(3)
    shape = [41, 200, 200]
    i3, _ = random_sparse(rng, 2, shape, 3000, 1)
    i2, _ = random_sparse(rng, 2, shape, 2500, 1)
    i2[:800] = i3[rng.choice(3000, 800, replace=False)]  # force overlaps
    i3 = i3[np.argsort(i3[:, 0], kind='stable')]
    i2 = i2[np.argsort(i2[:, 0], kind='stable')]
    c3, c2, s3, s2 = cpu.voxel_modality_split(i3, i2, 2)
    assert c3.shape == (3000, 5) and c2.shape == (2500, 5)
    assert s3.shape == s2.shape and s3.shape[0] >= 800
    # matched pairs carry equal float keys (sample 0 has offset 0)
    n0 = int((i3[:, 0] == 0).sum())
    m0 = int((i2[:, 0] == 0).sum())
    k3, k2 = cpu.float_key(i3[:, 1:]), cpu.float_key(i2[:, 1:])
    first = s3 < n0
    assert np.array_equal(k3[s3[first & (s2 < m0)]], k2[s2[first & (s2 < m0)]])
    assert c3[:, 1].sum() == s3.shape[0] and c2[:, 1].sum() == s2.shape[0]


def test_fusion_oracle_runs_end_to_end_small():
    """The CPU restatement of lift -> 4-scale voxels -> modality split -> GMA encoder executes on a
    small seeded scene and is self-consistent (shapes, channel widths, sorted strided outputs)."""
    import msmdfusion_b200 as m
    from msmdfusion_b200 import synthetic
    from oracle import model as omodel
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = m.Config.fromfile(os.path.join(root, 'configs', 'msmd_lc_hotpath.py')).hotpath
    torch.manual_seed(0)
    det = m.MSMDFusionDetector(**{k: cfg[k] for k in (
        'pts_voxel_layer', 'pts_voxel_encoder', 'pts_middle_encoder', 'multimodal_middle_encoder',
        'spatial_shapes', 'downscale_factors', 'fps_num_list', 'radius_list', 'max_cluster_samples_list',
        'dist_thresh_list')}).eval()
    sd = det.state_dict()
    scene = synthetic.lidar_scene(5, 1)[:3000]
    metas = [synthetic.camera_scene(5, scene, virtual_per_camera=500, real_per_camera=50)]
    ev, en, ec = omodel.voxelize_batch([scene], synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    emean = cpu.hard_simple_vfe(ev, en, 5)
    _, e_feats, _ = omodel.sparse_encoder(sd, dict(cfg.pts_middle_encoder), emean, ec, 1, prefix='pts_middle_encoder.')
    rng = np.random.default_rng(0)
    H, W = synthetic.INPUT_SHAPE
    comp = [np.abs(rng.standard_normal((6, 49, H // s, W // s))).astype(np.float32) for s in (4, 8, 16)]
    img_list = [comp[0]] + comp
    score_w = np.full((66,), 0.01, np.float32)
    v3l, v2l, s3l, s2l = [], [], [], []
    for i in range(4):
        v2 = omodel.fetch_2d_voxels(img_list[i], metas, score_w, 0.0, cfg.spatial_shapes[i],
                                    cfg.downscale_factors[i], synthetic.VOXEL_SIZE,
                                    synthetic.POINT_CLOUD_RANGE, 10, 160000)
        assert v2.features.shape[1] == 64
        c3, c2, s3, s2 = cpu.voxel_modality_split(e_feats[i].indices, v2.indices, 1)
        v3l.append(omodel.SpTensor(e_feats[i].features, c3, e_feats[i].spatial_shape, 1))
        v2l.append(omodel.SpTensor(v2.features, c2, v2.spatial_shape, 1))
        s3l.append(s3)
        s2l.append(s2)
    dummies = [rng.random((1, c), dtype=np.float32) for c in (16, 32, 64, 128)]
    outs = omodel.multimodal_encoder(sd, dict(cfg.multimodal_middle_encoder), v3l, v2l, s3l, s2l,
                                     cfg.fps_num_list, cfg.radius_list, cfg.max_cluster_samples_list,
                                     cfg.dist_thresh_list, dummies, prefix='multimodal_middle_encoder.')
    assert [o.features.shape[1] for o in outs] == [96, 128, 192, 192]
    assert [o.spatial_shape for o in outs] == [[21, 720, 720], [11, 360, 360], [5, 180, 180], [2, 180, 180]]
    for o in outs:
        s = o.spatial_shape
        lin = ((o.indices[:, 0].astype(np.int64) * s[0] + o.indices[:, 1]) * s[1] + o.indices[:, 2]) * s[2] + o.indices[:, 3]
        assert np.all(np.diff(lin) > 0) and np.isfinite(o.features).all()
    canvas = omodel.depth_canvas(metas, H, W)
    assert canvas.shape == (6, 1, H, W) and (canvas > 0).sum() > 0


# --------------------------------------------------------------------------------------
# sparse conv / rulebook restatement pinned against the reference's VENDORED spconv-1.x
# --------------------------------------------------------------------------------------
def spconv1x_goldens():
    return sorted(glob.glob(os.path.join(GOLDEN, 'spconv1x_*.npz')))


def oracle_conv_on_golden(g):
    idx = g['indices'].astype(np.int32)
    shape = [int(s) for s in g['spatial_shape']]
    ks, st, pd = [int(x) for x in g['ksize']], [int(x) for x in g['stride']], [int(x) for x in g['padding']]
    if int(g['subm']):
        pair, oi, oshape = cpu.subm_rulebook(idx, shape, ks, 1), idx, shape
    else:
        oi, pair, oshape = cpu.conv_rulebook(idx, shape, ks, st, pd, 1)
    return oi, cpu.spconv_fwd(g['features'], g['weight_krsc'], pair), list(oshape)


@pytest.mark.parametrize('path', spconv1x_goldens(), ids=os.path.basename)
def test_oracle_conv_matches_reference_spconv1x_golden(path):
    """Fixtures produced by the reference's own spconv-1.x CPU ops (tests/golden/make_golden_spconv.py):
    output index set (ascending linear order) bit-exact, features to fp32 rounding."""
    g = np.load(path)
    oi, of, oshape = oracle_conv_on_golden(g)
    assert oshape == [int(s) for s in g['out_shape']]
    assert np.array_equal(oi, g['out_indices'].astype(np.int32))
    assert np.abs(of - g['out_features']).max() < 1e-5


@pytest.mark.skipif(not os.path.isdir('/root/reference/mmdet3d/ops/spconv/src'), reason='reference tree not mounted')
def test_oracle_conv_matches_reference_spconv1x_live():
    from oracle import ref_spconv
    if ref_spconv.module() is None:
        pytest.skip('reference spconv-1.x could not be built here')
    rng = np.random.default_rng(21)
    for (shape, batch, n, cin, cout, ks, st, pd, subm) in (
            ([7, 18, 18], 2, 700, 8, 8, 3, 1, 1, True), ([7, 18, 18], 2, 700, 8, 16, 3, 2, 1, False),
            ([9, 12, 12], 1, 400, 4, 4, [3, 1, 1], [2, 1, 1], 0, False)):
        idx, feat = random_sparse(rng, batch, shape, n, cin)
        k3 = ks if isinstance(ks, list) else [ks] * 3
        w = rng.standard_normal((cout, *k3, cin)).astype(np.float32) * 0.2
        oi, of, oshape = ref_spconv.conv(idx, feat, w, shape, batch, ks, st, pd, 1, subm)
        if subm:
            ei, pair = idx, cpu.subm_rulebook(idx, shape, ks, 1)
        else:
            ei, pair, es = cpu.conv_rulebook(idx, shape, ks, st, pd, 1)
            assert list(es) == list(oshape)
        assert np.array_equal(oi, ei)
        assert np.abs(of - cpu.spconv_fwd(feat, w, pair)).max() < 1e-5


# --------------------------------------------------------------------------------------
# voxel_modality_split: the reference's own numba merge, executed from where it lies
# --------------------------------------------------------------------------------------
REF_DETECTOR = '/root/reference/mmdet3d/models/detectors/MSMDFusion.py'


def reference_type_assign():
    """Compile the reference's `type_assign` (MSMDFusion.py:26-45) from its own source text at test
    time -- the module itself cannot be imported here (mmcv / mmdet / spconv are absent)."""
    import numba  # noqa: F401
    lines = open(REF_DETECTOR).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('def type_assign'))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith('class ') or lines[i].startswith('def '))
    ns = {}
    exec('from numba import jit\nimport numpy as np\n@jit(nopython=True)\n' + '\n'.join(lines[start:end]), ns)
    return ns['type_assign']


@pytest.mark.skipif(not os.path.exists(REF_DETECTOR), reason='reference tree not mounted')
def test_modality_split_matches_reference_type_assign_live():
    """oracle.cpu.type_assign / voxel_modality_split against the reference's numba merge and the
    reference's key / sort expressions (MSMDFusion.py:271-275) evaluated with torch on the CPU."""
    ref_assign = reference_type_assign()
    rng = np.random.default_rng(31)
    shape = [41, 300, 300]
    i3, _ = random_sparse(rng, 1, shape, 4000, 1)
    i2, _ = random_sparse(rng, 1, shape, 3000, 1)
    i2[:900] = i3[rng.choice(4000, 900, replace=False)]
    i2[900:930, 3] += 1      # x-neighbours: collide under the float32 key for z >= 17
    c3, c2 = torch.from_numpy(i3[:, 1:]), torch.from_numpy(i2[:, 1:])
    k3 = c3[:, 0] * 1e6 + c3[:, 1] * 1e3 + c3[:, 2]          # :271 (int32 tensor * python float)
    k2 = c2[:, 0] * 1e6 + c2[:, 1] * 1e3 + c2[:, 2]
    assert np.array_equal(k3.numpy(), cpu.float_key(i3[:, 1:])) and k3.dtype == torch.float32
    v3, ind3 = torch.sort(k3, dim=-1, stable=True)           # :274 (stable: one legal outcome)
    v2, ind2 = torch.sort(k2, dim=-1, stable=True)
    t3, t2 = ref_assign(v3.numpy(), v2.numpy(), np.zeros_like(v3.numpy()), np.zeros_like(v2.numpy()))
    o3, o2 = cpu.type_assign(v3.numpy(), v2.numpy())
    assert np.array_equal(t3, o3) and np.array_equal(t2, o2) and t3.sum() >= 900
    # full restatement: mix flags and the matched row ids in sorted-key order
    e3, e2, s3, s2 = cpu.voxel_modality_split(i3, i2, 1)
    mix3 = np.zeros(i3.shape[0], np.int32)
    mix3[ind3.numpy()] = t3.astype(np.int32)
    mix2 = np.zeros(i2.shape[0], np.int32)
    mix2[ind2.numpy()] = t2.astype(np.int32)
    assert np.array_equal(e3[:, 1], mix3) and np.array_equal(e2[:, 1], mix2)
    assert np.array_equal(s3, ind3.numpy()[np.nonzero(t3)[0]]) and np.array_equal(s2, ind2.numpy()[np.nonzero(t2)[0]])
