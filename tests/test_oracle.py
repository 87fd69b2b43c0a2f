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
def split_case(seed, shape, n3, n2, shared):
    rng = np.random.default_rng(seed)
    i3, _ = random_sparse(rng, 1, shape, n3, 1)
    i2, _ = random_sparse(rng, 1, shape, n2, 1)
    i2[:shared] = i3[rng.choice(i3.shape[0], shared, replace=False)]
    i2[shared:shared + 30, 3] += 1      # x-neighbours: collide under the float32 key for z >= 17
    return i3, i2[rng.permutation(i2.shape[0])]


@pytest.mark.skipif(not __import__('oracle.ref_split', fromlist=['x']).available(), reason='reference tree not mounted')
def test_modality_split_matches_reference_type_assign_live():
    """oracle.cpu.type_assign / voxel_modality_split against the reference's numba merge and the
    reference's key / sort expressions (MSMDFusion.py:271-275) evaluated with torch on the CPU."""
    from oracle import ref_split
    i3, i2 = split_case(31, [41, 300, 300], 4000, 3000, 900)
    k3, k2 = np.sort(cpu.float_key(i3[:, 1:])), np.sort(cpu.float_key(i2[:, 1:]))
    o3, o2 = cpu.type_assign(k3, k2)
    r3, r2 = ref_split.type_assign()(k3, k2, np.zeros_like(k3), np.zeros_like(k2))
    assert np.array_equal(r3, o3) and np.array_equal(r2, o2) and r3.sum() >= 900
    # the float key itself: int32 tensor * python float, as MSMDFusion.py:271 writes it
    c3 = torch.from_numpy(i3[:, 1:])
    t3 = c3[:, 0] * 1e6 + c3[:, 1] * 1e3 + c3[:, 2]
    assert t3.dtype == torch.float32 and np.array_equal(t3.numpy(), cpu.float_key(i3[:, 1:]))


@pytest.mark.skipif(not __import__('oracle.ref_split', fromlist=['x']).available(), reason='reference tree not mounted')
@pytest.mark.parametrize('batch', [1, 2])
def test_modality_split_matches_reference_method_live(batch):
    """oracle.cpu.voxel_modality_split against MSMDFusionDetector.voxel_modality_split itself
    (MSMDFusion.py:251-325, whole method run in place): the (b, mix, z, y, x) index tensors and the
    batch-offset synchronisation indices, including duplicate rows and float32-key collisions."""
    from oracle import ref_split
    parts3, parts2 = [], []
    for b in range(batch):
        i3, i2 = split_case(40 + b, [41, 300, 300], 4000 - 500 * b, 3000 + 300 * b, 900)
        i2[50:60] = i2[40:50]
        i3[:, 0] = b
        i2[:, 0] = b
        parts3.append(i3)
        parts2.append(i2)
    i3, i2 = np.concatenate(parts3), np.concatenate(parts2)
    r3, r2, rs3, rs2 = ref_split.voxel_modality_split(i3, i2, batch)
    e3, e2, es3, es2 = cpu.voxel_modality_split(i3, i2, batch)
    assert r3.dtype == np.int32 and r3.shape == (i3.shape[0], 5)
    assert np.array_equal(r3, e3) and np.array_equal(r2, e2)
    assert rs3.dtype == np.int64 and np.array_equal(rs3, es3) and np.array_equal(rs2, es2)
    assert rs3.shape[0] >= 900 * batch


@pytest.mark.parametrize('name', ['split_dense_overlap', 'split_lidar_grid', 'split_disjoint'])
def test_modality_split_matches_reference_golden(name):
    """Committed outputs of the reference's own merge (tests/golden/make_golden_split.py)."""
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    e3, e2, s3, s2 = cpu.voxel_modality_split(g['indices3'], g['indices2'], 1)
    assert np.array_equal(e3[:, 1], g['mix3']) and np.array_equal(e2[:, 1], g['mix2'])
    assert np.array_equal(s3, g['syn3']) and np.array_equal(s2, g['syn2'])


# --------------------------------------------------------------------------------------
# fps_NN_fast: the reference's own method body, executed from where it lies
# --------------------------------------------------------------------------------------
ASSIGN_CASES = ['assign_scale0', 'assign_scale1', 'assign_scale2', 'assign_scale3', 'assign_direct']


@pytest.mark.skipif(not __import__('oracle.ref_assign', fromlist=['x']).available(), reason='reference tree not mounted')
@pytest.mark.parametrize('Q,K,fps_num,radius,nsample,thresh', [(300, 500, 512, 2., 8, 3.), (3000, 2000, 512, 3., 16, 4.),
                                                              (5000, 3000, 1024, 2., 8, 1.5), (2500, 40, 256, 6., 32, 8.)])
def test_fps_nn_fast_matches_reference_method_live(Q, K, fps_num, radius, nsample, thresh):
    """oracle.cpu.fps_nn_fast against `SparseMultiModalEncoderPaint.fps_NN_fast` itself
    (sparse_multimodal_encoder_painting.py:276-323): distance matrix, first-minimum tie-break, threshold
    and the duplicate-index scatter are the reference's code run by torch (oracle/ref_assign.py)."""
    from oracle import ref_assign
    rng = np.random.default_rng(Q + K)
    both, _ = random_sparse(rng, 1, [11, 60, 60], Q + K, 1)
    q, k = both[:Q], both[Q:]
    got = cpu.fps_nn_fast(q, k, fps_num, radius, nsample, thresh)
    exp = ref_assign.fps_nn_fast(q, k, fps_num, radius, nsample, thresh)
    assert exp.dtype == np.int64 and np.array_equal(got, exp) and (exp >= 0).any()


@pytest.mark.parametrize('name', ASSIGN_CASES)
def test_fps_nn_fast_matches_reference_golden(name):
    """Committed outputs of the reference's own method (tests/golden/make_golden_assign.py) at the four
    scales' parameters of configs/MSMDFusion_nusc_voxel_LC.py:146-149."""
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    fps_num, radius, nsample, thresh = g['params']
    got = cpu.fps_nn_fast(g['query'], g['key'], int(fps_num), float(radius), int(nsample), float(thresh))
    assert np.array_equal(got, g['assign'])


# --------------------------------------------------------------------------------------
# 2D -> 3D lift: the reference's own get_foreground2D / depth canvas, executed from where they lie
# --------------------------------------------------------------------------------------
def load_lift_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden_lift', os.path.join(GOLDEN, 'make_golden_lift.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    metas, feat, score_w, score_b = mod.lift_inputs()
    g = np.load(os.path.join(GOLDEN, 'lift_reference.npz'))
    assert mod.inputs_crc(metas, feat, score_w, score_b) == int(g['inputs_crc'][0]), 'synthetic generator drifted'
    return mod, g, metas, feat, score_w, score_b


@pytest.mark.skipif(not __import__('oracle.ref_lift', fromlist=['x']).available(), reason='reference tree not mounted')
def test_lift_matches_reference_methods_live():
    """oracle.model.get_foreground2d / depth_canvas against MSMDFusionDetector.get_foreground2D and
    .depth_aware_channel_compression (MSMDFusion.py:169-238, 335-369) run by torch (oracle/ref_lift.py)."""
    from oracle import model as omodel, ref_lift
    mod, _, metas, feat, score_w, score_b = load_lift_golden()
    # gate == 1 (w = 0, b = 1): pixel truncation, the (h, w) gather, camera / sample order, the empty
    # camera and the concatenation are compared bit for bit
    zero_w = np.zeros_like(score_w)
    for a, b in zip(ref_lift.get_foreground2d(feat, metas, zero_w, np.float32(1)),
                    omodel.get_foreground2d(feat, metas, zero_w, np.float32(1))):
        assert a.shape == b.shape and np.array_equal(a, b)
    # learned gate: the 66-term dot product is summed in a different order (torch addmm vs numpy)
    for a, b in zip(ref_lift.get_foreground2d(feat, metas, score_w, score_b),
                    omodel.get_foreground2d(feat, metas, score_w, score_b)):
        assert np.array_equal(a[:, :15], b[:, :15])
        assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(a).max())
        assert np.array_equal((a[:, 15:] != 0).any(1), (b[:, 15:] != 0).any(1))
    H, W = synthetic.INPUT_SHAPE
    shapes = [(H // s, W // s) for s in (4, 8, 16)]
    canvas = torch.from_numpy(omodel.depth_canvas(metas, H, W))
    for (h, w), ref in zip(shapes, ref_lift.depth_maps(shapes, metas)):
        assert np.array_equal(ref, F.interpolate(canvas, (h, w), mode='bilinear').numpy())


def test_lift_matches_reference_golden():
    from oracle import model as omodel
    mod, g, metas, feat, score_w, score_b = load_lift_golden()
    for b, a in enumerate(omodel.get_foreground2d(feat, metas, score_w, score_b)):
        assert a.shape[0] == int(g['count%d' % b][0])
        assert zlib.crc32(np.ascontiguousarray(a[:, :15]).tobytes()) == int(g['points_crc%d' % b][0])
        ref = g['rows%d' % b]
        assert np.abs(a[::mod.ROW_STEP] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    H, W = synthetic.INPUT_SHAPE
    canvas = omodel.depth_canvas(metas, H, W).reshape(-1)
    assert canvas.shape[0] == int(g['canvas_size'][0])
    nz = np.nonzero(canvas)[0]
    assert np.array_equal(nz, g['canvas_index']) and np.array_equal(canvas[nz], g['canvas_value'])


# --------------------------------------------------------------------------------------
# GMA encoder glue: the reference's own grouped_sparse_conv / forward, executed from where they lie
# --------------------------------------------------------------------------------------
@pytest.mark.skipif(not __import__('oracle.ref_encoder', fromlist=['x']).available(), reason='reference tree not mounted')
def test_multimodal_encoder_matches_reference_glue_live():
    """oracle.model.multimodal_encoder against SparseMultiModalEncoderPaint.forward /
    .grouped_sparse_conv / .pad_missing_batch_id / .fps_NN_fast themselves
    (sparse_multimodal_encoder_painting.py:208-225, 276-459), run by torch around the oracle's conv
    restatement (oracle/ref_encoder.py).  Batch of 2 with an empty camera: the per-sample assignment
    loop, its `base` offset, the dummy-embedding row and the batch padding are all on the path."""
    import msmdfusion_b200 as m
    from oracle import model as omodel, ref_encoder
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = m.Config.fromfile(os.path.join(root, 'configs', 'msmd_lc_hotpath.py')).hotpath
    torch.manual_seed(0)
    det = m.MSMDFusionDetector(**{k: cfg[k] for k in (
        'pts_voxel_layer', 'pts_voxel_encoder', 'pts_middle_encoder', 'multimodal_middle_encoder',
        'spatial_shapes', 'downscale_factors', 'fps_num_list', 'radius_list', 'max_cluster_samples_list',
        'dist_thresh_list')}).eval()
    gen = torch.Generator().manual_seed(9)
    for mod in det.modules():   # non-trivial eval-mode statistics
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=gen) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=gen) + 0.5)
    sd = det.state_dict()
    B = 2
    scenes = [synthetic.lidar_scene(5 + b, 1)[:4000] for b in range(B)]
    metas = [synthetic.camera_scene(5 + b, scenes[b], virtual_per_camera=2500 if b == 0 else 150, real_per_camera=50,
                                    empty_cameras=(() if b == 0 else (1,))) for b in range(B)]
    ev, en, ec = omodel.voxelize_batch(scenes, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    _, e_feats, _ = omodel.sparse_encoder(sd, dict(cfg.pts_middle_encoder), cpu.hard_simple_vfe(ev, en, 5), ec, B,
                                          prefix='pts_middle_encoder.')
    rng = np.random.default_rng(0)
    H, W = synthetic.INPUT_SHAPE
    comp = [np.abs(rng.standard_normal((6 * B, 49, H // s, W // s))).astype(np.float32) for s in (4, 8, 16)]
    img_list = [comp[0]] + comp
    v3l, v2l, s3l, s2l = [], [], [], []
    for i in range(4):
        v2 = omodel.fetch_2d_voxels(img_list[i], metas, np.full((66,), 0.01, np.float32), 0.0, cfg.spatial_shapes[i],
                                    cfg.downscale_factors[i], synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE,
                                    10, 160000)
        c3, c2, s3, s2 = cpu.voxel_modality_split(e_feats[i].indices, v2.indices, B)
        v3l.append(omodel.SpTensor(e_feats[i].features, c3, e_feats[i].spatial_shape, B))
        v2l.append(omodel.SpTensor(v2.features, c2, v2.spatial_shape, B))
        s3l.append(s3)
        s2l.append(s2)
    assert sum(s.shape[0] for s in s3l) > 100
    fps = [256] * 4     # smaller than most only-2D sets: the FPS + ball-query branch runs at every scale
    args = (v3l, v2l, s3l, s2l, fps, cfg.radius_list, cfg.max_cluster_samples_list, cfg.dist_thresh_list)
    torch.manual_seed(77)
    dummies = [torch.rand(1, c).numpy() for c in cfg.multimodal_middle_encoder['in_channels_3D']]
    mine = omodel.multimodal_encoder(sd, dict(cfg.multimodal_middle_encoder), *args, dummies,
                                     prefix='multimodal_middle_encoder.')
    ref = ref_encoder.forward(sd, dict(cfg.multimodal_middle_encoder), *args, 77, prefix='multimodal_middle_encoder.')
    for a, b in zip(ref, mine):
        assert a.spatial_shape == b.spatial_shape and np.array_equal(a.indices, b.indices)
        # the two gate MLPs run through torch addmm there and numpy matmul here
        assert np.abs(a.features - b.features).max() <= 2e-6 * max(1.0, np.abs(a.features).max())


# --------------------------------------------------------------------------------------
# LiDAR SparseEncoder: the reference's own class, executed from where it lies
# --------------------------------------------------------------------------------------
ENCODER_CFGS = {
    # configs/MSMDFusion_nusc_voxel_LC.py / transfusion_nusc_voxel_L.py pts_middle_encoder
    'basicblock': dict(type='SparseEncoder', in_channels=5, sparse_shape=[41, 1440, 1440], output_channels=128,
                       order=('conv', 'norm', 'act'),
                       encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
                       encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock'),
    # the class defaults (SECOND-style conv_module stack)
    'conv_module': dict(type='SparseEncoder', in_channels=4, sparse_shape=[41, 400, 352]),
}


def assert_same_layer_table(net, table):
    """`table`: [(qualified name, spec)] recorded by the reference's own constructor (oracle/ref_stubs.py);
    `net`: this package's module -- same names, same conv type / channels / kernel / stride / padding."""
    from msmdfusion_b200 import spconv as sp
    from msmdfusion_b200.sparse_block import SparseBasicBlock
    ours = dict(net.named_modules())
    triple = lambda v: [v] * 3 if isinstance(v, int) else list(v)  # noqa: E731
    for name, spec in table:
        mod = ours[name]
        if spec['kind'] == 'basicblock':
            assert isinstance(mod, SparseBasicBlock)
            assert mod.conv1.in_channels == spec['cin'] and mod.conv2.out_channels == spec['cout']
            assert isinstance(mod.conv1, sp.SubMConv3d) and isinstance(mod.conv2, sp.SubMConv3d)
            continue
        conv, bn, act = list(mod.children())
        assert type(conv).__name__ == spec['conv_type'] and isinstance(bn, torch.nn.BatchNorm1d)
        assert isinstance(act, torch.nn.ReLU) and bn.eps == spec['eps']
        assert (conv.in_channels, conv.out_channels) == (spec['cin'], spec['cout'])
        assert list(conv.kernel_size) == triple(spec['ksize']) and list(conv.stride) == triple(spec['stride'])
        assert list(conv.padding) == triple(spec['padding']) and conv.indice_key == spec['indice_key']
    assert {n for n, mod in ours.items() if isinstance(mod, SparseBasicBlock) or
            (isinstance(mod, sp.SparseSequential) and any(isinstance(c, sp.SparseConvolution) for c in mod.children()))
            } == {n for n, _ in table}


@pytest.mark.skipif(not __import__('oracle.ref_lidar_encoder', fromlist=['x']).available(), reason='reference tree not mounted')
@pytest.mark.parametrize('kind', list(ENCODER_CFGS))
def test_sparse_encoder_layer_table_matches_reference_class_live(kind):
    """The layer table the reference's own SparseEncoder constructor lays out (sparse_encoder.py:31-209,
    run in place by oracle/ref_lidar_encoder.py) against this package's module tree: same qualified
    names, same conv type / channels / kernel / stride / padding at every position."""
    import msmdfusion_b200 as m
    from oracle import ref_lidar_encoder
    cfg = ENCODER_CFGS[kind]
    ours = m.SparseEncoder(**{k: v for k, v in cfg.items() if k != 'type'})
    table = ref_lidar_encoder.layer_table(cfg)
    assert len(table) == (13 if kind == 'basicblock' else 12)
    assert_same_layer_table(ours, table)


@pytest.mark.skipif(not __import__('oracle.ref_lidar_encoder', fromlist=['x']).available(), reason='reference tree not mounted')
@pytest.mark.parametrize('kind', list(ENCODER_CFGS))
def test_sparse_encoder_forward_matches_reference_class_live(kind):
    """oracle.model.sparse_encoder against the reference's SparseEncoder.forward (sparse_encoder.py:96-133)
    run in place around the oracle's conv restatement: same encode_features in the same order, same dense
    (N, C*D, H, W) view."""
    import msmdfusion_b200 as m
    from oracle import model as omodel, ref_lidar_encoder
    cfg = dict(ENCODER_CFGS[kind])
    cfg['sparse_shape'] = [41, 96, 96]
    torch.manual_seed(3)
    net = m.SparseEncoder(**{k: v for k, v in cfg.items() if k != 'type'}).eval()
    sd = net.state_dict()
    rng = np.random.default_rng(8)
    idx, feat = random_sparse(rng, 2, cfg['sparse_shape'], 3000, cfg['in_channels'])
    e_spatial, e_feats, _ = omodel.sparse_encoder(sd, cfg, feat, idx, 2)
    r_spatial, r_feats = ref_lidar_encoder.forward(sd, cfg, feat, idx, 2)
    assert r_spatial.shape == e_spatial.shape and np.array_equal(r_spatial, e_spatial)
    assert len(r_feats) == len(e_feats) == 5
    for a, b in zip(r_feats, e_feats):
        assert a.spatial_shape == b.spatial_shape and np.array_equal(a.indices, b.indices)
        assert np.array_equal(a.features, b.features)


@pytest.mark.skipif(not __import__('oracle.ref_encoder', fromlist=['x']).available(), reason='reference tree not mounted')
def test_multimodal_encoder_layout_matches_reference_class_live():
    """Module layout of the reference's own SparseMultiModalEncoderPaint constructor
    (sparse_multimodal_encoder_painting.py:99-205, run in place) against this package's class: the 20
    sparse-conv blocks by qualified name and the gate MLPs' state-dict keys and shapes."""
    import msmdfusion_b200 as m
    from oracle import ref_encoder
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = dict(m.Config.fromfile(os.path.join(root, 'configs', 'msmd_lc_hotpath.py')).hotpath.multimodal_middle_encoder)
    ref = ref_encoder.build(cfg)
    ours = m.SparseMultiModalEncoderPaint(**{k: v for k, v in cfg.items() if k != 'type'})
    table = ref_encoder.ref_stubs.layer_table(ref)
    assert len(table) == 20
    assert_same_layer_table(ours, table)
    gates = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert len(gates) == 16 and all(k.startswith(('gate_control.', 'cross_gate_control.')) for k in gates)
    mine = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert all(mine[k] == shape for k, shape in gates.items())
    assert {k for k in mine if 'gate_control' in k} == set(gates)


# --------------------------------------------------------------------------------------
# Detector level: the reference's own extract_pts_feat, executed from where it lies
# --------------------------------------------------------------------------------------
def load_detector_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden_detector', os.path.join(GOLDEN, 'make_golden_detector.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, np.load(os.path.join(GOLDEN, 'detector_reference.npz'))


@pytest.mark.skipif(not __import__('oracle.ref_detector', fromlist=['x']).available(), reason='reference tree not mounted')
def test_extract_voxel_space_matches_reference_detector_live():
    """oracle.model.extract_voxel_space against MSMDFusionDetector.extract_pts_feat itself
    (MSMDFusion.py:421-445 and every method / class it reaches, run in place by oracle/ref_detector.py
    with the reference's own C++ voxelizer): the four stage outputs' indices bit for bit, their features
    and the (B, 640, 180, 180) tensor handed to bev_fusion within fp32 summation-order noise."""
    import _fixtures
    from oracle import model as omodel, ref_detector
    det, cfg = _fixtures.build_msmd_detector(1)
    sd = det.state_dict()
    scenes, metas, fpn = _fixtures.lc_scene(1, points=20000, virtual=(1500, 300))
    r_bev, r_outs = ref_detector.extract_voxel_space(sd, cfg, scenes, fpn, metas, 77)
    torch.manual_seed(77)
    dummies = [torch.rand(1, c).numpy() for c in cfg.multimodal_middle_encoder['in_channels_3D']]
    e_bev, e_outs, _ = omodel.extract_voxel_space(sd, cfg, scenes, fpn, metas, dummies)
    assert r_bev.shape == e_bev.shape == (1, 640, 180, 180)
    for a, b in zip(r_outs, e_outs):
        assert a.spatial_shape == b.spatial_shape and np.array_equal(a.indices, b.indices)
        assert np.abs(a.features - b.features).max() <= 1e-5 * max(1.0, np.abs(a.features).max())
    assert np.abs(r_bev - e_bev).max() <= 1e-5 * max(1.0, np.abs(r_bev).max())
    assert np.count_nonzero(r_bev) == np.count_nonzero(e_bev) > 100000


def check_against_detector_golden(mod, g, bev, outs, tol):
    """bev: (B,640,180,180) array; outs: [(indices (N,4) int32, features (N,C))] per stage."""
    assert list(bev.shape) == g['bev_shape'].tolist()
    for i, (idx, feat) in enumerate(outs):
        assert idx.shape[0] == int(g['count%d' % i][0])
        assert zlib.crc32(np.ascontiguousarray(idx, np.int32).tobytes()) == int(g['indices_crc%d' % i][0])
        ref = g['rows%d' % i]
        assert np.abs(feat[::mod.ROW_STEP] - ref).max() <= tol * max(1.0, np.abs(ref).max())
    flat = bev.reshape(-1)
    ref = g['bev_values']
    assert np.abs(flat[mod.bev_positions(flat.shape[0])] - ref).max() <= tol * max(1.0, float(g['bev_absmax'][0]))
    assert np.count_nonzero(flat) == int(g['bev_nonzero'][0])


def test_extract_voxel_space_matches_reference_golden():
    """The oracle chain against the committed output of the reference's own extract_pts_feat on the
    batch-2 scene (tests/golden/make_golden_detector.py) -- the same fixture the CUDA path is checked
    against on the GPU box."""
    import _fixtures
    from oracle import model as omodel
    mod, g = load_detector_golden()
    det, cfg = _fixtures.build_msmd_detector(mod.DET_SEED)
    sd = det.state_dict()
    assert _fixtures.state_dict_crc(sd) == int(g['weights_crc'][0]), 'random-init weights drifted'
    scenes, metas, fpn = _fixtures.lc_scene(mod.BATCH)
    assert mod.inputs_crc(scenes, metas, fpn) == int(g['inputs_crc'][0]), 'synthetic generator drifted'
    torch.manual_seed(mod.DUMMY_SEED)
    dummies = [torch.rand(1, c).numpy() for c in cfg.multimodal_middle_encoder['in_channels_3D']]
    bev, outs, _ = omodel.extract_voxel_space(sd, cfg, scenes, fpn, metas, dummies)
    check_against_detector_golden(mod, g, bev, [(o.indices, o.features) for o in outs], 1e-5)


@pytest.mark.skipif(not __import__('oracle.ref_detector', fromlist=['x']).available(), reason='reference tree not mounted')
def test_spp_module_matches_reference_class_live():
    """bev_fusion: this package's SPPModule against the reference's own class (MSMDFusion.py:47-90, run in
    place): identical state-dict names and shapes, identical output on the same weights (dense torch ops
    on both sides -- no kernel of this project; it closes the path behind the voxel-space tensor)."""
    from msmdfusion_b200.detector import SPPModule
    from oracle import ref_detector
    torch.manual_seed(4)
    ours = SPPModule().eval()
    ref = ref_detector.spp_module()().eval()
    assert [(k, tuple(v.shape)) for k, v in ours.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    for mod in ours.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(1, 640, 40, 40)
    with torch.no_grad():
        assert torch.equal(ours(x), ref(x))


# --------------------------------------------------------------------------------------
# sparse conv BACKWARD (config 5): the oracle restatement pinned against the reference's vendored
# spconv-1.x backward (live + committed fixtures) and against autograd through a dense conv3d
# --------------------------------------------------------------------------------------
def spconv1x_bwd_goldens():
    return sorted(glob.glob(os.path.join(GOLDEN, 'spconv1xbwd_*.npz')))


def load_bwd_golden(path):
    """(forward fixture, backward fixture, pair_fwd, n_in) of one spconv1xbwd_<name>.npz."""
    name = os.path.basename(path)[len('spconv1xbwd_'):-len('.npz')]
    g = np.load(os.path.join(GOLDEN, f'spconv1x_{name}.npz'))
    b = np.load(path)
    idx = g['indices'].astype(np.int32)
    shape = [int(s) for s in g['spatial_shape']]
    ks, st, pd = [int(x) for x in g['ksize']], [int(x) for x in g['stride']], [int(x) for x in g['padding']]
    if int(g['subm']):
        pair = cpu.subm_rulebook(idx, shape, ks, 1)
    else:
        _, pair, _ = cpu.conv_rulebook(idx, shape, ks, st, pd, 1)
    return g, b, pair, idx.shape[0]


def rel_err(got, ref):
    return float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize('path', spconv1x_bwd_goldens(), ids=os.path.basename)
def test_oracle_conv_backward_matches_reference_spconv1x_golden(path):
    """Fixtures produced by the reference's own ``indice_conv_backward_fp32``
    (tests/golden/make_golden_spconv_bwd.py)."""
    g, b, pair, n_in = load_bwd_golden(path)
    gi, gw = cpu.spconv_bwd(g['features'], g['weight_krsc'], pair, b['grad_out'].astype(np.float32))
    assert rel_err(gi, b['grad_features']) < 5e-6
    assert rel_err(gw, b['grad_weight']) < 5e-6


@pytest.mark.skipif(not os.path.isdir('/root/reference/mmdet3d/ops/spconv/src'), reason='reference tree not mounted')
def test_oracle_conv_backward_matches_reference_spconv1x_live():
    from oracle import ref_spconv
    if ref_spconv.module() is None:
        pytest.skip('reference spconv-1.x could not be built here')
    rng = np.random.default_rng(22)
    for (shape, batch, n, cin, cout, ks, st, pd, subm) in (
            ([7, 18, 18], 2, 700, 8, 12, 3, 1, 1, True), ([7, 18, 18], 2, 700, 8, 16, 3, 2, 1, False),
            ([9, 12, 12], 1, 400, 4, 6, [3, 1, 1], [2, 1, 1], 0, False)):
        idx, feat = random_sparse(rng, batch, shape, n, cin)
        k3 = ks if isinstance(ks, list) else [ks] * 3
        w = rng.standard_normal((cout, *k3, cin)).astype(np.float32) * 0.2
        if subm:
            pair = cpu.subm_rulebook(idx, shape, ks, 1)
        else:
            _, pair, _ = cpu.conv_rulebook(idx, shape, ks, st, pd, 1)
        go = rng.standard_normal((pair.shape[1], cout)).astype(np.float32)
        ri, rw = ref_spconv.conv_backward(idx, feat, w, go, shape, batch, ks, st, pd, 1, subm)
        gi, gw = cpu.spconv_bwd(feat, w, pair, go)
        assert rel_err(gi, ri) < 5e-6 and rel_err(gw, rw) < 5e-6


@pytest.mark.parametrize('subm,ksize,stride,padding', [(True, 3, 1, 1), (False, 3, 2, 1),
                                                       (False, (3, 1, 1), (2, 1, 1), 0), (False, 3, 2, (0, 1, 1))])
def test_conv_backward_vs_dense_autograd(subm, ksize, stride, padding):
    """Independent pin: d/dx, d/dW of a dense conv3d (float64 autograd) sampled at the active rows."""
    rng = np.random.default_rng(3)
    shape, batch, cin, cout = [7, 12, 10], 2, 5, 6
    idx, feat = random_sparse(rng, batch, shape, 300, cin)
    ks = cpu._triple(ksize)
    w = rng.standard_normal((cout, *ks, cin)).astype(np.float32)
    if subm:
        pair, out_idx = cpu.subm_rulebook(idx, shape, ksize, 1), idx
    else:
        out_idx, pair, _ = cpu.conv_rulebook(idx, shape, ksize, stride, padding, 1)
    go = rng.standard_normal((out_idx.shape[0], cout)).astype(np.float32)
    gi, gw = cpu.spconv_bwd(feat, w, pair, go)

    ft = torch.from_numpy(feat).double().requires_grad_(True)
    li = torch.from_numpy(idx.astype(np.int64))
    x = torch.zeros(batch, *shape, cin, dtype=torch.float64)          # channels last, then permute
    x = x.index_put((li[:, 0], li[:, 1], li[:, 2], li[:, 3]), ft).permute(0, 4, 1, 2, 3)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    y = F.conv3d(x, wt.permute(0, 4, 1, 2, 3), stride=cpu._triple(stride), padding=cpu._triple(padding))
    lo = torch.from_numpy(out_idx.astype(np.int64))
    rows = y[lo[:, 0], :, lo[:, 1], lo[:, 2], lo[:, 3]]
    (rows * torch.from_numpy(go).double()).sum().backward()
    assert rel_err(gi, ft.grad.numpy()) < 1e-5
    assert rel_err(gw, wt.grad.numpy()) < 1e-5


def test_pair_transpose_and_dgrad_as_forward():
    """dgrad = the FORWARD contraction over pair_bwd with the weight transposed (Cin <-> Cout): the
    identity the CUDA path relies on to reuse its forward kernels for msmd_spconv_bwd_data.  For SubM
    pair_bwd is pair_fwd with the kernel offsets reversed (o reads i through k  <=>  i reads o through
    K-1-k)."""
    rng = np.random.default_rng(5)
    shape, batch, cin, cout = [7, 14, 12], 2, 6, 9
    idx, feat = random_sparse(rng, batch, shape, 500, cin)
    for subm in (True, False):
        ks = [3, 3, 3]
        w = rng.standard_normal((cout, *ks, cin)).astype(np.float32)
        if subm:
            pair = cpu.subm_rulebook(idx, shape, 3, 1)
        else:
            _, pair, _ = cpu.conv_rulebook(idx, shape, 3, 2, 1, 1)
        pb = cpu.pair_transpose(pair, idx.shape[0])
        K, n_out = pair.shape
        for k in range(K):  # definition
            o = np.nonzero(pair[k] >= 0)[0]
            assert np.array_equal(pb[k, pair[k, o]], o)
            assert (pb[k] >= 0).sum() == o.shape[0]
        if subm:
            assert np.array_equal(pb, pair[::-1])
        go = rng.standard_normal((n_out, cout)).astype(np.float32)
        gi, _ = cpu.spconv_bwd(feat, w, pair, go, need_weight_grad=False)
        wt = np.ascontiguousarray(np.transpose(w, (4, 1, 2, 3, 0)))  # [Cin,kz,ky,kx,Cout]
        assert rel_err(cpu.spconv_fwd(go, wt, pb), gi) < 1e-5


def test_batchnorm_train_matches_torch():
    rng = np.random.default_rng(6)
    x = rng.standard_normal((257, 12)).astype(np.float32) * 3 + 1
    bn = torch.nn.BatchNorm1d(12, eps=1e-3, momentum=0.01).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    y = bn(torch.from_numpy(x)).detach().numpy()
    got, mean, var = cpu.batchnorm_train(x, bn.weight.detach().numpy(), bn.bias.detach().numpy(), 1e-3)
    assert np.abs(got - y).max() < 1e-5
    assert np.abs(mean * 0.01 - bn.running_mean.numpy()).max() < 1e-6


# --------------------------------------------------------------------------------------
# arithmetic model of the opt-in bf16x3 mode (csrc/spconv_tc16.cu): is it inside the parity bound
# over the WHOLE 21-layer LiDAR encoder?  (the kernel itself: tests/test_cuda_emul.py::test_tc16_*)
# --------------------------------------------------------------------------------------
def _bf16_round(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def test_bf16x3_accuracy_model(monkeypatch):
    """SparseEncoder (random spconv-default weights, BN statistics randomised) with every convolution
    computed as A_hi*B_hi + A_lo*B_hi + A_hi*B_lo on bf16 hi/lo parts, against the fp32 oracle: each of the
    five returned tensors and the BEV tensor within 1e-4 of its scale (north_star's bound), with margin.
    The single-MMA bf16 mode is NOT (that is why it is only the train-step arithmetic)."""
    from msmdfusion_b200 import synthetic
    from oracle import model as omodel
    import msmdfusion_b200 as m
    from _fixtures import hotpath_cfg, randomize_bn
    cfg = hotpath_cfg()
    torch.manual_seed(0)
    enc = m.registry.build_middle_encoder(cfg.pts_middle_encoder)
    randomize_bn(enc, 1)
    sd = enc.eval().state_dict()
    scene = synthetic.lidar_scene(5, 1)[:6000]
    ev, en, ec = omodel.voxelize_batch([scene], synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
    emean = cpu.hard_simple_vfe(ev, en, 5)
    ref_sp, ref_feats, _ = omodel.sparse_encoder(sd, dict(cfg.pts_middle_encoder), emean, ec, 1)
    exact = cpu.spconv_fwd

    def x3(features, weight, pair):
        fh, wh = _bf16_round(features), _bf16_round(weight)
        fl, wl = _bf16_round(features - fh), _bf16_round(weight - wh)
        return exact(fl, wh, pair) + exact(fh, wl, pair) + exact(fh, wh, pair)

    def x1(features, weight, pair):
        return exact(_bf16_round(features), _bf16_round(weight), pair)

    def run(fn):
        monkeypatch.setattr(cpu, 'spconv_fwd', fn)
        try:
            sp, feats, _ = omodel.sparse_encoder(sd, dict(cfg.pts_middle_encoder), emean, ec, 1)
        finally:
            monkeypatch.setattr(cpu, 'spconv_fwd', exact)
        errs = [rel_err(a.features, b.features) for a, b in zip(feats, ref_feats)] + [rel_err(sp, ref_sp)]
        return max(errs)

    e3, e1 = run(x3), run(x1)
    assert e3 < 3e-5, e3
    assert e1 > 1e-4, e1


# --------------------------------------------------------------------------------------
# BEV tail: SECOND + SECONDFPN against the reference's own classes run in place
# --------------------------------------------------------------------------------------
def _reference_bev_tail():
    """The reference's SECOND / SECONDFPN classes compiled from their source where it lies; mmcv's layer
    builders are absent, so the three builder names resolve to this package's registry mirrors (plain
    torch.nn layers on both sides -- what is compared is the reference's wiring of them)."""
    import numpy
    from torch import nn
    from msmdfusion_b200 import registry
    from oracle.ref_inplace import load_def
    ns = {'nn': nn, 'torch': torch, 'np': numpy, 'build_conv_layer': registry.build_conv_layer,
          'build_norm_layer': registry.build_norm_layer, 'build_upsample_layer': registry.build_upsample_layer,
          'auto_fp16': lambda *a, **k: (lambda f: f)}
    second = load_def('mmdet3d/models/backbones/second.py', 'SECOND', ns, keyword='class')
    fpn = load_def('mmdet3d/models/necks/second_fpn.py', 'SECONDFPN', ns, keyword='class')
    return second, fpn


def _randomize_bn2d(module, seed):
    g = torch.Generator().manual_seed(seed)
    for mod in module.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data = torch.rand(mod.weight.shape, generator=g) + 0.5
            mod.bias.data = torch.randn(mod.bias.shape, generator=g) * 0.1
            mod.running_mean.data = torch.randn(mod.running_mean.shape, generator=g) * 0.1
            mod.running_var.data = torch.rand(mod.running_var.shape, generator=g) + 0.5


@pytest.mark.skipif(not __import__('oracle.ref_detector', fromlist=['x']).available(), reason='reference tree not mounted')
def test_second_and_secondfpn_match_reference_classes_live():
    """Same constructor kwargs as configs/MSMDFusion_nusc_voxel_LC.py:191-206 (narrower channels to keep it
    quick): identical state-dict names and shapes, bit-identical output on the same weights in the layer-by-layer
    mode (eval and train), and the BN-folded channels-last inference path within 1e-5 of it."""
    import msmdfusion_b200 as m
    ref_second, ref_fpn = _reference_bev_tail()
    bb = dict(in_channels=32, out_channels=[16, 32], layer_nums=[5, 5], layer_strides=[1, 2],
              norm_cfg=dict(type='BN', eps=0.001, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False))
    nk = dict(in_channels=[16, 32], out_channels=[32, 32], upsample_strides=[1, 2],
              norm_cfg=dict(type='BN', eps=0.001, momentum=0.01), upsample_cfg=dict(type='deconv', bias=False),
              use_conv_for_no_stride=True)
    torch.manual_seed(3)
    ours_b = m.registry.build_backbone(dict(type='SECOND', **bb))
    ours_n = m.registry.build_neck(dict(type='SECONDFPN', **nk))
    _randomize_bn2d(ours_b, 1)
    _randomize_bn2d(ours_n, 2)
    rb, rn = ref_second(**bb), ref_fpn(**nk)
    for ours, ref in ((ours_b, rb), (ours_n, rn)):
        assert [(k, tuple(v.shape)) for k, v in ours.state_dict().items()] == \
            [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
        ref.load_state_dict(ours.state_dict())
    x = torch.randn(2, 32, 24, 24)
    for mode in ('eval', 'train'):
        for mod in (ours_b, ours_n, rb, rn):
            getattr(mod, mode)()
        a = ours_n(ours_b(x.clone()))     # grad enabled: layer-by-layer path in both modes
        b = rn(rb(x.clone()))
        assert len(a) == 1 and a[0].shape == (2, 64, 24, 24) and torch.equal(a[0], b[0])
    for mod in (ours_b, ours_n, rb, rn):
        mod.eval()
    with torch.no_grad():
        fused = ours_n(ours_b(x.clone()))[0]
        assert ours_b._folded is not None and ours_n._folded is not None      # the folded path ran
        plain = rn(rb(x.clone()))[0]
    assert rel_err(fused.numpy(), plain.numpy()) < 1e-5
    # the cache follows in-place parameter updates
    with torch.no_grad():
        ours_b.blocks[0][0].weight.mul_(1.5)
        rb.load_state_dict(ours_b.state_dict())
        assert rel_err(ours_n(ours_b(x.clone()))[0].numpy(), rn(rb(x.clone()))[0].numpy()) < 1e-5
    # the other level types: deconv with stride 1 (use_conv_for_no_stride=False) and a down-sampling level (stride 0.5)
    nk2 = dict(in_channels=[16, 32], out_channels=[8, 8], upsample_strides=[0.5, 1],
               norm_cfg=dict(type='BN', eps=0.001, momentum=0.01), upsample_cfg=dict(type='deconv', bias=False))
    torch.manual_seed(5)
    o2, r2 = m.SECONDFPN(**nk2).eval(), ref_fpn(**nk2).eval()
    r2.load_state_dict(o2.state_dict())
    xs = (torch.randn(1, 16, 24, 24), torch.randn(1, 32, 12, 12))
    with torch.no_grad():
        assert rel_err(o2(xs)[0].numpy(), r2(xs)[0].numpy()) < 1e-5


def test_detector_builds_the_bev_tail_from_the_unchanged_config_entries():
    """pts_backbone / pts_neck of the config resolve through the registries and run behind extract_pts_feat's
    BEV tensor: (B, 256, 180, 180) -> SECOND -> SECONDFPN -> (B, 512, 180, 180) (configs/...LC.py:191-206)."""
    import msmdfusion_b200 as m
    bb = m.registry.build_backbone(dict(type='SECOND', in_channels=256, out_channels=[128, 256], layer_nums=[5, 5],
                                        layer_strides=[1, 2], norm_cfg=dict(type='BN', eps=0.001, momentum=0.01),
                                        conv_cfg=dict(type='Conv2d', bias=False))).eval()
    nk = m.registry.build_neck(dict(type='SECONDFPN', in_channels=[128, 256], out_channels=[256, 256],
                                    upsample_strides=[1, 2], norm_cfg=dict(type='BN', eps=0.001, momentum=0.01),
                                    upsample_cfg=dict(type='deconv', bias=False), use_conv_for_no_stride=True)).eval()
    assert sum(p.numel() for p in bb.parameters()) + sum(p.numel() for p in nk.parameters()) == 4_576_768   # "SECOND+FPN ~4.6 M"
    with torch.no_grad():
        out = nk(bb(torch.randn(1, 256, 36, 36)))
    assert out[0].shape == (1, 512, 36, 36)
