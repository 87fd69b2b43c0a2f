"""SURVEY §8(f) rank 3 -- virtual-point wire format and loader (host side, CPU only).

msmdfusion_b200.loading against the reference's OWN pipeline classes
(mmdet3d/datasets/pipelines/my_loading_multi_proj.py) run in place by oracle/ref_loading.py on synthetic
wire files: every array of `foreground2D_info` bit for bit."""
import copy
import os
import zlib

import numpy as np
import pytest
import torch

from msmdfusion_b200 import loading, synthetic
from oracle import ref_loading

needs_reference = pytest.mark.skipif(not ref_loading.available(), reason='reference tree not mounted')
SCALE = np.array([0.5, 0.49777778, 0.5, 0.49777778], np.float32)    # 1600x900 -> 800x448 (keep_ratio resize)


@pytest.fixture()
def wire(tmp_path, monkeypatch):
    """Relative data root (the reference's path logic drops a leading '/', :126-128)."""
    monkeypatch.chdir(tmp_path)

    def make(**kw):
        return synthetic.write_foreground_wire('data', **kw)
    return make


def pipeline_meta(results, train=False):
    results = dict(results)
    results.update(scale_factor=SCALE, img_shape=(448, 800, 3))
    if train:
        results.update(transformation_3d_flow=['R', 'S', 'T', 'HF', 'VF'], pcd_rotation=torch.tensor(
            [[0.9553365, -0.29552022, 0.0], [0.29552022, 0.9553365, 0.0], [0.0, 0.0, 1.0]]).T,
            pcd_scale_factor=1.07, pcd_trans=np.array([0.3, -0.2, 0.05]), pcd_horizontal_flip=True,
            pcd_vertical_flip=False, flip=True, img_crop_offset=0)
    else:   # what GlobalRotScaleTrans with zero ranges + RandomFlip3D(no flip) leave behind
        results.update(transformation_3d_flow=['R', 'S', 'T'], pcd_rotation=torch.eye(3), pcd_scale_factor=1.0,
                       pcd_trans=np.zeros(3), pcd_horizontal_flip=False, pcd_vertical_flip=False, flip=False)
    return results


def ours_test_pipeline(sweeps_num=10, test_mode=True):
    return loading.build_pipeline([
        dict(type='LoadForeground2D', dataset='NuScenesDataset'),
        dict(type='LoadForeground2DFromMultiSweeps', dataset='NuScenesDataset', sweeps_num=sweeps_num,
             test_mode=test_mode),
        dict(type='GlobalRotTransFilterForeground2D', point_cloud_range=synthetic.POINT_CLOUD_RANGE),
        dict(type='ImgScaleCropFlipForeground2D')])


def assert_same_info(ref, got):
    assert len(ref['fg_pixels']) == len(got['fg_pixels']) == 6
    for c in range(6):
        for key in ('fg_pixels', 'fg_real_pixels', 'fg_real_points'):
            a, b = ref[key][c], got[key][c]
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), (key, c)
        a, b = ref['fg_points'][c].tensor, got['fg_points'][c].tensor
        assert a.dtype == b.dtype == torch.float32 and a.shape == b.shape and torch.equal(a, b), ('fg_points', c)


@needs_reference
@pytest.mark.parametrize('sweeps,missing,empty', [(10, (3,), (4,)), (0, (), ()), (4, (0, 1, 2, 3), (0, 5))])
def test_inference_pipeline_matches_reference(wire, sweeps, missing, empty):
    results = pipeline_meta(wire(seed=sweeps, sweeps=sweeps, missing_sweeps=missing, empty_cameras=empty,
                             virtual_per_camera=700, real_per_camera=120))
    ref = ref_loading.run(ref_loading.test_pipeline(synthetic.POINT_CLOUD_RANGE), copy.deepcopy(results))
    got = loading.run_pipeline(ours_test_pipeline(), copy.deepcopy(results))
    assert_same_info(ref['foreground2D_info'], got['foreground2D_info'])
    scene = got['foreground2D_info']['packed']
    assert scene.pixels.shape[0] == scene.offsets[-1] == sum(p.shape[0] for p in ref['foreground2D_info']['fg_pixels'])
    if sweeps == 10:   # the range filter removed something, and only from the (virtual + real) set
        raw = sum(np.load(os.path.join('data/samples', loading.FOREGROUND_DIR, f), allow_pickle=True).item()
                  ['virtual_pixel_indices'][0].shape[0] for f in os.listdir(os.path.join('data/samples', loading.FOREGROUND_DIR)))
        assert 0 < scene.offsets[1] - scene.real_offsets[1] < raw


@needs_reference
def test_training_pipeline_matches_reference(wire):
    """Augmentation flow replay (rotate, scale, translate, flips), image flip and the per-camera shuffle."""
    results = pipeline_meta(wire(seed=21, sweeps=3, virtual_per_camera=500, real_per_camera=90), train=True)
    c = ref_loading.classes()
    ref_stages = [c['LoadForeground2D'](dataset='NuScenesDataset'),
                  c['LoadForeground2DFromMultiSweeps'](dataset='NuScenesDataset', sweeps_num=10),
                  c['GlobalRotTransFilterForeground2D'](point_cloud_range=synthetic.POINT_CLOUD_RANGE),
                  c['ImgScaleCropFlipForeground2D'](), c['ShuffleForeground2D']()]
    our_stages = ours_test_pipeline(test_mode=False) + loading.build_pipeline([dict(type='ShuffleForeground2D')])
    torch.manual_seed(5)
    ref = ref_loading.run(ref_stages, copy.deepcopy(results))
    torch.manual_seed(5)
    got = loading.run_pipeline(our_stages, copy.deepcopy(results))
    assert_same_info(ref['foreground2D_info'], got['foreground2D_info'])


@needs_reference
@pytest.mark.parametrize('test_mode', [True, False])
def test_sweep_choice_matches_reference(wire, test_mode):
    """More sweeps on disk than `sweeps_num`: the first ones at test time, np.random.choice in training."""
    results = pipeline_meta(wire(seed=33, sweeps=7, virtual_per_camera=200, real_per_camera=40))
    c = ref_loading.classes()
    multi = c['LoadForeground2DFromMultiSweeps'](dataset='NuScenesDataset', sweeps_num=4)
    multi.test_mode = test_mode
    np.random.seed(11)
    ref = ref_loading.run([c['LoadForeground2D'](dataset='NuScenesDataset'), multi], copy.deepcopy(results))
    np.random.seed(11)
    got = loading.run_pipeline(ours_test_pipeline(sweeps_num=4, test_mode=test_mode)[:2], copy.deepcopy(results))
    ref_info = ref['foreground2D_info']
    for cam in range(6):   # before the later stages fg_pixels are still float32 arrays on both sides
        assert np.array_equal(ref_info['fg_pixels'][cam], got['foreground2D_info']['fg_pixels'][cam])
        assert torch.equal(ref_info['fg_points'][cam].tensor, got['foreground2D_info']['fg_points'][cam].tensor)


def scene_crc(info):
    crc = 0
    for key in ('fg_pixels', 'fg_real_pixels', 'fg_real_points'):
        for a in info[key]:
            crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
    for p in info['fg_points']:
        crc = zlib.crc32(np.ascontiguousarray(p.tensor.numpy()).tobytes(), crc)
    return crc


GOLDEN_CRC = 1701617899   # printed by `python tests/test_loading.py` (reference classes run in place)


def test_inference_pipeline_golden_crc(wire):
    """Travels without the reference: CRC of the reference pipeline's output on the seeded wire files
    (generated by `python tests/test_loading.py`, which prints it)."""
    results = pipeline_meta(wire(seed=1, sweeps=10, missing_sweeps=(3,), empty_cameras=(4,), virtual_per_camera=700,
                             real_per_camera=120))
    got = loading.run_pipeline(ours_test_pipeline(), results)
    assert scene_crc(got['foreground2D_info']) == GOLDEN_CRC


def test_packed_scene_is_what_the_detector_uploads(wire):
    """The per-camera lists are views of the packed arrays, and detector.PackedForeground packs the dict
    into exactly those arrays (CPU path)."""
    from msmdfusion_b200.detector import PackedForeground
    results = pipeline_meta(wire(seed=2, sweeps=2, virtual_per_camera=300, real_per_camera=50, empty_cameras=(1,)))
    info = loading.run_pipeline(ours_test_pipeline(), results)['foreground2D_info']
    scene = info['packed']
    assert all(np.shares_memory(v, scene.pixels) for v in info['fg_pixels'] if v.size)
    assert all(p.tensor.untyped_storage().data_ptr() == scene.points.untyped_storage().data_ptr()
               for p in info['fg_points'] if len(p))
    meta = dict(foreground2D_info=info, lidar2img=synthetic.camera_matrices(0))
    pk = PackedForeground([meta], 'cpu')
    assert np.array_equal(pk.pixels.numpy(), scene.pixels) and torch.equal(pk.points, scene.points)
    assert np.array_equal(pk.cam.numpy(), scene.cam_ids())
    assert np.array_equal(pk.real_pixels.numpy(), scene.real_pixels)
    assert pk.counts == [int(scene.offsets[-1])]
    # the same dict without the packed scene goes through the per-camera branch: identical result
    plain = {k: v for k, v in info.items() if k != 'packed'}
    pk2 = PackedForeground([dict(meta, foreground2D_info=plain), dict(meta, foreground2D_info=info)], 'cpu')
    m = pk.pixels.shape[0]
    assert torch.equal(pk2.pixels[:m], pk.pixels) and torch.equal(pk2.pixels[m:], pk.pixels)
    assert torch.equal(pk2.points[:m], pk.points) and torch.equal(pk2.points[m:], pk.points)
    assert torch.equal(pk2.cam[:m], pk.cam) and torch.equal(pk2.cam[m:], pk.cam + 6)
    r = pk.real_pixels.shape[0]
    assert torch.equal(pk2.real_pixels[r:], pk.real_pixels) and torch.equal(pk2.real_cam[r:], pk.real_cam + 6)
    assert torch.equal(pk2.real_cam[:r], pk.real_cam) and pk2.counts == [m, m]


def test_foreground_info_pickles_as_one_packed_scene(wire):
    """DataLoader workers pickle samples: the dict travels as the packed arrays only and comes back with
    its per-camera views rebuilt on them."""
    import pickle
    info = loading.run_pipeline(ours_test_pipeline(), pipeline_meta(wire(seed=4, sweeps=2, virtual_per_camera=400,
                                                                         real_per_camera=60)))['foreground2D_info']
    blob = pickle.dumps(info, protocol=pickle.HIGHEST_PROTOCOL)
    scene = info['packed']
    payload = scene.pixels.nbytes + scene.points.numel() * 4 + scene.real_pixels.nbytes + scene.real_points.nbytes
    assert len(blob) < 1.05 * payload + 4096            # not twice: the views are not serialised
    back = pickle.loads(blob)
    assert isinstance(back, loading.ForegroundInfo) and scene_crc(back) == scene_crc(info)
    assert all(np.shares_memory(v, back['packed'].pixels) for v in back['fg_pixels'] if v.size)
    assert not hasattr(back['packed'], 'key_frame')


def test_registry_and_errors(wire):
    for name in ('LoadForeground2D', 'LoadForeground2DFromMultiSweeps', 'GlobalRotTransFilterForeground2D',
                 'ImgScaleCropFlipForeground2D', 'ShuffleForeground2D'):
        assert name in loading.PIPELINES
    with pytest.raises(NotImplementedError):
        loading.LoadForeground2D(dataset='LyftDataset')({'pts_filename': 'a/b/c.bin'})
    with pytest.raises(FileNotFoundError):
        loading.LoadForeground2D()({'pts_filename': 'data/samples/LIDAR_TOP/absent.pcd.bin'})
    with pytest.raises(KeyError):
        loading.ImgScaleCropFlipForeground2D()({'foreground2D_info': {'fg_pixels': []}, 'img_shape': (1, 1, 3)})


if __name__ == '__main__':   # regenerate GOLDEN_CRC through the reference's own classes
    import tempfile
    os.chdir(tempfile.mkdtemp())
    res = pipeline_meta(synthetic.write_foreground_wire('data', seed=1, sweeps=10, missing_sweeps=(3,), empty_cameras=(4,),
                                                    virtual_per_camera=700, real_per_camera=120))
    out = ref_loading.run(ref_loading.test_pipeline(synthetic.POINT_CLOUD_RANGE), res)
    print('GOLDEN_CRC =', scene_crc(out['foreground2D_info']))
