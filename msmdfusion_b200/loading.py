"""Virtual-point wire format -> packed foreground scene (host side, numpy/torch CPU).

Mirror of the reference's foreground-2D pipeline stages
(``mmdet3d/datasets/pipelines/my_loading_multi_proj.py``): ``LoadForeground2D`` (:14-160),
``LoadForeground2DFromMultiSweeps`` (:162-337), ``GlobalRotTransFilterForeground2D`` (:341-416),
``ImgScaleCropFlipForeground2D`` (:419-455), ``ShuffleForeground2D`` (:457-489) -- same registered names,
same constructor kwargs, same keys read from / written to the ``results`` dict, results identical bit
for bit (``tests/test_loading.py`` runs the reference's own classes in place beside these).

This is dataset-side host code by definition -- the reference's stages are numpy on DataLoader
workers -- not a CPU fallback of any kernel: nothing here has a CUDA counterpart that it stands in for.

What is different is the memory plan.  The reference keeps six Python lists of per-camera arrays and
grows them by ``np.concatenate`` once per sweep and camera (every merge copies everything merged so
far, in float64), then wraps each camera in a ``LiDARPoints``.  Here one sample is ONE set of packed,
camera-major arrays sized once from the file headers' row counts:

    pixels       (M, 3)  f32   u, v, depth          rows of camera c: offsets[c] : offsets[c+1]
    points       (M, 15) f32   xyz | 10 one-hot + score | dt      (torch tensor sharing the memory)
    real_pixels  (R, 3)  f32   real_points (R, 15) f64            rows: real_offsets[c] : real_offsets[c+1]

and every stage works on those arrays (the per-sweep rigid transform is done in float64 on the
sweep's own rows, exactly as the reference does, and rounded to float32 once on the way into the
packed buffer).  ``results['foreground2D_info']`` stays a dict with the reference's four keys -- lists of
per-camera VIEWS into the packed arrays, ``fg_points[c].tensor`` included -- so every consumer of the
reference format (``MSMDFusionDetector.get_foreground2D``, ``MyCollect3D``) keeps working; the extra key
``'packed'`` holds the :class:`ForegroundScene`, which is already the layout ``detector.PackedForeground``
uploads.
"""
import os

import numpy as np
import torch

from .registry import Registry

PIPELINES = Registry('pipeline')      # mmdet.datasets.builder.PIPELINES
FOREGROUND_DIR = 'FOREGROUND_MIXED_6NN_WITH_DEPTH'   # my_loading_multi_proj.py:128,307
POINT_DIM = 15


class CameraPoints:
    """What the path reads of a ``LiDARPoints`` (core/points/base_points.py:25-43): ``.tensor``."""
    __slots__ = ('tensor', 'points_dim')

    def __init__(self, tensor):
        self.tensor = tensor
        self.points_dim = tensor.shape[-1]

    def __len__(self):
        return self.tensor.shape[0]


class ForegroundInfo(dict):
    """``results['foreground2D_info']``: the reference's keys (per-camera views) + ``'packed'``.  Pickles as
    the packed scene alone (DataLoader workers hand samples over by pickle): the views are rebuilt on the
    other side instead of being serialised a second time as copies."""

    def __reduce__(self):
        return (_info_from_scene, (self['packed'],))


def _info_from_scene(scene):
    return scene.reference_dict()


class ForegroundScene:
    """Packed, camera-major foreground info of one sample (see the module docstring)."""

    def __init__(self, ncam, pixels, points, offsets, real_pixels, real_points, real_offsets):
        self.ncam = ncam
        self.pixels, self.points, self.offsets = pixels, points, offsets
        self.real_pixels, self.real_points, self.real_offsets = real_pixels, real_points, real_offsets

    def cam_ids(self):
        """(M,) int32 camera id of every row (what PackedForeground uploads as ``cam``)."""
        return np.repeat(np.arange(self.ncam, dtype=np.int32), np.diff(self.offsets))

    def reference_dict(self):
        """The reference's ``foreground2D_info`` dict: per-camera views, no copies."""
        o, r = self.offsets, self.real_offsets
        cams = range(self.ncam)
        return ForegroundInfo({'fg_pixels': [self.pixels[o[c]:o[c + 1]] for c in cams],
                               'fg_points': [CameraPoints(self.points[o[c]:o[c + 1]]) for c in cams],
                               'fg_real_pixels': [self.real_pixels[r[c]:r[c + 1]] for c in cams],
                               'fg_real_points': [self.real_points[r[c]:r[c + 1]] for c in cams],
                               'packed': self})

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('key_frame', None)     # the raw key-frame dict is only needed between the two load stages
        return state


def scene_of(results):
    info = results['foreground2D_info']
    scene = info.get('packed') if isinstance(info, dict) else None
    if scene is None:
        raise KeyError("results['foreground2D_info'] was not produced by msmdfusion_b200.loading")
    return scene


def foreground_path(pts_filename):
    """my_loading_multi_proj.py:126-128 -- kept as written, including the loss of a leading '/'."""
    tokens = pts_filename.split('/')
    return os.path.join(*tokens[:-2], FOREGROUND_DIR, tokens[-1] + '.pkl.npy')


def read_wire(path):
    """One sweep's saved dict (four keys, lists of per-camera float32 arrays)."""
    return np.load(path, allow_pickle=True).item()


class _Part:
    """One sweep's contribution: the raw dict plus what the merge needs (None = the key frame)."""
    __slots__ = ('raw', 'dt', 'dt_real', 'rot', 'trans')

    def __init__(self, raw, dt=0.0, dt_real=0.0, rot=None, trans=None):
        self.raw, self.dt, self.dt_real, self.rot, self.trans = raw, dt, dt_real, rot, trans


def pack_parts(parts):
    """Key frame + sweeps -> ForegroundScene, one allocation per array.

    Per camera the row order is the reference's: for each part in order, its virtual rows, then its real
    rows (:65,77 / :190,203 and the merge order of :255-275)."""
    ncam = len(parts[0].raw['virtual_pixel_indices'])
    parts = [p for p in parts if len(p.raw['virtual_pixel_indices']) == ncam]    # :249, else skipped
    nv = np.array([[p.raw['virtual_pixel_indices'][c].shape[0] for p in parts] for c in range(ncam)], np.int64)
    nr = np.array([[p.raw['real_pixel_indices'][c].shape[0] for p in parts] for c in range(ncam)], np.int64)
    offsets = np.concatenate([[0], np.cumsum((nv + nr).sum(1))])
    real_offsets = np.concatenate([[0], np.cumsum(nr.sum(1))])
    M, R = int(offsets[-1]), int(real_offsets[-1])
    pixels = np.empty((M, 3), np.float32)
    points = np.empty((M, POINT_DIM), np.float32)
    real_pixels = np.empty((R, 3), np.float32)
    real_points = np.empty((R, POINT_DIM), np.float64)
    dts = np.array([p.dt for p in parts], np.float64)
    dts_real = np.array([p.dt_real for p in parts], np.float64)
    for c in range(ncam):
        # one C-level concatenation per destination block instead of one slice assignment per
        # (sweep, source array): the lists below are views / small per-sweep temporaries
        pix, lab, xyz, rpix, rlab, rxyz = [], [], [], [], [], []
        for p in parts:
            raw = p.raw
            vp, rp = raw['virtual_pixel_indices'][c], raw['real_pixel_indices'][c]
            vq, rq = raw['virtual_points'][c], raw['real_points'][c]
            pix += [vp[:, :3], rp[:, :3]]
            rpix.append(rp[:, :3])
            if vq.shape[1] == 3:                    # :71 "append label after xyz"
                lab += [vp[:, -11:], rp[:, -11:]]
                rlab.append(rp[:, -11:])
            else:
                lab += [vq[:, 3:14], rq[:, 3:14]]
                rlab.append(rq[:, 3:14])
            if p.rot is None:
                xyz += [vq[:, :3], rq[:, :3]]
                rxyz.append(rq[:, :3])
            else:
                # :262-263 / :273-274 -- float64, on the same operand shapes as the reference: the merged
                # (virtual + real) set for fg_points, the real set on its own for fg_real_points
                both = np.concatenate([vq[:, :3], rq[:, :3]], 0).astype(np.float64) @ p.rot.T
                xyz.append(both + p.trans)
                rxyz.append(rq[:, :3].astype(np.float64) @ p.rot.T + p.trans)
        sl, rs = slice(int(offsets[c]), int(offsets[c + 1])), slice(int(real_offsets[c]), int(real_offsets[c + 1]))
        np.concatenate(pix, 0, out=pixels[sl])
        np.concatenate(lab, 0, out=points[sl, 3:14])
        np.concatenate(xyz, 0, out=points[sl, :3], casting='same_kind')     # float64 -> float32 here
        points[sl, 14] = np.repeat(dts, nv[c] + nr[c])                      # (base_points.py:30)
        np.concatenate(rpix, 0, out=real_pixels[rs])
        np.concatenate(rlab, 0, out=real_points[rs, 3:14], casting='same_kind')
        np.concatenate(rxyz, 0, out=real_points[rs, :3], casting='same_kind')
        real_points[rs, 14] = np.repeat(dts_real, nr[c])
    return ForegroundScene(ncam, pixels, torch.from_numpy(points), offsets, real_pixels, real_points, real_offsets)


@PIPELINES.register_module()
class LoadForeground2D:
    """my_loading_multi_proj.py:14-160 (nuScenes branch).  Reads the key frame's wire file; the result is
    already a packed scene (one part)."""

    def __init__(self, dataset='NuScenesDataset', **kwargs):
        self.dataset = dataset

    def __call__(self, results):
        if self.dataset != 'NuScenesDataset':
            raise NotImplementedError('foreground2D info of {} dataset is unavailable!'.format(self.dataset))
        raw = read_wire(foreground_path(results['pts_filename']))
        scene = pack_parts([_Part(raw)])
        scene.key_frame = raw           # the multi-sweep stage packs once more, from the raw parts
        results['foreground2D_info'] = scene.reference_dict()
        return results


@PIPELINES.register_module()
class LoadForeground2DFromMultiSweeps:
    """my_loading_multi_proj.py:162-337.  ``test_mode`` is read by the reference (:300) but never set by
    its constructor; it is a keyword here (default False = the training-time random choice)."""

    def __init__(self, dataset='NuScenesDataset', sweeps_num=10, test_mode=False):
        self.dataset = dataset
        self.sweeps_num = sweeps_num
        self.test_mode = test_mode

    def __call__(self, results):
        if self.dataset != 'NuScenesDataset':
            return None                  # the reference falls off the end of __call__ (:296-337)
        key = scene_of(results)
        raw_key = getattr(key, 'key_frame', None)
        if raw_key is None:
            raise KeyError('LoadForeground2DFromMultiSweeps must follow LoadForeground2D')
        sweeps = results['sweeps']
        if len(sweeps) <= self.sweeps_num:
            choices = np.arange(len(sweeps))
        elif self.test_mode:
            choices = np.arange(self.sweeps_num)
        else:
            choices = np.random.choice(len(sweeps), self.sweeps_num, replace=False)
        ts = results['timestamp']
        parts = [_Part(raw_key)]
        for idx in choices:
            sweep = sweeps[idx]
            path = foreground_path(sweep['data_path'])
            if not os.path.exists(path):
                continue
            sweep_ts = sweep['timestamp'] / 1e6
            parts.append(_Part(read_wire(path), dt=ts - sweep_ts, dt_real=ts - sweep_ts / 1e-6,   # :206,215
                               rot=np.asarray(sweep['sensor2lidar_rotation']),
                               trans=np.asarray(sweep['sensor2lidar_translation'])))
        results['foreground2D_info'] = pack_parts(parts).reference_dict()
        return results


def _rot_mat_T(t, rotation, axis=2):
    """BasePoints.rotate (base_points.py:77-116) for LiDARPoints (rotation_axis = 2)."""
    if not isinstance(rotation, torch.Tensor):
        rotation = t.new_tensor(rotation)
    assert rotation.shape == torch.Size([3, 3]) or rotation.numel() == 1
    if rotation.numel() == 1:
        rot_sin, rot_cos = torch.sin(rotation), torch.cos(rotation)
        return rotation.new_tensor([[rot_cos, -rot_sin, 0], [rot_sin, rot_cos, 0], [0, 0, 1]]).T
    return rotation


@PIPELINES.register_module()
class GlobalRotTransFilterForeground2D:
    """my_loading_multi_proj.py:341-416: replay the point-cloud augmentation flow on the virtual points,
    then keep the rows inside ``point_cloud_range`` (pixels follow; the real pixels are not filtered)."""

    def __init__(self, point_cloud_range=None):
        self.pcd_range = np.array(point_cloud_range, dtype=np.float32) if point_cloud_range else None

    def __call__(self, input_dict):
        scene = scene_of(input_dict)
        t, o = scene.points, scene.offsets
        rot = input_dict['pcd_rotation'] if 'pcd_rotation' in input_dict else np.eye(3)
        scale = input_dict['pcd_scale_factor'] if 'pcd_scale_factor' in input_dict else 1.
        trans = input_dict['pcd_trans'] if 'pcd_trans' in input_dict else np.zeros(3)
        hflip = input_dict['pcd_horizontal_flip'] if 'pcd_horizontal_flip' in input_dict else False
        vflip = input_dict['pcd_vertical_flip'] if 'pcd_vertical_flip' in input_dict else False
        flow = input_dict['transformation_3d_flow'] if 'transformation_3d_flow' in input_dict else []
        for op in flow:
            assert op in ('T', 'S', 'R', 'HF', 'VF'), f'This 3D data transformation op ({op}) is not supported'
        for c in range(scene.ncam):       # per camera, so every matmul has the reference's operand shapes
            v = t[o[c]:o[c + 1]]
            for op in flow:
                if op == 'T':
                    tv = trans if isinstance(trans, torch.Tensor) else v.new_tensor(trans)
                    v[:, :3] += tv.squeeze(0)
                elif op == 'S':
                    if not (isinstance(scale, (int, float)) and scale == 1):   # x * 1 is x, bit for bit
                        v[:, :3] *= scale
                elif op == 'R':
                    v[:, :3] = v[:, :3] @ _rot_mat_T(v, rot)
                elif op == 'HF' and hflip:
                    v[:, 1] = -v[:, 1]
                elif op == 'VF' and vflip:
                    v[:, 0] = -v[:, 0]
        if isinstance(self.pcd_range, (list, tuple, np.ndarray)):
            r = self.pcd_range        # base_points.py:143-166 (this fork's -0.0001 on the upper bounds)
            tn = t.numpy()
            xyz = np.ascontiguousarray(tn[:, :3])
            lo = r[:3]
            hi = np.array([r[3] - 0.0001, r[4] - 0.0001, r[5] - 0.0001], np.float32)   # float32, as np.float32 - float is
            inside = (xyz > lo) & (xyz < hi)
            keep_np = inside[:, 0] & inside[:, 1] & inside[:, 2]
            rows = np.flatnonzero(keep_np)
            csum = np.concatenate([[0], np.cumsum(keep_np)])
            counts = csum[o[1:]] - csum[o[:-1]]
            scene = ForegroundScene(scene.ncam, scene.pixels.take(rows, axis=0),
                                    torch.from_numpy(tn.take(rows, axis=0)),
                                    np.concatenate([[0], np.cumsum(counts)]), scene.real_pixels,
                                    scene.real_points, scene.real_offsets)
        input_dict['foreground2D_info'] = scene.reference_dict()
        return input_dict


@PIPELINES.register_module()
class ImgScaleCropFlipForeground2D:
    """my_loading_multi_proj.py:419-455: replay the image resize / crop / flip on the pixel coordinates."""

    def __init__(self, **kwargs):
        pass

    def __call__(self, input_dict):
        scene = scene_of(input_dict)
        img_scale_factor = input_dict['scale_factor'][:2] if 'scale_factor' in input_dict else [1., 1.]
        img_flip = input_dict['flip'] if 'flip' in input_dict else False
        img_crop_offset = input_dict['img_crop_offset'] if 'img_crop_offset' in input_dict else 0
        img_shape = input_dict['img_shape'][:2]
        for pix in (scene.pixels, scene.real_pixels):
            pix[:, :2] = pix[:, :2] * img_scale_factor
            if not (np.isscalar(img_crop_offset) and img_crop_offset == 0):   # x - 0 is x, bit for bit
                pix -= img_crop_offset      # all three columns, depth included, as the reference does (:448)
            if img_flip:
                orig_h, orig_w = img_shape
                pix[:, 0] = orig_w - 1 - pix[:, 0]
        input_dict['foreground2D_info'] = scene.reference_dict()
        return input_dict


@PIPELINES.register_module()
class ShuffleForeground2D:
    """my_loading_multi_proj.py:457-489: one ``torch.randperm`` per camera, in camera order, applied to the
    camera's pixels and points alike."""

    def __init__(self, **kwargs):
        pass

    def __call__(self, input_dict):
        scene = scene_of(input_dict)
        o = scene.offsets
        pix = torch.from_numpy(scene.pixels)
        for c in range(scene.ncam):
            n = int(o[c + 1] - o[c])
            perm = torch.randperm(n)
            pix[o[c]:o[c + 1]] = pix[o[c]:o[c + 1]][perm]
            scene.points[o[c]:o[c + 1]] = scene.points[o[c]:o[c + 1]][perm]
        input_dict['foreground2D_info'] = scene.reference_dict()
        return input_dict


def build_pipeline(cfgs):
    """[dict(type=..., **kwargs)] -> stage objects (mmdet ``Compose`` without the rest of mmdet)."""
    from .registry import build_from_cfg
    return [build_from_cfg(c, PIPELINES) for c in cfgs]


def run_pipeline(stages, results):
    for stage in stages:
        results = stage(results)
        if results is None:
            return None
    return results
