"""MSMDFusionDetector / TransFusionDetector -- the voxel-space part of the reference detectors.

Mirrors ``mmdet3d/models/detectors/MSMDFusion.py:92-452`` (and ``transfusion.py:61-74``,
``mvx_two_stage.py:22-97`` for construction): same registry names, constructor kwargs, method
names and parameter names (``conv1x1_blocks``, ``score_net``, ``bev_fusion``).  In scope: everything from
``points`` + FPN image features to the fused BEV tensor.  The image backbone / neck, the BEV
backbone / neck and the head are NOT part of this path: they are built only when their type is
registered by the host project and are otherwise kept as config dicts.

B200-first differences that cannot change results:
* the per-(sample, camera) ``torch.from_numpy(...).to(device)`` copies of
  ``get_foreground2D`` / ``depth_aware_channel_compression`` (``:202-209``, ``:349-350``) are one
  packed pinned-memory upload per batch, reused by the four scales;
* pixel-feature gather x score gate is one kernel over all cameras (``ops.lift_gather``);
* ``hard_voxelize`` + ``HardSimpleVFE`` are fused, so the (160000, 10, 64) zero-filled staging
  buffer (410 MB per call) never exists;
* ``voxel_modality_split`` runs on the device (stable radix sort + binary-search merge) instead
  of GPU sort -> CPU numba merge -> GPU.
"""
import contextlib
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops, registry, spconv
from .registry import DETECTORS
from .voxel import Voxelization


class SPPModule(nn.Module):
    """``MSMDFusion.py:47-90`` (dense BEV fusion; plain cuDNN convolutions)."""

    def __init__(self):
        super().__init__()

        def branch(cin, k, pad, dil):
            return nn.Sequential(
                nn.Conv2d(cin, 256, kernel_size=k, stride=1, padding=pad, dilation=dil, bias=False),
                nn.BatchNorm2d(256, eps=0.001, momentum=0.01), nn.ReLU())
        self.conv1x1 = branch(384 + 256, 1, 0, 1)
        self.conv3x3 = branch(384 + 256, 3, 1, 1)
        self.dilated_conv3x3_rate6 = branch(384 + 256, 3, 6, 6)
        self.dilated_conv3x3_rate12 = branch(384 + 256, 3, 12, 12)
        self.fuse = branch(256 * 4, 1, 0, 1)

    def forward(self, x):
        return self.fuse(torch.cat([self.conv1x1(x), self.conv3x3(x), self.dilated_conv3x3_rate6(x),
                                    self.dilated_conv3x3_rate12(x)], dim=1))


def _as_numpy(a, dtype=np.float32):
    if hasattr(a, 'tensor'):  # LiDARPoints (core/points/base_points.py:25-30)
        a = a.tensor
    if torch.is_tensor(a):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dtype)


class PackedForeground:
    """All cameras' virtual / real foreground points of a batch in one upload.

    pixels (M,3) u,v,depth; cam (M,) flat camera id sample*ncam+cam; points (M,P);
    lidar2img (B*ncam,16); real_pixels (R,3); real_cam (R,); counts[b] = points of sample b.
    Order = samples, then cameras, then the per-camera order: what the reference's nested
    concatenation (``MSMDFusion.py:189-226``) produces.
    """

    def __init__(self, img_metas, device, ncam=None, staging=None):
        B = len(img_metas)
        pix, pts, cam, l2i, rpix, rcam, counts = [], [], [], [], [], [], []
        for b, meta in enumerate(img_metas):
            info = meta['foreground2D_info']
            n = len(info['fg_pixels']) if ncam is None else ncam
            total = 0
            scene = info.get('packed') if isinstance(info, dict) else None
            if scene is not None and scene.ncam == n:
                # produced by msmdfusion_b200.loading: already camera-major packed arrays (the per-camera
                # lists are views of them) -- one block per array instead of one per camera
                pix.append(scene.pixels)
                pts.append(_as_numpy(scene.points))
                cam.append(scene.cam_ids() + np.int32(b * n))
                rpix.append(scene.real_pixels)
                rcam.append(np.repeat(np.arange(n, dtype=np.int32), np.diff(scene.real_offsets)) + np.int32(b * n))
                for v in range(n):
                    l2i.append(np.asarray(meta['lidar2img'][v], np.float64).reshape(16).astype(np.float32))
                counts.append(int(scene.offsets[-1]))
                self.ncam = n
                continue
            for v in range(n):
                p = _as_numpy(info['fg_pixels'][v]).reshape(-1, 3)
                q = _as_numpy(info['fg_points'][v])
                q = q.reshape(p.shape[0], -1) if p.shape[0] else q.reshape(0, q.shape[-1] if q.ndim > 1 else 15)
                pix.append(p)
                pts.append(q)
                cam.append(np.full((p.shape[0],), b * n + v, np.int32))
                l2i.append(np.asarray(meta['lidar2img'][v], np.float64).reshape(16).astype(np.float32))
                total += p.shape[0]
                if 'fg_real_pixels' in info:
                    r = _as_numpy(info['fg_real_pixels'][v]).reshape(-1, 3)
                    rpix.append(r)
                    rcam.append(np.full((r.shape[0],), b * n + v, np.int32))
            counts.append(total)
            self.ncam = n
        self.batch_size = B
        self.counts = counts
        # persistent pinned staging buffers (grow-only, owned by the detector): cudaHostAlloc per
        # call costs milliseconds and synchronises the device
        self._staging = staging if staging is not None else {}
        ev = self._staging.get('event')
        if ev is not None:
            ev.synchronize()  # the previous upload has left the staging buffers
        pdim = pts[0].shape[1] if pts else 15
        self.pixels = self._upload('pixels', pix, 3, np.float32, device)
        self.points = self._upload('points', pts, pdim, np.float32, device)
        self.cam = self._upload('cam', cam, None, np.int32, device)
        self.lidar2img = self._upload('lidar2img', l2i, None, np.float32, device).view(-1, 16)
        self.real_pixels = self._upload('real_pixels', rpix, 3, np.float32, device)
        self.real_cam = self._upload('real_cam', rcam, None, np.int32, device)
        if torch.device(device).type == 'cuda':
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            self._staging['event'] = ev
        self.h2d_bytes = sum(int(t.numel()) * t.element_size() for t in
                             (self.pixels, self.points, self.cam, self.lidar2img, self.real_pixels,
                              self.real_cam))

    def _upload(self, name, arrs, width, dtype, device):
        """Concatenate ``arrs`` DIRECTLY into a persistent pinned staging buffer (no temporary host
        array, no per-call cudaHostAlloc) and start the asynchronous upload."""
        rows = sum(a.shape[0] for a in arrs)
        shape = (rows,) if width is None else (rows, width)
        numel = rows * (1 if width is None else width)
        if numel == 0 or torch.device(device).type != 'cuda':
            a = np.concatenate(arrs, 0) if arrs else np.zeros(shape, dtype)
            return torch.from_numpy(np.ascontiguousarray(a.astype(dtype, copy=False))).to(device)
        tdtype = torch.float32 if dtype == np.float32 else torch.int32
        buf = self._staging.get(name)
        if buf is None or buf.numel() < numel or buf.dtype != tdtype:
            buf = self._staging[name] = torch.empty((int(numel * 1.25) + 16,), dtype=tdtype).pin_memory()
        stage = buf[:numel].view(shape)
        np.concatenate([np.asarray(a, dtype).reshape((-1,) + shape[1:]) for a in arrs], 0, out=stage.numpy())
        return stage.to(device, non_blocking=True)


def _maybe_build(cfg, builder):
    """Build a sub-module outside this path only if the host project registered its type."""
    if cfg is None:
        return None
    try:
        return builder(cfg)
    except KeyError:
        return None


class _VoxelPathMixin:
    """Construction shared by both detectors (``mvx_two_stage.py:22-97``)."""

    def _build_common(self, pts_voxel_layer, pts_voxel_encoder, pts_middle_encoder,
                      multimodal_middle_encoder, pts_backbone, pts_neck, pts_bbox_head, img_backbone,
                      img_neck, train_cfg, test_cfg):
        if pts_voxel_layer:
            self.pts_voxel_layer = Voxelization(**pts_voxel_layer)
        if pts_voxel_encoder:
            self.pts_voxel_encoder = registry.build_voxel_encoder(pts_voxel_encoder)
        if pts_middle_encoder:
            self.pts_middle_encoder = registry.build_middle_encoder(pts_middle_encoder)
        if multimodal_middle_encoder:
            self.multimodal_middle_encoder = registry.build_middle_encoder(multimodal_middle_encoder)
        self.out_of_path_cfg = dict(pts_backbone=pts_backbone, pts_neck=pts_neck,
                                    pts_bbox_head=pts_bbox_head, img_backbone=img_backbone,
                                    img_neck=img_neck)
        for name, builder in (('pts_backbone', registry.build_backbone), ('pts_neck', registry.build_neck),
                              ('img_backbone', registry.build_backbone), ('img_neck', registry.build_neck)):
            mod = _maybe_build(self.out_of_path_cfg[name], builder)
            if mod is not None:
                setattr(self, name, mod)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

    @property
    def with_pts_backbone(self):
        return hasattr(self, 'pts_backbone') and self.pts_backbone is not None

    @property
    def with_pts_neck(self):
        return hasattr(self, 'pts_neck') and self.pts_neck is not None

    @torch.no_grad()
    def voxelize(self, points, downscale_factor=1.0):
        """``MSMDFusion.py:462-491`` / ``mvx_two_stage.py`` voxelize: per-sample hard_voxelize,
        batch id padded in front of the coordinates.  Returns (voxels, num_points, coors)."""
        self.pts_voxel_layer.voxel_size = [0.075, 0.075, 0.2]  # hard-coded reset (:475)
        self.pts_voxel_layer.voxel_size = [x * downscale_factor for x in self.pts_voxel_layer.voxel_size]
        voxels, coors, num_points = [], [], []
        for i, res in enumerate(points):
            v, c, n = self.pts_voxel_layer(res)
            voxels.append(v)
            coors.append(F.pad(c, (1, 0), mode='constant', value=i))
            num_points.append(n)
        return torch.cat(voxels, 0), torch.cat(num_points, 0), torch.cat(coors, 0)

    @torch.no_grad()
    def voxelize_mean(self, points, num_features, downscale_factor=1.0):
        """voxelize + HardSimpleVFE fused: (mean (V,F), coors (V,4) int32, per-sample counts)."""
        self.pts_voxel_layer.voxel_size = [0.075, 0.075, 0.2]
        self.pts_voxel_layer.voxel_size = [x * downscale_factor for x in self.pts_voxel_layer.voxel_size]
        means, coors, counts = [], [], []
        for i, res in enumerate(points):
            m, c, _ = self.pts_voxel_layer.forward_mean(res, num_features, batch_idx=i)
            means.append(m)
            coors.append(c)
            counts.append(c.shape[0])
        if len(means) == 1:
            return means[0], coors[0], counts
        return torch.cat(means, 0), torch.cat(coors, 0), counts


@DETECTORS.register_module()
class TransFusionDetector(nn.Module, _VoxelPathMixin):
    """LiDAR-only detector (``transfusion.py:18-102``): voxelize -> VFE -> SparseEncoder."""

    def __init__(self, freeze_img=True, pts_voxel_layer=None, pts_voxel_encoder=None,
                 pts_middle_encoder=None, pts_fusion_layer=None, img_backbone=None, pts_backbone=None,
                 img_neck=None, pts_neck=None, pts_bbox_head=None, img_roi_head=None,
                 img_rpn_head=None, train_cfg=None, test_cfg=None, pretrained=None, **kwargs):
        nn.Module.__init__(self)
        self.freeze_img = freeze_img
        self._build_common(pts_voxel_layer, pts_voxel_encoder, pts_middle_encoder, None, pts_backbone,
                           pts_neck, pts_bbox_head, img_backbone, img_neck, train_cfg, test_cfg)

    def extract_pts_feat(self, pts, img_feats=None, img_metas=None):
        """``transfusion.py:61-74``."""
        nf = min(self.pts_voxel_encoder.num_features, pts[0].shape[1])
        voxel_features, coors, _ = self.voxelize_mean(pts, nf)
        x, _ = self.pts_middle_encoder(voxel_features, coors, len(pts))
        if self.with_pts_backbone:
            x = self.pts_backbone(x)
            if self.with_pts_neck:
                x = self.pts_neck(x)
        return x


@DETECTORS.register_module()
class MSMDFusionDetector(nn.Module, _VoxelPathMixin):
    """``MSMDFusion.py:92-452``."""

    def __init__(self, freeze_img=True, discard_views=[], pts_voxel_layer=None, pts_voxel_encoder=None,
                 pts2D_voxel_encoder=None, pts_middle_encoder=None, pseudo_pts_middle_encoder=None,
                 multimodal_middle_encoder=None, pts_fusion_layer=None, img_backbone=None,
                 pts_backbone=None, img_neck=None, pts_neck=None, pts_bbox_head=None,
                 img_roi_head=None, img_rpn_head=None, train_cfg=None, test_cfg=None, pretrained=None,
                 spatial_shapes=None, downscale_factors=None, fps_num_list=None, radius_list=None,
                 max_cluster_samples_list=None, dist_thresh_list=None, **kwargs):
        nn.Module.__init__(self)
        self.freeze_img = freeze_img
        self.discard_views = discard_views
        self._build_common(pts_voxel_layer, pts_voxel_encoder, pts_middle_encoder,
                           multimodal_middle_encoder, pts_backbone, pts_neck, pts_bbox_head,
                           img_backbone, img_neck, train_cfg, test_cfg)
        self.spatial_shapes = spatial_shapes
        self.downscale_factors = downscale_factors
        self.fps_num_list = fps_num_list
        self.radius_list = radius_list
        self.max_cluster_samples_list = max_cluster_samples_list
        self.dist_thresh_list = dist_thresh_list

        def compress(k):  # channel compression for the FPN levels (:108-124)
            return nn.Sequential(
                nn.Conv2d(256 + 1, 49, kernel_size=k, stride=1, padding=k // 2, bias=False),
                nn.BatchNorm2d(49, eps=0.001, momentum=0.01), nn.ReLU())
        self.conv1x1_blocks = nn.ModuleList([compress(5), compress(5), compress(3)])
        self.score_net = nn.Sequential(nn.Linear(50 + 16, 1), nn.ReLU())
        self.bev_fusion = SPPModule()
        self._packed = (None, None)

    # -- packed host->device upload, shared by the four scales ----------------------------
    def packed_foreground(self, img_metas, device):
        key, val = self._packed
        if key is img_metas and val is not None and val.pixels.device == torch.device(device):
            return val
        if not hasattr(self, '_staging'):
            self._staging = {}
        val = PackedForeground(img_metas, device, staging=self._staging)
        self._packed = (img_metas, val)
        return val

    def _score_bias(self):
        """score_net bias as a host float, read back once per parameter version (not per call)."""
        b = self.score_net[0].bias
        if b is None:
            return 0.0
        key = (b.data_ptr(), b._version)
        if getattr(self, '_bias_cache', (None, 0.0))[0] != key:
            self._bias_cache = (key, float(b.detach().item()))
        return self._bias_cache[1]

    def _xyz_normalizer(self, device):
        t = getattr(self, '_xyz_norm', None)
        if t is None or t.device != device:
            t = self._xyz_norm = torch.tensor([13.5, 13.5, 2.0], device=device)  # :388
        return t

    # -- :169-238 ---------------------------------------------------------------------------
    def get_foreground2D(self, img_feats, img_metas):
        """Per sample: (M_b, 15 + C) = [virtual point | gated C-channel pixel feature]."""
        B = len(img_metas)
        BN, C, H, W = img_feats.shape
        downscale_factor = img_feats.shape[-1] / img_metas[0]['input_shape'][-1]
        pk = self.packed_foreground(img_metas, img_feats.device)
        lin = self.score_net[0]
        out = ops.lift_gather(img_feats.float(), pk.pixels, pk.cam, pk.points, pk.lidar2img,
                              downscale_factor, lin.weight, self._score_bias())
        return list(torch.split(out, pk.counts, dim=0)) if B > 1 else [out]

    # -- :335-369 ---------------------------------------------------------------------------
    def depth_aware_channel_compression(self, feat_list, img_metas, _static_ok=False):
        B = len(img_metas)
        device = feat_list[0].device
        H, W = img_metas[0]['pad_shape'][:2]
        pk = self.packed_foreground(img_metas, device)
        ncanvas = B * pk.ncam
        canvas = torch.zeros(ncanvas * H * W, device=device)
        if pk.real_pixels.shape[0]:
            coors = pk.real_pixels[:, :2].long()  # truncation toward zero (:351)
            lin = (pk.real_cam.long() * H + coors[:, 1]) * W + coors[:, 0]
            # index_put_ with duplicate pixels (:356): the last real point in input order wins
            order = torch.arange(lin.shape[0], device=device)
            winner = torch.full_like(canvas, -1, dtype=torch.long).scatter_reduce_(
                0, lin, order, reduce='amax', include_self=True)
            depth = pk.real_pixels[:, 2].contiguous()
            canvas = torch.where(winner >= 0, depth[winner.clamp_min(0)], canvas)  # no host sync
        canvas = canvas.view(ncanvas, 1, H, W)
        # `_static_ok`: the caller consumes the result within the step (extract_voxel_space); everybody else gets
        # tensors of their own
        if (_static_ok and self.compress_graph and canvas.is_cuda and not torch.is_grad_enabled() and
                not self.conv1x1_blocks.training and all(f.is_cuda and f.dtype == torch.float32 for f in feat_list[:3])):
            return self._compress_graphed(feat_list[:3], canvas)
        return self._compress_static(feat_list, canvas)

    def _compress_static(self, feat_list, canvas):
        """The shape-static half of the block (:360-369): depth map resampled to each FPN level, concatenated,
        Conv2d(257 -> 49) + BatchNorm2d + ReLU (cuDNN)."""
        out = []
        for i in range(3):
            img_feat = feat_list[i]
            h, w = img_feat.shape[-2:]
            sp_depth_map = F.interpolate(canvas, (h, w), mode='bilinear')
            out.append(self.conv1x1_blocks[i](torch.cat([img_feat, sp_depth_map], dim=1)))
        return out

    # Those ~20 launches have the same shapes every step, and the step is bound by the host issuing launches
    # (DESIGN.md 3d): in the inference form they are captured once into a CUDA graph (inputs copied into the graph's
    # static buffers: 45 MB of device-to-device copies, ~30 us) and replayed with one launch.  Same cuDNN kernels on
    # the same operands.  MSMD_COMPRESS_GRAPH=0 switches it off; the graph is rebuilt when a shape or a parameter's
    # storage changes (parameter VALUES are read at replay, so in-place updates need nothing).
    compress_graph = os.environ.get('MSMD_COMPRESS_GRAPH', '1') not in ('', '0')

    def _compress_graphed(self, feat_list, canvas):
        key = (tuple((tuple(f.shape), f.device) for f in feat_list), tuple(canvas.shape),
               tuple(t.data_ptr() for t in list(self.conv1x1_blocks.parameters()) + list(self.conv1x1_blocks.buffers())))
        rec = self.__dict__.get('_compress_graph_rec')
        if rec is None or rec['key'] != key:
            cur = torch.cuda.current_stream(canvas.device)
            static_in = [torch.empty_like(f, memory_format=torch.contiguous_format) for f in feat_list]
            static_canvas = torch.empty_like(canvas)
            for dst, src in zip(static_in, feat_list):
                dst.copy_(src)
            static_canvas.copy_(canvas)
            self._compress_static(static_in, static_canvas)     # lazy cuDNN initialisation happens outside the capture
            cur.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):   # other threads / streams keep working
                static_out = self._compress_static(static_in, static_canvas)
            rec = self.__dict__['_compress_graph_rec'] = dict(key=key, graph=graph, static_in=static_in,
                                                              static_canvas=static_canvas, static_out=static_out)
        for dst, src in zip(rec['static_in'], feat_list):
            dst.copy_(src)
        rec['static_canvas'].copy_(canvas)
        rec['graph'].replay()
        # the graph owns its outputs and rewrites them at the next replay: consumers are stream-ordered before it (the
        # side stream waits for an event of the main stream at the start of every step)
        return list(rec['static_out'])

    # -- :371-393 ---------------------------------------------------------------------------
    def fetch_2D_voxels(self, img_feat, img_metas, voxel_size, downscale_factor, B):
        batch_fg = self.get_foreground2D(img_feat, img_metas)
        for i in range(B):
            if batch_fg[i].shape[0] == 0:  # empty sample -> 100 all-zero points (:376-380)
                batch_fg[i] = torch.zeros(100, batch_fg[i].shape[1], device=batch_fg[i].device)
        feat_dim = batch_fg[0].shape[-1]
        self.pts_voxel_encoder.num_features = feat_dim  # never restored in the reference (:386)
        fg_voxel_features, fg_coors, _ = self.voxelize_mean(batch_fg, feat_dim, downscale_factor)
        xyz_normalizer = self._xyz_normalizer(fg_voxel_features.device)
        fg_voxel_features[:, :3] = fg_voxel_features[:, :3] / xyz_normalizer[None, :]
        return spconv.SparseConvTensor(fg_voxel_features, fg_coors, voxel_size, B)

    # -- :251-325 ---------------------------------------------------------------------------
    def voxel_modality_split(self, voxel_3D, voxel_2D, B):
        """Marks voxels present in both modalities: indices become (b, mix, z, y, x); returns the
        row ids of the matched pairs (sorted-key order, previous-sample offset quirk kept)."""
        coord_3D, coord_2D = voxel_3D.indices.int(), voxel_2D.indices.int()
        if B == 1:
            n3 = [coord_3D.shape[0]]
            n2 = [coord_2D.shape[0]]
        else:
            n3 = torch.bincount(coord_3D[:, 0].long(), minlength=B)[:B].tolist()
            n2 = torch.bincount(coord_2D[:, 0].long(), minlength=B)[:B].tolist()
        mix3, mix2, syn3, syn2 = [], [], [], []
        o3 = o2 = 0
        last3 = last2 = 0
        for i in range(B):
            # rows of a sample are contiguous (per-sample voxelize / ascending conv outputs)
            m3, m2, s3, s2 = ops.modality_split_single(coord_3D[o3:o3 + n3[i]], coord_2D[o2:o2 + n2[i]],
                                                       offset3=last3, offset2=last2)
            mix3.append(m3); mix2.append(m2); syn3.append(s3); syn2.append(s2)
            o3 += n3[i]; o2 += n2[i]
            last3, last2 = n3[i], n2[i]  # previous sample's length only (:294-295,:313-314)
        cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs, 0))
        m3, m2 = cat(mix3), cat(mix2)
        voxel_3D.indices = torch.cat([coord_3D[:, :1], m3[:, None], coord_3D[:, 1:]], dim=1)
        voxel_2D.indices = torch.cat([coord_2D[:, :1], m2[:, None], coord_2D[:, 1:]], dim=1)
        # kept for SparseMultiModalEncoderPaint's sync-free group selection (one sample per GPU)
        voxel_3D._mix, voxel_3D._bzyx = m3, coord_3D
        voxel_2D._mix, voxel_2D._bzyx = m2, coord_2D
        return voxel_3D, voxel_2D, cat(syn3), cat(syn2)

    # -- :400-418 ---------------------------------------------------------------------------
    def extract_multiscale_voxel_feat(self, img_feats, encode_features, img_metas, spatial_shapes,
                                      downscale_factors, batch_size, compressed=None):
        img_feats = compressed if compressed is not None else self.depth_aware_channel_compression(img_feats, img_metas)
        img_feat_list = [img_feats[0]] + list(img_feats)
        v3l, v2l, s3l, s2l = [], [], [], []
        for i in range(4):
            voxel_2D = self.fetch_2D_voxels(img_feat_list[i], img_metas, spatial_shapes[i],
                                            downscale_factors[i], batch_size)
            v3, v2, s3, s2 = self.voxel_modality_split(encode_features[i], voxel_2D, batch_size)
            v3l.append(v3); v2l.append(v2); s3l.append(s3); s2l.append(s2)
            enc = getattr(self, 'multimodal_middle_encoder', None)
            if enc is not None and hasattr(enc, 'prelaunch_assign') and self.fps_num_list is not None:
                # coordinates of this scale are final: start its FPS / NN-assignment chain now
                enc.prelaunch_assign(i, v3, v2, s3.shape[0], self.fps_num_list[i], self.radius_list[i],
                                     self.max_cluster_samples_list[i], self.dist_thresh_list[i])
        return v3l, v2l, s3l, s2l

    # -- :421-452 ---------------------------------------------------------------------------
    # Inference schedule (CUDA, no grad, LiDAR encoder on the native executor): everything on the image side of the
    # fusion -- 4 x (lift, voxelize, modality split) and the FPS / nearest-voxel assignment chains (2047 serial
    # rounds each) -- depends on the LiDAR branch through voxel COORDINATES only.  Those are complete on the
    # executor's geometry stream ~0.4 ms into the encoder call, long before its 21 convolutions have run, so that
    # work is issued on a side stream that waits for the geometry and overlaps the LiDAR convolutions instead of
    # queueing behind them (the reference runs all of it in sequence, MSMDFusion.py:421-445).  Same kernels, same
    # operands: results are unchanged.  (A three-strand variant that also starts the FPS chains from a
    # coordinates-only voxelisation was measured slower, profiles/r02h_lc_timeline.txt: the step is bound by the host
    # issuing ~400 launches, and the extra voxelisation calls cost more host time than the earlier FPS start saves.)
    overlap_image_side = os.environ.get('MSMD_LC_OVERLAP', '1') not in ('', '0')   # A/B switch

    # The compression block is ~35 small torch launches (0.7 ms of host time, profiles/r02k_lc_timeline_fps512.txt) that
    # depend on nothing the calling thread does next (LiDAR voxelisation + encoder call, ~1 ms of host time, most of it
    # inside C calls that release the GIL): a helper thread issues them on the side stream meanwhile.  Same stream, same
    # kernels, same order on that stream as before -- only the host work overlaps.  Worth 0.1 ms of the step (r02s), and
    # nothing once the block's static half is a CUDA graph (r02v: 7.441 vs 7.444 ms): OFF by default, MSMD_LC_HOST_THREAD=1
    # switches it on.
    host_thread = os.environ.get('MSMD_LC_HOST_THREAD', '0') not in ('', '0')

    def _helper(self):
        ex = self.__dict__.get('_host_helper')
        if ex is None:
            from concurrent.futures import ThreadPoolExecutor
            ex = self.__dict__['_host_helper'] = ThreadPoolExecutor(max_workers=1, thread_name_prefix='msmd-image-side')
        return ex

    def _side(self, device):
        st = self.__dict__.get('_image_side_stream')
        if st is None or st.device != device:
            st = self.__dict__['_image_side_stream'] = torch.cuda.Stream(device=device)
        return st

    def extract_voxel_space(self, pts, img_feats, img_metas):
        """The voxel-space fusion hot path: returns the (B, 256 + 384, 180, 180) BEV tensor that
        ``bev_fusion`` consumes, plus the multimodal stage outputs."""
        batch_size = len(pts)
        nf = min(self.pts_voxel_encoder.num_features, pts[0].shape[1])  # [:64] of 5 dims after :386
        dev = pts[0].device
        overlap = (self.overlap_image_side and dev.type == 'cuda' and not torch.is_grad_enabled() and
                   getattr(self.pts_middle_encoder, 'use_executor', False))
        compressed = None
        main = torch.cuda.current_stream(dev) if overlap else None
        if overlap:
            # the compression convolutions (cuDNN, ~1.2 ms) depend on nothing from the LiDAR branch: they open the
            # side stream, so the LiDAR voxelisation's read-back on `main` does not wait behind them
            side = self._side(dev)
            start = torch.cuda.Event()
            start.record(main)
            side.wait_event(start)
            pending = None
            if self.host_thread:
                with torch.cuda.stream(side):
                    self.packed_foreground(img_metas, dev)   # the upload (and its cache entry) from this thread

                def compress():
                    with torch.no_grad(), torch.cuda.stream(side):   # grad mode and current stream are per thread
                        return self.depth_aware_channel_compression(img_feats, img_metas, _static_ok=True)
                pending = self._helper().submit(compress)
            else:
                with torch.cuda.stream(side):
                    compressed = self.depth_aware_channel_compression(img_feats, img_metas, _static_ok=True)
        voxel_features, coors, _ = self.voxelize_mean(pts, nf)
        # a frozen LiDAR encoder (tools/train.py:185-211) has no grad-requiring input either (voxelize is
        # no_grad, :462-464), so autograd would skip it anyway: run it on the inference path
        frozen = torch.is_grad_enabled() and not any(p.requires_grad for p in self.pts_middle_encoder.parameters())
        with (torch.no_grad() if frozen else contextlib.nullcontext()):
            x, encode_features = self.pts_middle_encoder(voxel_features, coors, batch_size)
        if overlap:
            from . import executor
            if pending is not None:
                compressed = pending.result()   # re-raises whatever the helper thread raised
            geom_done = torch.cuda.Event()
            if getattr(self.pts_middle_encoder, 'ran_on_executor', False):
                geom_done.record(executor.geometry_stream(dev))   # index sets of the four LiDAR scales
            else:
                geom_done.record(main)
            side.wait_event(geom_done)
            # Allocator note: tensors created under `side` and read later on `main` are safe without record_stream.
            # `main` waits for `image_side_done` before it reads them, and the side stream only ever starts a step's
            # work after waiting for an event recorded on `main` (start), i.e. after every main-stream reader of the
            # previous step's blocks has been queued AND finished before the blocks can be rewritten.
            with torch.cuda.stream(side):
                v3l, v2l, s3l, s2l = self.extract_multiscale_voxel_feat(
                    img_feats, encode_features, img_metas, self.spatial_shapes, self.downscale_factors, batch_size,
                    compressed=compressed)
                image_side_done = torch.cuda.Event()
                image_side_done.record(side)
            main.wait_event(image_side_done)
        else:
            v3l, v2l, s3l, s2l = self.extract_multiscale_voxel_feat(
                img_feats, encode_features, img_metas, self.spatial_shapes, self.downscale_factors, batch_size,
                compressed=compressed)
        stage_outs = self.multimodal_middle_encoder(
            v3l, v2l, s3l, s2l, self.fps_num_list, self.radius_list, self.max_cluster_samples_list,
            self.dist_thresh_list)
        multimodal_out_dense = stage_outs[-1].dense()
        N, C, D, H, W = multimodal_out_dense.shape
        x_mm = multimodal_out_dense.view(N, C * D, H, W)
        return torch.cat([x, x_mm], dim=1), stage_outs

    def extract_pts_feat(self, pts, img_feats, img_metas):
        x, _ = self.extract_voxel_space(pts, img_feats, img_metas)
        x = self.bev_fusion(x)
        if self.with_pts_backbone:
            x = self.pts_backbone(x)
            if self.with_pts_neck:
                x = self.pts_neck(x)
        return x
