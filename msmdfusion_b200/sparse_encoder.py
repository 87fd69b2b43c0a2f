"""SparseEncoder -- mirrors ``mmdet3d/models/middle_encoders/sparse_encoder.py:10-209``
(same constructor kwargs, sub-module names and return value)."""
from torch import nn

import torch

from . import executor, ops, spconv
from .registry import MIDDLE_ENCODERS
from .sparse_block import SparseBasicBlock, make_sparse_convmodule


@MIDDLE_ENCODERS.register_module()
class SparseEncoder(nn.Module):

    def __init__(self, in_channels, sparse_shape, order=('conv', 'norm', 'act'),
                 norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), base_channels=16,
                 output_channels=128,
                 encoder_channels=((16, ), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                 encoder_paddings=((1, ), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                 block_type='conv_module'):
        super().__init__()
        assert block_type in ['conv_module', 'basicblock']
        self.sparse_shape = sparse_shape
        self.in_channels = in_channels
        self.order = tuple(order)
        self.base_channels = base_channels
        self.output_channels = output_channels
        self.encoder_channels = encoder_channels
        self.encoder_paddings = encoder_paddings
        self.stage_num = len(self.encoder_channels)
        self.fp16_enabled = False
        assert isinstance(self.order, tuple) and len(self.order) == 3
        assert set(self.order) == {'conv', 'norm', 'act'}

        if self.order[0] != 'conv':  # pre activate
            self.conv_input = make_sparse_convmodule(in_channels, self.base_channels, 3,
                                                     norm_cfg=norm_cfg, padding=1,
                                                     indice_key='subm1', conv_type='SubMConv3d',
                                                     order=('conv', ))
        else:  # post activate
            self.conv_input = make_sparse_convmodule(in_channels, self.base_channels, 3,
                                                     norm_cfg=norm_cfg, padding=1,
                                                     indice_key='subm1', conv_type='SubMConv3d')
        encoder_out_channels = self.make_encoder_layers(make_sparse_convmodule, norm_cfg,
                                                        self.base_channels, block_type=block_type)
        self.conv_out = make_sparse_convmodule(encoder_out_channels, self.output_channels,
                                               kernel_size=(3, 1, 1), stride=(2, 1, 1),
                                               norm_cfg=norm_cfg, padding=0,
                                               indice_key='spconv_down2', conv_type='SparseConv3d')

    # ``use_executor``: run the whole chain through the native executor (csrc/executor.cu) -- one
    # C-ABI call instead of ~55 -- whenever the module tree is in the fused-inference form
    # (CUDA, no grad, eval-mode BatchNorm).  The module-by-module path below is the general one.
    use_executor = True

    def _plan_for(self):
        cached = getattr(self, '_plan', None)
        if cached is not None and len(cached) == 4:
            key = cached[3].key()
            if key is not None and cached[0] == key:
                return cached[1], cached[2]
        mods = [self.conv_input, self.encoder_layers, self.conv_out]
        watch = executor.PlanWatch(mods)
        key = watch.key()
        plan = executor.SparseNetPlan()
        cur = plan.add(self.conv_input, 0)
        marks = [cur]
        for encoder_layer in self.encoder_layers:
            cur = plan.add(encoder_layer, cur)
            marks.append(cur)
        marks.append(plan.add(self.conv_out, cur))
        plan.finalize()
        self._plan = (key, plan, marks, watch)
        return plan, marks

    def forward(self, voxel_features, coors, batch_size):
        """sparse_encoder.py:96-133 -> (spatial_features (B, C*D, H, W), encode_features)."""
        coors = coors.int()
        self.ran_on_executor = False   # True: the geometry of this call ran on the executor's geometry stream
        if self.use_executor and voxel_features.is_cuda and not torch.is_grad_enabled():
            try:
                plan, marks = self._plan_for()
            except executor.Unsupported:
                plan = None
            if plan is not None:
                B = int(batch_size)
                acts = plan.run(voxel_features, coors, self.sparse_shape, B)
                encode_features = [spconv.SparseConvTensor(acts[i][0], acts[i][1], acts[i][2], B)
                                   for i in marks[:-1]]
                f, idx, shape = acts[marks[-1]]
                spatial_features = ops.to_dense(idx, f, shape, B)
                N, C, D, H, W = spatial_features.shape
                self.ran_on_executor = True
                return spatial_features.view(N, C * D, H, W), encode_features
        x = spconv.SparseConvTensor(voxel_features, coors, self.sparse_shape, batch_size)
        x = self.conv_input(x)
        encode_features = [x]
        for encoder_layer in self.encoder_layers:
            x = encoder_layer(x)
            encode_features.append(x)
        out = self.conv_out(encode_features[-1])
        spatial_features = out.dense()
        N, C, D, H, W = spatial_features.shape
        spatial_features = spatial_features.view(N, C * D, H, W)
        return spatial_features, encode_features

    def make_encoder_layers(self, make_block, norm_cfg, in_channels, block_type='conv_module',
                            conv_cfg=dict(type='SubMConv3d')):
        """sparse_encoder.py:135-209."""
        assert block_type in ['conv_module', 'basicblock']
        self.encoder_layers = spconv.SparseSequential()
        for i, blocks in enumerate(self.encoder_channels):
            blocks_list = []
            for j, out_channels in enumerate(tuple(blocks)):
                padding = tuple(self.encoder_paddings[i])[j]
                if i != 0 and j == 0 and block_type == 'conv_module':
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg,
                                                  stride=2, padding=padding,
                                                  indice_key=f'spconv{i + 1}',
                                                  conv_type='SparseConv3d'))
                elif block_type == 'basicblock':
                    if j == len(blocks) - 1 and i != len(self.encoder_channels) - 1:
                        blocks_list.append(make_block(in_channels, out_channels, 3,
                                                      norm_cfg=norm_cfg, stride=2, padding=padding,
                                                      indice_key=f'spconv{i + 1}',
                                                      conv_type='SparseConv3d'))
                    else:
                        blocks_list.append(SparseBasicBlock(out_channels, out_channels,
                                                            norm_cfg=norm_cfg, conv_cfg=conv_cfg))
                else:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg,
                                                  padding=padding, indice_key=f'subm{i + 1}',
                                                  conv_type='SubMConv3d'))
                in_channels = out_channels
            stage_name = f'encoder_layer{i + 1}'
            self.encoder_layers.add_module(stage_name, spconv.SparseSequential(*blocks_list))
        return out_channels
