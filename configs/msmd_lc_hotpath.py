# Hot-path slice of the reference's configs/MSMDFusion_nusc_voxel_LC.py (values from
# :1-6, :143-190): only the keys the voxel-space fusion path consumes.  The full reference
# config file also loads unchanged through msmdfusion_b200.Config.fromfile.
point_cloud_range = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
voxel_size = [0.075, 0.075, 0.2]

hotpath = dict(
    spatial_shapes=[[41, 1440, 1440], [21, 720, 720], [11, 360, 360], [5, 180, 180]],
    downscale_factors=[1, 2, 4, 8],
    fps_num_list=[2048] * 4,
    radius_list=[6, 3, 2, 1],
    max_cluster_samples_list=[200, 100, 50, 25],
    dist_thresh_list=[13.3, 6.6, 3.3, 1.6],
    pts_voxel_layer=dict(max_num_points=10, voxel_size=voxel_size, max_voxels=(120000, 160000),
                         point_cloud_range=point_cloud_range),
    pts_voxel_encoder=dict(type='HardSimpleVFE', num_features=5),
    pts_middle_encoder=dict(
        type='SparseEncoder', in_channels=5, sparse_shape=[41, 1440, 1440], output_channels=128,
        order=('conv', 'norm', 'act'),
        encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
        encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)),
        block_type='basicblock'),
    multimodal_middle_encoder=dict(
        type='SparseMultiModalEncoderPaint', in_channels_3D=(16, 32, 64, 128),
        in_channels_2D=(64, 64, 64, 64), out_channels=(32, 64, 128, 128),
        padding=(1, 1, [0, 1, 1], 0), order=('conv', 'norm', 'act'),
        norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01)),
)
