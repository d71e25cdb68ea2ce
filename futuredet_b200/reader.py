"""VoxelFeatureExtractorV3 (det3d/models/readers/voxel_encoder.py:9-24): mean of the points of a voxel.

In the fused path (VoxelNet.forward_points) the mean is produced by the voxelizer kernel itself and this
module is bypassed; `forward` keeps the reference signature for callers that hold padded voxels."""
import torch
from torch import nn

from . import ops
from .registry import READERS


@READERS.register_module
class VoxelFeatureExtractorV3(nn.Module):
    def __init__(self, num_input_features=4, norm_cfg=None, name="VoxelFeatureExtractorV3"):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors=None):
        assert self.num_input_features == features.shape[-1]
        return ops.vfe_mean(features, num_voxels)
