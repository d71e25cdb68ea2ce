"""SpMiddleResNetFHD: the CenterPoint/VoxelNet sparse 3-D backbone on the native kernels.

Same constructor, forward signature, topology and state_dict key layout as
det3d/models/backbones/scn.py:37-80 (SparseBasicBlock) and :83-176 (SpMiddleResNetFHD).  21 sparse
convolutions, each one fused kernel launch (conv + bias + eval-BN + residual + ReLU); 4 SubM rulebooks
(keys res0..res3, shared by 5/4/4/4 convs) and 4 strided rulebooks per forward; the final `.dense()` +
`view(N, C*D, H, W)` (scn.py:165-168) is fused into extra_conv's epilogue.
"""
import numpy as np
import torch
from torch import nn

from .registry import BACKBONES
from .sparse import SparseConv3d, SparseConvTensor, SparseModule, SparseSequential, SubMConv3d


def build_norm_1d(norm_cfg, planes):
    """BatchNorm1d from a det3d norm_cfg dict (det3d/models/utils/norm.py:59-108, type 'BN1d')."""
    cfg = dict(norm_cfg)
    kind = cfg.pop("type")
    if kind not in ("BN1d", "BN"):
        raise KeyError("unsupported norm type for sparse features: %s" % kind)
    cfg.setdefault("eps", 1e-5)
    cfg.pop("requires_grad", None)
    return nn.BatchNorm1d(planes, **cfg)


def conv3x3(in_planes, out_planes, stride=1, indice_key=None, bias=True):
    return SubMConv3d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=bias, indice_key=indice_key)


class SparseBasicBlock(SparseModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, norm_cfg=None, downsample=None, indice_key=None):
        super().__init__()
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)
        bias = norm_cfg is not None          # always True after defaulting, as in scn.py:51-54
        self.conv1 = conv3x3(inplanes, planes, stride, indice_key=indice_key, bias=bias)
        self.bn1 = build_norm_1d(norm_cfg, planes)
        self.relu = nn.ReLU()
        self.conv2 = conv3x3(planes, planes, indice_key=indice_key, bias=bias)
        self.bn2 = build_norm_1d(norm_cfg, planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = self.conv1(x, bn=self.bn1, relu=True)
        # relu(bn2(conv2(out)) + identity) in one epilogue (scn.py:71-78)
        return self.conv2(out, bn=self.bn2, residual=identity._feat, relu=True)


@BACKBONES.register_module
class SpMiddleResNetFHD(nn.Module):
    def __init__(self, num_input_features=128, norm_cfg=None, name="SpMiddleResNetFHD", **kwargs):
        super().__init__()
        self.name = name
        self.dcn = None
        self.zero_init_residual = False
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)
        bn = lambda c: build_norm_1d(norm_cfg, c)
        self.conv_input = SparseSequential(
            SubMConv3d(num_input_features, 16, 3, bias=False, indice_key="res0"), bn(16), nn.ReLU(inplace=True))
        self.conv1 = SparseSequential(
            SparseBasicBlock(16, 16, norm_cfg=norm_cfg, indice_key="res0"),
            SparseBasicBlock(16, 16, norm_cfg=norm_cfg, indice_key="res0"))
        self.conv2 = SparseSequential(
            SparseConv3d(16, 32, 3, 2, padding=1, bias=False), bn(32), nn.ReLU(inplace=True),
            SparseBasicBlock(32, 32, norm_cfg=norm_cfg, indice_key="res1"),
            SparseBasicBlock(32, 32, norm_cfg=norm_cfg, indice_key="res1"))
        self.conv3 = SparseSequential(
            SparseConv3d(32, 64, 3, 2, padding=1, bias=False), bn(64), nn.ReLU(inplace=True),
            SparseBasicBlock(64, 64, norm_cfg=norm_cfg, indice_key="res2"),
            SparseBasicBlock(64, 64, norm_cfg=norm_cfg, indice_key="res2"))
        self.conv4 = SparseSequential(
            SparseConv3d(64, 128, 3, 2, padding=[0, 1, 1], bias=False), bn(128), nn.ReLU(inplace=True),
            SparseBasicBlock(128, 128, norm_cfg=norm_cfg, indice_key="res3"),
            SparseBasicBlock(128, 128, norm_cfg=norm_cfg, indice_key="res3"))
        self.extra_conv = SparseSequential(
            SparseConv3d(128, 128, (3, 1, 1), (2, 1, 1), bias=False), bn(128), nn.ReLU())

    def forward(self, voxel_features, coors, batch_size, input_shape, n_dev=None, n_cap=None, out_fmt="fp32"):
        """voxel_features [M,>=5], coors [M,4] (b,z,y,x), input_shape = grid (x,y,z).
        Returns (dense [B, 256, H, W] (channels-last memory), dict of per-stage SparseConvTensors).
        n_dev/n_cap: optional device-resident row count + capacity (fused voxelizer path).
        out_fmt="split" (fused pipeline only) returns the BEV map as an ops.Feat in split bf16 rows [B,H,W,256]."""
        sparse_shape = np.array([int(v) for v in input_shape][::-1]) + [1, 0, 0]
        ret = SparseConvTensor(voxel_features, coors, sparse_shape, batch_size, n_dev=n_dev, n_cap=n_cap)
        x = self.conv_input(ret)
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        # [B, H, W, C*D] channels-last; channel = c*D + d as scn.py:165-168, except on the fused split path where the
        # rows are written d-major (channel = d*C + c, contiguous vector stores) and tagged so that the consumer
        # (RPN) permutes its first conv's input channels accordingly
        dmajor = out_fmt == "split"
        bev = self.extra_conv(x_conv4, bev_last=True, out_fmt=out_fmt, bev_dmajor=dmajor)
        if dmajor:
            c_out = self.extra_conv[0].out_channels
            bev.bev_dmajor = (c_out, bev.ctot // c_out)
        ret = bev if out_fmt == "split" else bev.permute(0, 3, 1, 2)    # logical [B, C*D, H, W]
        return ret, {"conv1": x_conv1, "conv2": x_conv2, "conv3": x_conv3, "conv4": x_conv4}
