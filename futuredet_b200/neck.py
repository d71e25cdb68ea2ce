"""RPN BEV neck on the native fused conv+BN+ReLU kernels.

Same constructor arguments, module tree and state_dict keys as det3d/models/necks/rpn.py:22-160
(blocks.{i}.{0: ZeroPad2d, 1: Conv2d, 2: BN, 3: ReLU, 4,7,..: Conv2d, 5,8,..: BN}, deblocks.{i}.{0,1}).
The torch modules are parameter containers only: forward() runs one native kernel per conv with the
eval-mode BatchNorm and ReLU folded into the epilogue, activations channels-last, and the deblocks write
straight into channel slices of the concatenated output (torch.cat of rpn.py:156-157 is free).
"""
import logging

import numpy as np
import torch
from torch import nn

from . import ops
from . import precision as _precision
from .registry import NECKS
from .sparse import folded_epilogue


def build_norm_2d(norm_cfg, planes):
    cfg = dict(norm_cfg)
    kind = cfg.pop("type")
    if kind != "BN":
        raise KeyError("unsupported norm type for the neck: %s" % kind)
    cfg.setdefault("eps", 1e-5)
    cfg.pop("requires_grad", None)
    return nn.BatchNorm2d(planes, **cfg)


def conv_weight_kio(conv):
    """[K, Cin, Cout] view of a Conv2d / ConvTranspose2d weight, cached until the parameter changes."""
    w = conv.weight
    key = (w.data_ptr(), w._version)
    cache = conv.__dict__.setdefault("_kio_cache", {})
    if cache.get("k") != key:
        with torch.no_grad():
            if isinstance(conv, nn.ConvTranspose2d):     # [Cin, Cout, kh, kw]
                v = w.detach().permute(2, 3, 0, 1)
            else:                                        # [Cout, Cin, kh, kw]
                v = w.detach().permute(2, 3, 1, 0)
            cache["v"] = v.reshape(-1, v.shape[2], v.shape[3]).contiguous().float()
        cache["k"] = key
    return cache["v"]


def act_fmt(precision=None):
    """Inter-layer activation format of a precision: fp32 rows, or split bf16 hi/lo rows on the tensor-core arm."""
    return _precision.act_fmt(precision)


def conv_weight_kio_dmajor(conv, C, D):
    """Weight for an input whose channels arrive d-major (j = d*C + c) instead of the reference's c*D + d."""
    w = conv_weight_kio(conv)
    # keyed on the parameter itself: `w` is a derived temporary whose address the allocator recycles between
    # optimizer steps (a stale hit would run the first neck conv with pre-update weights)
    key = (conv.weight.data_ptr(), conv.weight._version, C, D)
    cache = conv.__dict__.setdefault("_kio_dmajor_cache", {})
    if cache.get("k") != key:
        j = torch.arange(C * D, device=w.device)
        cache["v"] = w[:, (j % C) * D + j // C, :].contiguous()
        cache["k"] = key
    return cache["v"]


def run_conv(x, conv, bn, relu, out=None, pad=None, precision=None, out_fmt=None, dmajor=None):
    """x channels-last [B,H,W,C] (tensor or ops.Feat) -> fused conv(+bias)+BN(eval)+ReLU."""
    scale, shift = folded_epilogue(conv, bn)
    transposed = isinstance(conv, nn.ConvTranspose2d)
    padding = conv.padding if pad is None else pad
    prec = _precision.resolve(precision)
    if prec != "fp32" and conv.in_channels % 8 != 0:
        prec = "fp32"
    w = conv_weight_kio(conv) if dmajor is None else conv_weight_kio_dmajor(conv, *dmajor)
    return ops.conv2d_nhwc(x, w, conv.kernel_size, conv.stride, padding, scale, shift, relu,
                           out=out, precision=prec, transposed=transposed,
                           out_fmt=out_fmt or act_fmt(prec))


def as_nhwc_feat(x, fmt):
    """Logical NCHW tensor / channels-last Feat -> Feat in the requested inter-layer format."""
    if isinstance(x, ops.Feat):
        if x.fmt == fmt:
            return x
        return ops.to_split(x) if fmt == "split" else ops.Feat(x.to_fp32())
    x = to_nhwc(x)
    return ops.to_split(x) if fmt == "split" else ops.Feat(x)


def to_nhwc(x):
    """Logical NCHW tensor -> channels-last [B,H,W,C] view (copies only if the memory is not already NHWC)."""
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


@NECKS.register_module
class RPN(nn.Module):
    def __init__(self, layer_nums, ds_layer_strides, ds_num_filters, us_layer_strides, us_num_filters,
                 num_input_features, norm_cfg=None, name="rpn", logger=None, **kwargs):
        super().__init__()
        self._layer_strides = ds_layer_strides
        self._num_filters = ds_num_filters
        self._layer_nums = layer_nums
        self._upsample_strides = us_layer_strides
        self._num_upsample_filters = us_num_filters
        self._num_input_features = num_input_features
        if norm_cfg is None:
            norm_cfg = dict(type="BN", eps=1e-3, momentum=0.01)
        self._norm_cfg = norm_cfg
        assert len(ds_layer_strides) == len(layer_nums) == len(ds_num_filters)
        assert len(us_num_filters) == len(us_layer_strides)
        self._upsample_start_idx = len(layer_nums) - len(us_layer_strides)
        self.precision = None     # None -> precision.default_precision()
        ratios = [us_layer_strides[i] / np.prod(ds_layer_strides[: i + self._upsample_start_idx + 1])
                  for i in range(len(us_layer_strides))]
        assert all(r == ratios[0] for r in ratios)

        in_filters = [num_input_features, *ds_num_filters[:-1]]
        blocks, deblocks = [], []
        for i, n in enumerate(layer_nums):
            layers = [nn.ZeroPad2d(1), nn.Conv2d(in_filters[i], ds_num_filters[i], 3, stride=ds_layer_strides[i], bias=False),
                      build_norm_2d(norm_cfg, ds_num_filters[i]), nn.ReLU()]
            for _ in range(n):
                layers += [nn.Conv2d(ds_num_filters[i], ds_num_filters[i], 3, padding=1, bias=False),
                           build_norm_2d(norm_cfg, ds_num_filters[i]), nn.ReLU()]
            blocks.append(nn.Sequential(*layers))
            j = i - self._upsample_start_idx
            if j >= 0:
                stride = us_layer_strides[j]
                if stride > 1:
                    up = nn.ConvTranspose2d(ds_num_filters[i], us_num_filters[j], stride, stride=stride, bias=False)
                else:
                    s = int(np.round(1 / stride))
                    up = nn.Conv2d(ds_num_filters[i], us_num_filters[j], s, stride=s, bias=False)
                deblocks.append(nn.Sequential(up, build_norm_2d(norm_cfg, us_num_filters[j]), nn.ReLU()))
        self.blocks = nn.ModuleList(blocks)
        self.deblocks = nn.ModuleList(deblocks)
        (logger or logging.getLogger("RPN")).info("Finish RPN Initialization")

    @property
    def downsample_factor(self):
        factor = np.prod(self._layer_strides)
        if len(self._upsample_strides) > 0:
            factor /= self._upsample_strides[-1]
        return factor

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)

    def forward(self, x, out_fmt="fp32"):
        """x logical [B,C,H,W] (or a channels-last ops.Feat) -> logical [B, sum(us_num_filters), H', W']
        (channels-last memory); out_fmt="split" (fused pipeline) returns a split-row ops.Feat [B,H',W',C]."""
        prec = _precision.resolve(self.precision)
        fmt = act_fmt(prec)
        dmajor = getattr(x, "bev_dmajor", None)          # set by the fused backbone (channel = d*C + c)
        x = as_nhwc_feat(x, fmt)
        B = x.t.shape[0]
        final_fmt = out_fmt if fmt == "split" else "fp32"
        out = None
        col = 0
        n_up = len(self.deblocks)
        for i, block in enumerate(self.blocks):
            mods = list(block)
            x = ops.as_feat(run_conv(x, mods[1], mods[2], True, pad=(1, 1), precision=prec,
                                     dmajor=dmajor if i == 0 else None))   # ZeroPad2d(1) + conv(pad 0)
            for k in range(4, len(mods), 3):
                x = ops.as_feat(run_conv(x, mods[k], mods[k + 1], True, precision=prec))
            j = i - self._upsample_start_idx
            if j >= 0:
                up, bn = self.deblocks[j][0], self.deblocks[j][1]
                H_, W_ = x.t.shape[1], x.t.shape[2]
                if isinstance(up, nn.ConvTranspose2d):
                    Ho, Wo = H_ * up.stride[0], W_ * up.stride[1]
                else:
                    Ho, Wo = H_ // up.stride[0], W_ // up.stride[1]
                if out is None:
                    out = ops.Feat(torch.empty((B, Ho, Wo, sum(self._num_upsample_filters)), dtype=torch.float32,
                                               device=x.t.device), final_fmt)
                c = self._num_upsample_filters[j]
                run_conv(x, up, bn, True, out=out.slice(col, c), precision=prec)
                col += c
        res = x if n_up == 0 else out
        if res.fmt == "split":
            return res if out_fmt == "split" else res.to_fp32().permute(0, 3, 1, 2)
        return res.to_fp32().permute(0, 3, 1, 2)
